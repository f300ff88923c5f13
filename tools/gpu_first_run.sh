#!/bin/bash
# First GPU contact: kernel numerics (one pytest process per test group so a device trap cannot take
# the rest down) + micro-benchmarks. Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
for k in linear_decode gemm_plain gemm_block gemm_bias gemm_swiglu gemm_b_transposed gemm_fp32 norm elementwise rope_kv attention_alibi; do
  timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -m gpu -k "$k" > gpurun_out/test_$k.log 2>&1
  echo "$k exit=$?" | tee -a gpurun_out/summary.txt
  tail -3 gpurun_out/test_$k.log
done
timeout 600 python tools/kernel_bench.py > gpurun_out/kernel_bench.log 2>&1; echo "bench exit=$?" | tee -a gpurun_out/summary.txt
tail -40 gpurun_out/kernel_bench.log
