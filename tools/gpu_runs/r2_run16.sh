#!/bin/bash
# Round 2, call 16 (1 GPU): the 2-CTA (cta_group::2) GEMM — numerics vs the 1-CTA kernel, then the kernel table vs 1-CTA and cuBLAS.
mkdir -p gpurun_out
S=gpurun_out/r2_16_summary.txt; : > $S
export PETALS_B200_RUN_UNVALIDATED=1
timeout 300 python -m pytest tests/test_kernels_gpu.py -q --timeout=120 -x -k "2cta or ll_collectives" > gpurun_out/r2_16_tests.log 2>&1; echo "2cta tests exit=$?" | tee -a $S
timeout 300 python -m pytest tests/test_fp8_gpu.py -q --timeout=120 -k "2cta" > gpurun_out/r2_16_tests_fp8.log 2>&1; echo "fp8 2cta tests exit=$?" | tee -a $S
grep -E "passed|failed|Error|assert|stuck|watchdog" gpurun_out/r2_16_tests_fp8.log | tail -8 | cut -c1-300 | tee -a $S
grep -E "passed|failed|Error|assert|stuck" gpurun_out/r2_16_tests.log | tail -8 | cut -c1-300 | tee -a $S
timeout 600 python tools/engine_error_stats.py 2>/dev/null | grep "^{" > gpurun_out/r2_16_engine_error_stats.txt; cat gpurun_out/r2_16_engine_error_stats.txt | cut -c1-400 | tee -a $S
if grep -q "passed" gpurun_out/r2_16_tests.log && ! grep -q "failed" gpurun_out/r2_16_tests.log; then
  timeout 600 python tools/kernel_bench.py --only gemm_2cta > gpurun_out/r2_16_kernel_bench.log 2>&1; echo "kernel bench exit=$?" | tee -a $S
  grep "gemm2" gpurun_out/r2_16_kernel_bench.log | cut -c1-300 | tee -a $S
fi
if grep -q "passed" gpurun_out/r2_16_tests_fp8.log && ! grep -q "failed" gpurun_out/r2_16_tests_fp8.log; then
  PETALS_B200_GEMM_2CTA=1 timeout 600 python tools/kernel_bench.py --only gemm_fp8 > gpurun_out/r2_16_kernel_bench_fp8_2cta.log 2>&1; echo "fp8 2cta kernel bench exit=$?" | tee -a $S
  grep "gemm_mxfp8" gpurun_out/r2_16_kernel_bench_fp8_2cta.log | cut -c1-300 | tee -a $S
fi
