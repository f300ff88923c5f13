#!/bin/bash
# Round 2, call 12 (1 GPU): block-scaled MXFP8 GEMM — numerics tests, kernel bench vs the bf16 tcgen05 GEMM, fp8 engine test + bench.
mkdir -p gpurun_out
S=gpurun_out/r2_12_summary.txt; : > $S
timeout 600 python -m pytest tests/test_fp8_gpu.py -q --timeout=200 > gpurun_out/r2_12_fp8_tests.log 2>&1; echo "fp8 tests exit=$?" | tee -a $S
grep -E "passed|failed|Error|assert" gpurun_out/r2_12_fp8_tests.log | tail -12 | cut -c1-300 | tee -a $S
timeout 600 python tools/kernel_bench.py --only gemm_fp8 > gpurun_out/r2_12_kernel_bench_fp8.log 2>&1; echo "kernel bench exit=$?" | tee -a $S
grep "gemm_mxfp8" gpurun_out/r2_12_kernel_bench_fp8.log | cut -c1-300 | tee -a $S
tail -3 gpurun_out/r2_12_kernel_bench_fp8.log | grep -iE "error|Traceback" | cut -c1-300 | tee -a $S
run() { name=$1; shift
  timeout 900 python bench.py --steps 16 --warmup 3 "$@" > gpurun_out/r2_12_$name.log 2>&1; echo "$name exit=$?" | tee -a $S
  grep '^{' gpurun_out/r2_12_$name.log | python -c "import sys,json; [print({k:d[k] for k in ('value','ms_per_step','gpu_launches') if k in d}, 'prefill', json.dumps(d.get('prefill'))[:300], 'fp8', json.dumps(d.get('fp8_weights'))[:600]) for d in map(json.loads, sys.stdin)]" | tee -a $S
  grep -iE "error|Traceback" gpurun_out/r2_12_$name.log | head -3 | cut -c1-300 | tee -a $S
}
run 70b_fp8
