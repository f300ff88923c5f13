#!/bin/bash
# Round 2, call 10 (1 GPU): grouped-GEMM MoE prefill, full GPU suite, Mixtral bench, racecheck on the shared-memory heavy kernels.
mkdir -p gpurun_out
S=gpurun_out/r2_10_summary.txt; : > $S
timeout 600 python -m pytest tests/test_kernels_gpu.py -q --timeout=150 -k "moe" > gpurun_out/r2_10_moe_tests.log 2>&1; echo "moe tests exit=$?" | tee -a $S
tail -15 gpurun_out/r2_10_moe_tests.log | cut -c1-300 | tee -a $S
timeout 1500 python -m pytest tests -m gpu -q --timeout=300 -x > gpurun_out/r2_10_gpu_suite.log 2>&1; echo "gpu suite exit=$?" | tee -a $S
tail -8 gpurun_out/r2_10_gpu_suite.log | cut -c1-300 | tee -a $S
run() { name=$1; shift
  timeout 600 python bench.py --steps 24 --warmup 4 --skip-fp8 "$@" > gpurun_out/r2_10_$name.log 2>&1; echo "$name exit=$?" | tee -a $S
  grep '^{' gpurun_out/r2_10_$name.log | python -c "import sys,json; [print({k:d[k] for k in ('value','ms_per_step','gpu_launches') if k in d}, d.get('e2e',{}).get('value'), d.get('roofline',{}).get('frac_of_measured_hbm'), json.dumps(d.get('prefill'))[:400]) for d in map(json.loads, sys.stdin)]" | tee -a $S
  grep -iE "error|Traceback" gpurun_out/r2_10_$name.log | head -3 | cut -c1-300 | tee -a $S
}
run mixtral --model mixtral-8x7b
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 20 python -m pytest tests/test_kernels_gpu.py -q --timeout=800 -x -k "moe_prefill or attention_bwd or rmsnorm_bwd or gemm_plain" > gpurun_out/r2_10_racecheck.log 2>&1; echo "racecheck exit=$?" | tee -a $S
grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/r2_10_racecheck.log | tail -8 | cut -c1-300 | tee -a $S
