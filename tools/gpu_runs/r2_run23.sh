#!/bin/bash
# Round 2, call 23 (1 GPU): the very last tree — full GPU suite and smoke() (handler / backend changed after call 19), short bench.
mkdir -p gpurun_out
S=gpurun_out/r2_23_summary.txt; : > $S
timeout 900 python -m pytest tests -m gpu -q --timeout=300 -x > gpurun_out/r2_23_gpu_suite.log 2>&1; echo "gpu suite exit=$?" | tee -a $S
tail -4 gpurun_out/r2_23_gpu_suite.log | cut -c1-300 | tee -a $S
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_23_smoke.log 2>&1; echo "smoke exit=$?" | tee -a $S
tail -1 gpurun_out/r2_23_smoke.log | cut -c1-300 | tee -a $S
timeout 600 python bench.py --steps 24 --warmup 4 --skip-fp8 > gpurun_out/r2_23_bench.log 2>&1; echo "bench exit=$?" | tee -a $S
grep '^{' gpurun_out/r2_23_bench.log | python -c "import sys,json; [print({k:d[k] for k in ('value','ms_per_step','gpu_launches') if k in d}, 'e2e', d.get('e2e',{}).get('value'), 'prefill', json.dumps(d.get('prefill'))[:200]) for d in map(json.loads, sys.stdin)]" | tee -a $S
