#!/bin/bash
# Round 2, call 2: first hardware run of the persistent span kernel (csrc/decode_span.cu): numerics, then decode rate on one GPU
# (70B, 8B) and for one rank's share of tp8 / tp4 / tp2.
mkdir -p gpurun_out
S=gpurun_out/r2_2_summary.txt; : > $S
timeout 600 python -m pytest tests/test_decode_span_gpu.py -q -x --timeout=150 > gpurun_out/r2_2_span_tests.log 2>&1; echo "span tests exit=$?" | tee -a $S
tail -25 gpurun_out/r2_2_span_tests.log | cut -c1-300 | tee -a $S
timeout 600 python -m pytest tests -q -m gpu -x --timeout=150 > gpurun_out/r2_2_pytest.log 2>&1; echo "pytest -m gpu exit=$?" | tee -a $S
tail -4 gpurun_out/r2_2_pytest.log | cut -c1-300 | tee -a $S
run() { # name, extra args...
  name=$1; shift
  timeout 600 python bench.py --steps 24 --warmup 4 --skip-fp8 "$@" > gpurun_out/r2_2_$name.log 2>&1; echo "$name exit=$?" | tee -a $S
  grep '^{' gpurun_out/r2_2_$name.log | python -c "import sys,json; [print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches') if k in d}, d.get('roofline',{}).get('frac_of_measured_hbm'), (d.get('prefill') or {}).get('tokens_per_s')) for d in map(json.loads, sys.stdin)]" | tee -a $S
  grep -iE "error|Traceback" gpurun_out/r2_2_$name.log | head -5 | tee -a $S
}
run tp8emu_span --tp-emulate 8 --skip-prefill
PETALS_B200_SPAN_KERNEL=0 run tp8emu_nospan --tp-emulate 8 --skip-prefill
run tp4emu_span --tp-emulate 4 --skip-prefill
run tp2emu_span --tp-emulate 2 --skip-prefill
run 70b_span
run 8b_span --model llama-3-8b --skip-prefill
PETALS_B200_SPAN_KERNEL=0 run 8b_nospan --model llama-3-8b --skip-prefill
