#!/bin/bash
# Round 2, call 7: span kernel v4 (L2 prefetch-ahead while the ring is full, norm statistics fused into the gathers, batched combine)
mkdir -p gpurun_out
S=gpurun_out/r2_7_summary.txt; : > $S
timeout 600 python -m pytest tests/test_decode_span_gpu.py -q --timeout=150 > gpurun_out/r2_7_span_tests.log 2>&1; echo "span tests exit=$?" | tee -a $S
tail -3 gpurun_out/r2_7_span_tests.log | cut -c1-250 | tee -a $S
for shape in 70b-tp8 70b 8b; do
  for pf in 0; do
    echo "PF=$pf" | tee -a $S
    PETALS_B200_SPAN_PF=$pf timeout 300 python tools/span_probe.py --shape $shape 2>&1 | grep '^{' | tee -a $S
  done
done
run() { name=$1; shift
  timeout 600 python bench.py --steps 24 --warmup 4 --skip-fp8 "$@" > gpurun_out/r2_7_$name.log 2>&1; echo "$name exit=$?" | tee -a $S
  grep '^{' gpurun_out/r2_7_$name.log | python -c "import sys,json; [print({k:d[k] for k in ('value','ms_per_step','gpu_launches') if k in d}, d.get('e2e',{}).get('value'), d.get('roofline',{}).get('frac_of_measured_hbm'), (d.get('prefill') or {}).get('tokens_per_s')) for d in map(json.loads, sys.stdin)]" | tee -a $S
  grep -iE "error|Traceback" gpurun_out/r2_7_$name.log | head -3 | cut -c1-300 | tee -a $S
}
run tp8emu_span --tp-emulate 8 --skip-prefill
run 70b_span --skip-prefill
run 8b_span --model llama-3-8b --skip-prefill
