#!/bin/bash
# Round 2, call 6: locate the span-kernel v2 stall on real shapes (contextual watchdog), LoRA tests.
mkdir -p gpurun_out
S=gpurun_out/r2_6_summary.txt; : > $S
for shape in 8b 70b-tp8; do
  timeout 120 python tools/span_probe.py --shape $shape --layers 3 --steps 2 > gpurun_out/r2_6_probe_$shape.log 2>&1; echo "probe $shape exit=$?" | tee -a $S
  grep '^{' gpurun_out/r2_6_probe_$shape.log | tee -a $S
  grep "decode_span stuck" gpurun_out/r2_6_probe_$shape.log | sort | uniq -c | sort -rn | head -40 | tee -a $S
done
timeout 300 python -m pytest tests/test_lora_engine_gpu.py -q --timeout=120 > gpurun_out/r2_6_lora.log 2>&1; echo "lora exit=$?" | tee -a $S
grep -E "passed|failed|Error|assert" gpurun_out/r2_6_lora.log | tail -8 | cut -c1-300 | tee -a $S
