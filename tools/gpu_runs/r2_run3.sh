#!/bin/bash
# Round 2, call 3: (a) where the span kernel's time goes (phase stamps; polls off / math off), (b) first hardware run of the backward kernels.
mkdir -p gpurun_out
S=gpurun_out/r2_3_summary.txt; : > $S
for shape in 70b-tp8 8b 70b; do
  for dbg in 0 1 2 3; do
    PETALS_B200_SPAN_DEBUG=$dbg timeout 300 python tools/span_probe.py --shape $shape 2>&1 | grep '^{' | tee -a $S
  done
done
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x --timeout=120 -k "backward or rope_backward" > gpurun_out/r2_3_bwd_kernels.log 2>&1; echo "bwd kernel tests exit=$?" | tee -a $S
tail -15 gpurun_out/r2_3_bwd_kernels.log | cut -c1-250 | tee -a $S
timeout 600 python -m pytest tests/test_engine_gpu.py -q -x --timeout=200 -k "backward" > gpurun_out/r2_3_bwd_engine.log 2>&1; echo "bwd engine tests exit=$?" | tee -a $S
tail -15 gpurun_out/r2_3_bwd_engine.log | cut -c1-250 | tee -a $S
timeout 600 python -m pytest tests/test_decode_span_gpu.py -q --timeout=150 > gpurun_out/r2_3_span_tests.log 2>&1; echo "span tests exit=$?" | tee -a $S
tail -5 gpurun_out/r2_3_span_tests.log | cut -c1-250 | tee -a $S
for tc in 1 0; do
  PETALS_B200_ENGINE_BACKWARD=$tc timeout 600 python benchmarks/benchmark_training.py --model llama-3-8b --n_steps 8 --warmup_steps 3 --batch_size 8 --seq_len 128 > gpurun_out/r2_3_train_$tc.log 2>&1
  echo "ENGINE_BACKWARD=$tc $(grep 'Final result' gpurun_out/r2_3_train_$tc.log) $(grep -iE 'Error|Traceback' gpurun_out/r2_3_train_$tc.log | head -2)" | tee -a $S
done
