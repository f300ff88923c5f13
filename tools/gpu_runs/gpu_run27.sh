#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/summary27.txt; : > $S
timeout 600 python -m pytest tests/test_engine_gpu.py -q -m gpu -x > gpurun_out/t27_engine.log 2>&1; echo "engine tests exit=$?" | tee -a $S
tail -12 gpurun_out/t27_engine.log | cut -c1-500 | tee -a $S
for tc in 1 0; do
  PETALS_B200_TC_BACKWARD=$tc timeout 600 python benchmarks/benchmark_training.py --model llama-3-8b --n_steps 5 --batch_size 8 --seq_len 128 > gpurun_out/train27_tc$tc.log 2>&1; echo "benchmark_training 8b TC_BACKWARD=$tc exit=$?" | tee -a $S
  grep "Final result" gpurun_out/train27_tc$tc.log | tee -a $S
  tail -3 gpurun_out/train27_tc$tc.log | cut -c1-300
done
