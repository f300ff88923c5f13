#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/summary28.txt; : > $S
for tc in 0 1 0 1; do
  PETALS_B200_TC_BACKWARD=$tc timeout 600 python benchmarks/benchmark_training.py --model llama-3-8b --n_steps 8 --warmup_steps 3 --batch_size 8 --seq_len 128 > gpurun_out/train28.log 2>&1
  echo "TC_BACKWARD=$tc $(grep 'Final result' gpurun_out/train28.log)" | tee -a $S
done
