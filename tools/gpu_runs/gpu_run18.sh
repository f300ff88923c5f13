#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/summary18.txt; : > $S
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/t18_all.log 2>&1; echo "pytest -m gpu exit=$?" | tee -a $S
tail -6 gpurun_out/t18_all.log | cut -c1-400 | tee -a $S
timeout 600 python bench.py --steps 64 --warmup 4 > gpurun_out/b18_70b.log 2>&1; echo "bench 70b exit=$?" | tee -a $S
grep -E "^\{" gpurun_out/b18_70b.log | tail -1 | cut -c1-3000 | tee -a $S
timeout 600 python benchmarks/benchmark_training.py --model llama-3-8b --n_steps 5 --batch_size 8 --seq_len 128 > gpurun_out/train_8b.log 2>&1; echo "benchmark_training 8b exit=$?" | tee -a $S
grep "Final result" gpurun_out/train_8b.log | tee -a $S
timeout 600 python benchmarks/benchmark_forward.py --model llama-3-8b --n_steps 10 --batch_size 8 --seq_len 512 > gpurun_out/fwd_8b.log 2>&1; echo "benchmark_forward 8b exit=$?" | tee -a $S
grep -i "final\|tokens/sec" gpurun_out/fwd_8b.log | tail -2 | tee -a $S
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke18.log 2>&1; echo "smoke exit=$?" | tee -a $S
