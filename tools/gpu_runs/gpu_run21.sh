#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/summary21.txt; : > $S
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_fp8_gpu.py -q -m gpu -x -k "fp8" > gpurun_out/t21_k.log 2>&1; echo "fp8 tests exit=$?" | tee -a $S
tail -3 gpurun_out/t21_k.log | cut -c1-400 | tee -a $S
python tools/kernel_bench.py --fp8 > gpurun_out/kbench_fp8.log 2>&1; echo "kernel_bench fp8 exit=$?" | tee -a $S
grep "gemv.fp8" gpurun_out/kbench_fp8.log | cut -c1-200 | tee -a $S
timeout 600 python bench.py --steps 32 --warmup 4 > gpurun_out/b21_full_70b.log 2>&1; echo "bench full exit=$?" | tee -a $S
grep -E "^\{" gpurun_out/b21_full_70b.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['prefill'], d['fp8_weights'])" | tee -a $S
timeout 300 python bench.py --model llama-3-8b --steps 32 --warmup 4 --skip-prefill > gpurun_out/b21_8b.log 2>&1
grep -E "^\{" gpurun_out/b21_8b.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('8b', d['ms_per_step'], d['value'], d['fp8_weights'])" | tee -a $S
