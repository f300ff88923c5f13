#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/summary7.txt; : > $S
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x > gpurun_out/t7_kernels.log 2>&1; echo "kernels exit=$?" | tee -a $S
tail -5 gpurun_out/t7_kernels.log | cut -c1-300
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_fp8_gpu.py -q -m gpu -x > gpurun_out/t7_engine.log 2>&1; echo "engine exit=$?" | tee -a $S
tail -5 gpurun_out/t7_engine.log | cut -c1-300
timeout 600 python bench.py --steps 64 --warmup 4 --skip-prefill > gpurun_out/bench7_70b_pdl.log 2>&1; echo "bench pdl exit=$?" | tee -a $S
grep -E "^\{" gpurun_out/bench7_70b_pdl.log | tail -1 | cut -c1-700
PETALS_B200_PDL=0 timeout 600 python bench.py --steps 64 --warmup 4 --skip-prefill > gpurun_out/bench7_70b_nopdl.log 2>&1; echo "bench nopdl exit=$?" | tee -a $S
grep -E "^\{" gpurun_out/bench7_70b_nopdl.log | tail -1 | cut -c1-700
timeout 300 python bench.py --model llama-3-8b --steps 64 --warmup 4 --skip-prefill > gpurun_out/bench7_8b_pdl.log 2>&1; echo "bench 8b pdl exit=$?" | tee -a $S
grep -E "^\{" gpurun_out/bench7_8b_pdl.log | tail -1 | cut -c1-500
PETALS_B200_PDL=0 timeout 300 python bench.py --model llama-3-8b --steps 64 --warmup 4 --skip-prefill > gpurun_out/bench7_8b_nopdl.log 2>&1; echo "bench 8b nopdl exit=$?" | tee -a $S
grep -E "^\{" gpurun_out/bench7_8b_nopdl.log | tail -1 | cut -c1-500
