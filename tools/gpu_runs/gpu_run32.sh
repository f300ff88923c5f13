#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/summary32.txt; : > $S
timeout 200 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "skinny" > gpurun_out/t32.log 2>&1; echo "skinny tests exit=$?" | tee -a $S
tail -12 gpurun_out/t32.log | cut -c1-500 | tee -a $S
timeout 200 python tools/kernel_bench.py --only skinny > gpurun_out/kb32.log 2>&1; echo "kernel_bench skinny exit=$?" | tee -a $S
grep "gemv" gpurun_out/kb32.log | cut -c1-200 | tee -a $S
tail -3 gpurun_out/kb32.log | cut -c1-300
