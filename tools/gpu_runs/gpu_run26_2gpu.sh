#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/summary26.txt; : > $S
timeout 900 python -m pytest tests/test_multi_gpu.py -q -m gpu -x > gpurun_out/t26_multi.log 2>&1; echo "test_multi_gpu exit=$?" | tee -a $S
tail -5 gpurun_out/t26_multi.log | cut -c1-600 | tee -a $S
