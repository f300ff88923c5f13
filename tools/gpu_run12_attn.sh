#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/summary12.txt; : > $S
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "attention" > gpurun_out/t12_attn.log 2>&1; echo "attention tests exit=$?" | tee -a $S
tail -15 gpurun_out/t12_attn.log | cut -c1-400
timeout 120 python tools/profile_kernels.py attn_time > gpurun_out/attn_time.log 2>&1; echo "attn_time exit=$?" | tee -a $S
tail -3 gpurun_out/attn_time.log | cut -c1-600 | tee -a $S
