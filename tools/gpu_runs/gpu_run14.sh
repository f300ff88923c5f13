#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/summary14.txt; : > $S
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "attention or fused_rope" > gpurun_out/t14_attn.log 2>&1; echo "attention+rope tests exit=$?" | tee -a $S
tail -5 gpurun_out/t14_attn.log | cut -c1-400
timeout 120 python tools/profile_kernels.py attn_time > gpurun_out/attn_time.log 2>&1; echo "attn_time exit=$?" | tee -a $S
tail -3 gpurun_out/attn_time.log | cut -c1-600 | tee -a $S
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_tc -s 2 -c 1 -o gpurun_out/ncu_attn_tc -f env ATTN_IMPL=2 python tools/profile_kernels.py attn > gpurun_out/ncu_attn_tc.log 2>&1; echo "ncu exit=$?" | tee -a $S
bash tools/gpu_runs/gpu_run15.sh
