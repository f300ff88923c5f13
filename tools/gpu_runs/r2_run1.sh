#!/bin/bash
# Round 2, call 1: does the tree still pass on hardware after the refactors; first hardware run of the round-1 leftovers;
# per-kernel launch lists of a decode step (1 GPU, and one rank's share of tp8) to see where the boundary time goes.
mkdir -p gpurun_out
S=gpurun_out/r2_1_summary.txt; : > $S
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader | tee -a $S
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/r2_1_pytest.log 2>&1; echo "pytest -m gpu exit=$?" | tee -a $S
tail -5 gpurun_out/r2_1_pytest.log | cut -c1-300 | tee -a $S
PETALS_B200_RUN_UNVALIDATED=1 PETALS_B200_LORA_ENGINE=1 timeout 600 python -m pytest tests/test_l2_prefetch_gpu.py tests/test_lora_engine_gpu.py -q > gpurun_out/r2_1_unvalidated.log 2>&1; echo "unvalidated exit=$?" | tee -a $S
tail -5 gpurun_out/r2_1_unvalidated.log | cut -c1-300 | tee -a $S
timeout 900 python bench.py --steps 24 --warmup 4 > gpurun_out/r2_1_bench70b.log 2>&1; echo "bench exit=$?" | tee -a $S
grep '^{' gpurun_out/r2_1_bench70b.log | cut -c1-900 | tee -a $S
timeout 600 python bench.py --steps 24 --warmup 4 --tp-emulate 8 --skip-prefill --skip-fp8 > gpurun_out/r2_1_tp8emu.log 2>&1
grep '^{' gpurun_out/r2_1_tp8emu.log | cut -c1-400 | tee -a $S
# launch lists (serialised, cold cache: compare shares)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2600 -c 420 --csv --log-file gpurun_out/r2_1_launches_tp8emu.csv \
  python bench.py --steps 6 --warmup 3 --tp-emulate 8 --skip-prefill --skip-fp8 > gpurun_out/r2_1_ncu_tp8emu.log 2>&1; echo "ncu tp8emu exit=$?" | tee -a $S
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2600 -c 420 --csv --log-file gpurun_out/r2_1_launches_70b.csv \
  python bench.py --steps 6 --warmup 3 --skip-prefill --skip-fp8 > gpurun_out/r2_1_ncu_70b.log 2>&1; echo "ncu 70b exit=$?" | tee -a $S
# compute-sanitizer memcheck over the kernel numerics tests (bounded)
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py -q -x -k "not tp and not chain" > gpurun_out/r2_1_memcheck.log 2>&1; echo "memcheck exit=$?" | tee -a $S
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2_1_memcheck.log | tail -3 | tee -a $S
