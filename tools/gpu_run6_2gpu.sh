#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/summary6.txt; : > $S
timeout 400 python -m pytest tests/test_engine_gpu.py tests/test_fp8_gpu.py -q -m gpu -k "mixtral or fp8" > gpurun_out/t_moe_fp8.log 2>&1; echo "engine(mixtral)+fp8 stage exit=$?" | tee -a $S
tail -4 gpurun_out/t_moe_fp8.log | cut -c1-300
timeout 200 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "moe or fp8 or linear_decode_plain" > gpurun_out/t_k.log 2>&1; echo "kernels(moe,fp8,gemv) exit=$?" | tee -a $S
tail -3 gpurun_out/t_k.log | cut -c1-300
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29601 tools/tp_selftest.py > gpurun_out/tp_selftest.log 2>&1; echo "tp_selftest exit=$?" | tee -a $S
grep -E "^\{" gpurun_out/tp_selftest.log | tail -2 | cut -c1-600
timeout 300 $TR --master-port 29602 tools/pp_selftest.py > gpurun_out/pp_selftest.log 2>&1; echo "pp_selftest exit=$?" | tee -a $S
grep -E "^\{|Error" gpurun_out/pp_selftest.log | tail -3 | cut -c1-600
timeout 400 $TR --master-port 29603 bench.py --gpus 2 --model llama-3-8b --steps 32 --warmup 4 > gpurun_out/bench_8b_tp2.log 2>&1; echo "bench 8b tp2 exit=$?" | tee -a $S
grep -E "^\{" gpurun_out/bench_8b_tp2.log | tail -1 | cut -c1-2500
timeout 600 $TR --master-port 29604 bench.py --gpus 2 --steps 32 --warmup 4 > gpurun_out/bench_70b_tp2.log 2>&1; echo "bench 70b tp2 exit=$?" | tee -a $S
grep -E "^\{" gpurun_out/bench_70b_tp2.log | tail -1 | cut -c1-2500
timeout 600 $TR --master-port 29605 bench.py --gpus 2 --steps 32 --warmup 4 --parallelism pp2 > gpurun_out/bench_70b_pp2.log 2>&1; echo "bench 70b pp2 exit=$?" | tee -a $S
grep -E "^\{" gpurun_out/bench_70b_pp2.log | tail -1 | cut -c1-2500
