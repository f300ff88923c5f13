#!/bin/bash
# 2-GPU validation: backward test, fp8 kernels, tensor-parallel and pipeline (fused hop) self-tests, TP/PP benches.
mkdir -p gpurun_out
S=gpurun_out/summary5.txt; : > $S
nvidia-smi --query-gpu=index,name --format=csv,noheader | tee -a $S
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 240 python -m pytest tests/test_engine_gpu.py -q -m gpu -k "backward or generate" > gpurun_out/t_engine.log 2>&1; echo "engine(backward,generate) exit=$?" | tee -a $S
tail -3 gpurun_out/t_engine.log | cut -c1-300
timeout 240 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "fp8 or dequant or reduce_parts" > gpurun_out/t_fp8.log 2>&1; echo "fp8 kernels exit=$?" | tee -a $S
tail -3 gpurun_out/t_fp8.log | cut -c1-300
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29601 tools/tp_selftest.py > gpurun_out/tp_selftest.log 2>&1; echo "tp_selftest exit=$?" | tee -a $S
grep -E "tp_selftest|Error|error" gpurun_out/tp_selftest.log | tail -5 | cut -c1-600
timeout 300 $TR --master-port 29602 tools/pp_selftest.py > gpurun_out/pp_selftest.log 2>&1; echo "pp_selftest exit=$?" | tee -a $S
grep -E "pp_selftest|Error|error" gpurun_out/pp_selftest.log | tail -5 | cut -c1-600
timeout 400 $TR --master-port 29603 bench.py --gpus 2 --model llama-3-8b --steps 32 --warmup 4 > gpurun_out/bench_8b_tp2.log 2>&1; echo "bench 8b tp2 exit=$?" | tee -a $S
tail -1 gpurun_out/bench_8b_tp2.log | cut -c1-2500
timeout 600 $TR --master-port 29604 bench.py --gpus 2 --steps 32 --warmup 4 > gpurun_out/bench_70b_tp2.log 2>&1; echo "bench 70b tp2 exit=$?" | tee -a $S
tail -1 gpurun_out/bench_70b_tp2.log | cut -c1-2500
timeout 600 $TR --master-port 29605 bench.py --gpus 2 --steps 32 --warmup 4 --parallelism pp2 > gpurun_out/bench_70b_pp2.log 2>&1; echo "bench 70b pp2 exit=$?" | tee -a $S
tail -1 gpurun_out/bench_70b_pp2.log | cut -c1-2500
