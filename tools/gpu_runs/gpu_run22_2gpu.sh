#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/summary22.txt; : > $S
timeout 900 python -m pytest tests/test_multi_gpu.py -q -m gpu -x > gpurun_out/t22_multi.log 2>&1; echo "test_multi_gpu exit=$?" | tee -a $S
tail -15 gpurun_out/t22_multi.log | cut -c1-600 | tee -a $S
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29604 bench.py --gpus 2 --steps 32 --warmup 4 > gpurun_out/bench_70b_tp2.log 2>&1; echo "bench 70b tp2 exit=$?" | tee -a $S
grep -E "^\{" gpurun_out/bench_70b_tp2.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['prefill'])" | tee -a $S
PETALS_B200_CHAIN=1 timeout 600 $TR --master-port 29605 bench.py --gpus 2 --steps 32 --warmup 4 --skip-prefill > gpurun_out/bench_70b_tp2_chain.log 2>&1; echo "bench tp2 chain exit=$?" | tee -a $S
grep -E "^\{" gpurun_out/bench_70b_tp2_chain.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('chain', d['ms_per_step'], d['value'])" | tee -a $S
