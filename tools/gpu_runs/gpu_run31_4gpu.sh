#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/summary31.txt; : > $S
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=4 --master-addr 127.0.0.1"
TP_SELFTEST_HIDDEN=2048 timeout 300 $TR --master-port 29601 tools/tp_selftest.py > gpurun_out/tp_selftest_4.log 2>&1; echo "tp_selftest(4, hidden 2048) exit=$?" | tee -a $S
grep -E "^\{" gpurun_out/tp_selftest_4.log | tail -1 | cut -c1-600 | tee -a $S
timeout 300 $TR --master-port 29604 bench.py --gpus 4 --steps 32 --warmup 4 --skip-prefill > gpurun_out/bench_70b_tp4.log 2>&1; echo "bench 70b tp4 exit=$?" | tee -a $S
grep -E "^\{" gpurun_out/bench_70b_tp4.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['value'])" | tee -a $S
