#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/summary23.txt; : > $S
for N in 8 4; do
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1"
  timeout 600 $TR --master-port 2960$N bench.py --gpus $N --steps 32 --warmup 4 > gpurun_out/bench_70b_tp$N.log 2>&1; echo "bench 70b tp$N exit=$?" | tee -a $S
  grep -E "^\{" gpurun_out/bench_70b_tp$N.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['prefill'])" | tee -a $S
done
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29611 bench.py --gpus 8 --model llama-3-8b --steps 32 --warmup 4 --skip-prefill > gpurun_out/bench_8b_tp8.log 2>&1; echo "bench 8b tp8 exit=$?" | tee -a $S
grep -E "^\{" gpurun_out/bench_8b_tp8.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('8b tp8', d['ms_per_step'], d['value'])" | tee -a $S
timeout 300 $TR --master-port 29612 bench.py --gpus 8 --model llama-3-8b --steps 32 --warmup 4 --parallelism pp8 > gpurun_out/bench_8b_pp8.log 2>&1; echo "bench 8b pp8 exit=$?" | tee -a $S
grep -E "^\{" gpurun_out/bench_8b_pp8.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('8b pp8', d['ms_per_step'], d['value'])" | tee -a $S
