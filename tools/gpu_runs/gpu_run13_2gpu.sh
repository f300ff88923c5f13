#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/summary13.txt; : > $S
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29601 tools/tp_selftest.py > gpurun_out/tp_selftest_$N.log 2>&1; echo "tp_selftest($N) exit=$?" | tee -a $S
grep -E "^\{" gpurun_out/tp_selftest_$N.log | tail -2 | cut -c1-700 | tee -a $S
timeout 900 $TR --master-port 29604 bench.py --gpus $N --steps 32 --warmup 4 > gpurun_out/bench_70b_tp$N.log 2>&1; echo "bench 70b tp$N exit=$?" | tee -a $S
grep -E "^\{" gpurun_out/bench_70b_tp$N.log | tail -1 | cut -c1-3000 | tee -a $S
if [ "$N" != "2" ]; then
  timeout 300 $TR --master-port 29602 tools/pp_selftest.py > gpurun_out/pp_selftest_$N.log 2>&1; echo "pp_selftest($N) exit=$?" | tee -a $S
  grep -E "^\{" gpurun_out/pp_selftest_$N.log | tail -1 | cut -c1-600 | tee -a $S
  timeout 600 $TR --master-port 29605 bench.py --gpus $N --steps 32 --warmup 4 --parallelism pp$N > gpurun_out/bench_70b_pp$N.log 2>&1; echo "bench 70b pp$N exit=$?" | tee -a $S
  grep -E "^\{" gpurun_out/bench_70b_pp$N.log | tail -1 | cut -c1-2500 | tee -a $S
fi
