#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/summary11.txt; : > $S
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29601 tools/tp_selftest.py > gpurun_out/tp_selftest.log 2>&1; echo "tp_selftest exit=$?" | tee -a $S
grep -E "^\{" gpurun_out/tp_selftest.log | tail -2 | cut -c1-600
timeout 300 $TR --master-port 29602 tools/pp_selftest.py > gpurun_out/pp_selftest.log 2>&1; echo "pp_selftest exit=$?" | tee -a $S
grep -E "^\{|Error" gpurun_out/pp_selftest.log | tail -3 | cut -c1-600
timeout 900 $TR --master-port 29604 bench.py --gpus 2 --steps 32 --warmup 4 > gpurun_out/bench_70b_tp2.log 2>&1; echo "bench 70b tp2 exit=$?" | tee -a $S
grep -E "^\{" gpurun_out/bench_70b_tp2.log | tail -1 | cut -c1-3000
PETALS_B200_PDL_MASK=31 timeout 600 $TR --master-port 29606 bench.py --gpus 2 --steps 32 --warmup 4 --skip-prefill > gpurun_out/bench_70b_tp2_m31.log 2>&1; echo "bench 70b tp2 mask31 exit=$?" | tee -a $S
grep -E "^\{" gpurun_out/bench_70b_tp2_m31.log | tail -1 | cut -c1-400
PETALS_B200_PDL_MASK=0 PETALS_B200_PF_LINES_WAIT=0 timeout 600 $TR --master-port 29607 bench.py --gpus 2 --steps 32 --warmup 4 --skip-prefill > gpurun_out/bench_70b_tp2_m0.log 2>&1; echo "bench 70b tp2 mask0 pf0 exit=$?" | tee -a $S
grep -E "^\{" gpurun_out/bench_70b_tp2_m0.log | tail -1 | cut -c1-400
timeout 600 $TR --master-port 29605 bench.py --gpus 2 --steps 32 --warmup 4 --parallelism pp2 > gpurun_out/bench_70b_pp2.log 2>&1; echo "bench 70b pp2 exit=$?" | tee -a $S
grep -E "^\{" gpurun_out/bench_70b_pp2.log | tail -1 | cut -c1-2500
for R in 2 4 8; do timeout 200 python bench.py --tp-emulate $R --steps 64 --warmup 4 --skip-prefill --skip-fp8 > gpurun_out/emul_tp$R.log 2>&1; echo "emulate tp$R: $(grep -E '^\{' gpurun_out/emul_tp$R.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'])")" | tee -a $S; done
