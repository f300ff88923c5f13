#!/bin/bash
# Round 2, call 22 (4 GPUs): pipeline of tensor-parallel pairs with the leaders' NVLink fabric (2 x TP2), with and without it; tp4 default line.
mkdir -p gpurun_out
S=gpurun_out/r2_22_summary.txt; : > $S
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
run() { name=$1; shift
  timeout 420 $TR --master-port 29742 bench.py --gpus 4 --steps 24 --warmup 4 "$@" > gpurun_out/r2_22_$name.log 2>&1; echo "$name exit=$?" | tee -a $S
  grep '^{' gpurun_out/r2_22_$name.log | python -c "import sys,json
for d in map(json.loads, sys.stdin):
    print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches','invalid')}, 'e2e', d.get('e2e',{}).get('value'), 'prefill', json.dumps(d.get('prefill'))[:300])
    print('   config:', json.dumps(d.get('config'))[:500])
    print('   selftests:', json.dumps(d.get('selftests'))[:300], 'pipeline:', json.dumps(d.get('pipeline'))[:700])" | tee -a $S
  grep -iE "Traceback|Error|watchdog" gpurun_out/r2_22_$name.log | head -6 | cut -c1-300 | tee -a $S
}
run mixtral_pp2xtp2_fabric --model mixtral-8x7b --parallelism pp2xtp2
PETALS_B200_PPTP_FABRIC=0 run mixtral_pp2xtp2_rpc --model mixtral-8x7b --parallelism pp2xtp2 --skip-prefill
run tp4_default
