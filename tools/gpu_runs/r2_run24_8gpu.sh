#!/bin/bash
# Round 2, call 24 (8 GPUs): Mixtral-8x7B as 4 pipeline stages x TP 2 (BASELINE config #4 layout) with the leaders' NVLink fabric.
mkdir -p gpurun_out
S=gpurun_out/r2_24_summary.txt; : > $S
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 420 $TR --master-port 29742 bench.py --gpus 8 --steps 24 --warmup 4 --model mixtral-8x7b --parallelism pp4xtp2 > gpurun_out/r2_24_mixtral_pp4xtp2.log 2>&1; echo "mixtral_pp4xtp2 (fabric) exit=$?" | tee -a $S
grep '^{' gpurun_out/r2_24_mixtral_pp4xtp2.log | python -c "import sys,json
for d in map(json.loads, sys.stdin):
    print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches','invalid')}, 'e2e', d.get('e2e',{}).get('value'), 'prefill', json.dumps(d.get('prefill'))[:400])
    print('   config:', json.dumps(d.get('config'))[:600])" | tee -a $S
grep -iE "Traceback|Error|watchdog" gpurun_out/r2_24_mixtral_pp4xtp2.log | head -6 | cut -c1-300 | tee -a $S
