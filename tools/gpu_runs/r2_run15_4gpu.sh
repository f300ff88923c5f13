#!/bin/bash
# Round 2, call 15 (4 GPUs): the pipeline-of-tensor-parallel-groups layout (BASELINE config #4 is 4 x TP2 on Mixtral-8x7B) at 2 x TP2,
# and the tp4 line as the driver runs it (self-tests + pipeline record).
mkdir -p gpurun_out
S=gpurun_out/r2_15_summary.txt; : > $S
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
run() { name=$1; shift
  timeout 600 $TR --master-port 29742 bench.py --gpus 4 --steps 24 --warmup 4 "$@" > gpurun_out/r2_15_$name.log 2>&1; echo "$name exit=$?" | tee -a $S
  grep '^{' gpurun_out/r2_15_$name.log | python -c "import sys,json
for d in map(json.loads, sys.stdin):
    print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches','invalid')}, 'e2e', d.get('e2e',{}).get('value'), 'prefill', json.dumps(d.get('prefill'))[:400])
    print('   config:', json.dumps(d.get('config'))[:700])
    print('   selftests:', json.dumps(d.get('selftests'))[:300], 'pipeline:', json.dumps(d.get('pipeline'))[:900])" | tee -a $S
  grep -iE "Traceback|Error" gpurun_out/r2_15_$name.log | head -6 | cut -c1-300 | tee -a $S
}
run mixtral_pp2xtp2 --model mixtral-8x7b --parallelism pp2xtp2
run tp4_default
