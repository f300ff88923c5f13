#!/bin/bash
# Round 2, call 17 (8 GPUs): Mixtral-8x7B as 4 pipeline stages x tensor-parallel pairs (BASELINE config #4 layout, bf16), Mixtral tp8,
# and the pipelined prompt ingestion with 128- vs 256-token chunks.
mkdir -p gpurun_out
S=gpurun_out/r2_17_summary.txt; : > $S
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
run() { name=$1; shift
  timeout 420 $TR --master-port 29742 bench.py --gpus 8 --steps 24 --warmup 4 "$@" > gpurun_out/r2_17_$name.log 2>&1; echo "$name exit=$?" | tee -a $S
  grep '^{' gpurun_out/r2_17_$name.log | python -c "import sys,json
for d in map(json.loads, sys.stdin):
    print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches','invalid')}, 'e2e', d.get('e2e',{}).get('value'), 'prefill', json.dumps(d.get('prefill'))[:500])
    print('   config:', json.dumps(d.get('config'))[:500])
    print('   selftests:', json.dumps(d.get('selftests'))[:400])" | tee -a $S
  grep -iE "Traceback|Error|watchdog" gpurun_out/r2_17_$name.log | head -6 | cut -c1-300 | tee -a $S
}
run mixtral_pp4xtp2 --model mixtral-8x7b --parallelism pp4xtp2
run mixtral_tp8 --model mixtral-8x7b --skip-pipeline
run pp8_chunk128 --parallelism pp8 --pp-chunk-tokens 128 --skip-selftests
run pp8_chunk256 --parallelism pp8 --pp-chunk-tokens 256 --skip-selftests
