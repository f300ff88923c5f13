#!/bin/bash
# Round 2, call 14 (2 GPUs): sparse-MoE blocks in the tensor-parallel engine (experts' FFN columns split over the pair, replicated router):
# numerics self-test on mixtral-tiny, Mixtral-8x7B tp2 bench (decode + sequence-parallel prefill), dense self-test again.
mkdir -p gpurun_out
S=gpurun_out/r2_14_summary.txt; : > $S
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
sel() { name=$1; shift
  env "$@" timeout 300 $TR --master-port 29741 tools/tp_selftest.py > gpurun_out/r2_14_sel_$name.log 2>&1; echo "tp_selftest $name exit=$?" | tee -a $S
  grep '^{' gpurun_out/r2_14_sel_$name.log | tail -1 | cut -c1-500 | tee -a $S
  grep -iE "Error|stuck" gpurun_out/r2_14_sel_$name.log | head -6 | cut -c1-300 | tee -a $S
}
sel mixtral TP_SELFTEST_MODEL=mixtral-tiny
sel mixtral_noll TP_SELFTEST_MODEL=mixtral-tiny PETALS_B200_TP_LL=0
sel llama TP_SELFTEST_MODEL=llama-tiny
run() { name=$1; shift
  timeout 900 $TR --master-port 29742 bench.py --gpus 2 --steps 24 --warmup 4 "$@" > gpurun_out/r2_14_$name.log 2>&1; echo "$name exit=$?" | tee -a $S
  grep '^{' gpurun_out/r2_14_$name.log | python -c "import sys,json
for d in map(json.loads, sys.stdin):
    print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches','invalid')}, 'e2e', d.get('e2e',{}).get('value'), 'hbm', d.get('roofline',{}).get('frac_of_measured_hbm'), 'prefill', json.dumps(d.get('prefill'))[:300])
    print('   selftests:', json.dumps(d.get('selftests'))[:600])" | tee -a $S
  grep -iE "Traceback|Error" gpurun_out/r2_14_$name.log | head -5 | cut -c1-300 | tee -a $S
}
run mixtral_tp2 --model mixtral-8x7b --skip-pipeline
