#!/bin/bash
# Round 2, call 13 (8 GPUs): exactly what the driver runs at N=8 (self-tests + tp8 decode/prefill + pp8 pipeline record), then the
# tensor-parallel decode alone without the span kernel and with unicast pushes instead of multimem.st.
mkdir -p gpurun_out
S=gpurun_out/r2_13_summary.txt; : > $S
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
run() { name=$1; shift
  timeout 900 $TR --master-port 29742 bench.py --gpus 8 --steps 24 --warmup 4 "$@" > gpurun_out/r2_13_$name.log 2>&1; echo "$name exit=$?" | tee -a $S
  grep '^{' gpurun_out/r2_13_$name.log | python -c "import sys,json
for d in map(json.loads, sys.stdin):
    print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches','invalid')}, 'e2e', d.get('e2e',{}).get('value'), 'hbm', d.get('roofline',{}).get('frac_of_measured_hbm'), 'prefill', (d.get('prefill') or {}).get('tokens_per_s'))
    print('   selftests:', json.dumps(d.get('selftests'))[:1500])
    print('   pipeline:', json.dumps(d.get('pipeline'))[:1800])" | tee -a $S
  grep -iE "Traceback|Error|stuck" gpurun_out/r2_13_$name.log | head -5 | cut -c1-300 | tee -a $S
}
run tp8_default
PETALS_B200_SPAN_KERNEL=0 run tp8_nospan --skip-pipeline --skip-prefill --skip-selftests
PETALS_B200_SPAN_NVLS=0 run tp8_span_unicast --skip-pipeline --skip-prefill --skip-selftests
nvidia-smi topo -m > gpurun_out/r2_13_topo.txt 2>&1
