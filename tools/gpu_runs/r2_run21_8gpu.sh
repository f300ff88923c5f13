#!/bin/bash
# Round 2, call 21 (8 GPUs): the final tree exactly as the driver runs it at N = 8 (self-tests, tp8 decode + sequence-parallel prefill on the
# pair GEMM, pipeline record).
mkdir -p gpurun_out
S=gpurun_out/r2_21_summary.txt; : > $S
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29742 bench.py --gpus 8 --steps 24 --warmup 4 > gpurun_out/r2_21_tp8_default.log 2>&1; echo "tp8_default exit=$?" | tee -a $S
grep '^{' gpurun_out/r2_21_tp8_default.log | python -c "import sys,json
for d in map(json.loads, sys.stdin):
    print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches','invalid')}, 'e2e', d.get('e2e',{}).get('value'), 'hbm', d.get('roofline',{}).get('frac_of_measured_hbm'), 'prefill', json.dumps(d.get('prefill'))[:400])
    print('   selftests:', json.dumps(d.get('selftests'))[:1200])
    print('   pipeline:', json.dumps(d.get('pipeline'))[:1800])" | tee -a $S
grep -iE "Traceback|Error|watchdog" gpurun_out/r2_21_tp8_default.log | head -5 | cut -c1-300 | tee -a $S
