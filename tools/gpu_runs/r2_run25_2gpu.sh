#!/bin/bash
# Round 2, call 25 (2 GPUs): the final tree in the driver's N = 2 job (self-tests incl. training over the fabric, tp2, pipeline record) and the
# multi-GPU pytest file.
mkdir -p gpurun_out
S=gpurun_out/r2_25_summary.txt; : > $S
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29742 bench.py --gpus 2 --steps 24 --warmup 4 > gpurun_out/r2_25_tp2_default.log 2>&1; echo "tp2_default exit=$?" | tee -a $S
grep '^{' gpurun_out/r2_25_tp2_default.log | python -c "import sys,json
for d in map(json.loads, sys.stdin):
    print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches','invalid')}, 'e2e', d.get('e2e',{}).get('value'), 'prefill', json.dumps(d.get('prefill'))[:300])
    print('   selftests:', json.dumps(d.get('selftests'))[:1200])
    print('   pipeline:', json.dumps(d.get('pipeline'))[:1500])" | tee -a $S
grep -iE "Traceback|Error|watchdog" gpurun_out/r2_25_tp2_default.log | head -5 | cut -c1-300 | tee -a $S
timeout 600 python -m pytest tests/test_multi_gpu.py -q --timeout=400 > gpurun_out/r2_25_pytest.log 2>&1; echo "multi-gpu pytest exit=$?" | tee -a $S
tail -3 gpurun_out/r2_25_pytest.log | cut -c1-300 | tee -a $S
