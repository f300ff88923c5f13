#!/bin/bash
# Round 2, call 11 (2 GPUs): TP self-tests with the span kernel after the flag-epoch fix, pipeline self-test with training over the fabric
# (forward + gradient hops), multi-GPU pytest, tp2 bench with the pipeline sub-record.
mkdir -p gpurun_out
S=gpurun_out/r2_11_summary.txt; : > $S
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
sel() { name=$1; shift
  env "$@" timeout 300 $TR --master-port 29741 tools/tp_selftest.py > gpurun_out/r2_11_sel_$name.log 2>&1; echo "tp_selftest $name exit=$?" | tee -a $S
  grep '^{' gpurun_out/r2_11_sel_$name.log | tail -1 | cut -c1-400 | tee -a $S
  grep -iE "Error|stuck" gpurun_out/r2_11_sel_$name.log | head -4 | cut -c1-250 | tee -a $S
}
sel span_st PETALS_B200_SPAN_NVLS=st
sel span_reduce PETALS_B200_SPAN_NVLS=reduce
sel span_unicast PETALS_B200_SPAN_NVLS=0
sel ipc_heap PETALS_B200_SYMM_MEM=0
timeout 400 $TR --master-port 29743 tools/pp_selftest.py > gpurun_out/r2_11_pp_selftest.log 2>&1; echo "pp_selftest exit=$?" | tee -a $S
grep '^{' gpurun_out/r2_11_pp_selftest.log | tail -1 | cut -c1-900 | tee -a $S
grep -iE "Error|Traceback" gpurun_out/r2_11_pp_selftest.log | head -6 | cut -c1-300 | tee -a $S
timeout 900 python -m pytest tests/test_multi_gpu.py -q --timeout=400 > gpurun_out/r2_11_pytest.log 2>&1; echo "multi-gpu pytest exit=$?" | tee -a $S
tail -5 gpurun_out/r2_11_pytest.log | cut -c1-300 | tee -a $S
run() { name=$1; shift
  timeout 900 $TR --master-port 29742 bench.py --gpus 2 --steps 24 --warmup 4 "$@" > gpurun_out/r2_11_$name.log 2>&1; echo "$name exit=$?" | tee -a $S
  grep '^{' gpurun_out/r2_11_$name.log | python -c "import sys,json
for d in map(json.loads, sys.stdin):
    print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d.get('e2e',{}).get('value'), 'hbm', d.get('roofline',{}).get('frac_of_measured_hbm'), 'prefill', (d.get('prefill') or {}).get('tokens_per_s'))
    print('   pipeline:', json.dumps(d.get('pipeline'))[:1500])" | tee -a $S
  grep -iE "Traceback|Error" gpurun_out/r2_11_$name.log | head -3 | cut -c1-300 | tee -a $S
}
run tp2_default
