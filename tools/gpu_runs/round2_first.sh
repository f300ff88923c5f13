#!/bin/bash
# First GPU call of the next round (8 GPUs): validate what was written after round 1's GPU budget ran out, then measure the
# L2-prefetch experiment against the default at tp8 where kernel boundaries dominate.
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 1500 -- bash tools/gpu_runs/round2_first.sh
mkdir -p gpurun_out
S=gpurun_out/summary_r2_first.txt; : > $S
export PETALS_B200_RUN_UNVALIDATED=1
timeout 300 python -m pytest tests/test_l2_prefetch_gpu.py -q -x > gpurun_out/r2_l2pf_tests.log 2>&1; echo "l2 prefetch tests exit=$?" | tee -a $S
tail -3 gpurun_out/r2_l2pf_tests.log | cut -c1-300 | tee -a $S
timeout 500 python -m pytest tests/test_multi_gpu.py -q -x -k backward > gpurun_out/r2_tp_backward.log 2>&1; echo "tp backward exit=$?" | tee -a $S
tail -3 gpurun_out/r2_tp_backward.log | cut -c1-300 | tee -a $S
unset PETALS_B200_RUN_UNVALIDATED
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py \
      --gpus 8 --steps 48 --warmup 6 --skip-prefill --skip-fp8 > gpurun_out/r2_tp8_$name.log 2>&1
  echo "tp8 $name exit=$?" | tee -a $S
  grep -E "^\{" gpurun_out/r2_tp8_$name.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['unit'], d['ms_per_step'], 'ms', 'e2e', d['e2e']['value'])" 2>&1 | tee -a $S
}
run default PETALS_B200_L2_PREFETCH=0
run l2pf_a1_c8 PETALS_B200_L2_PREFETCH=1 PETALS_B200_L2_PREFETCH_AHEAD=1 PETALS_B200_L2_PREFETCH_CTAS=8
run l2pf_a2_c8 PETALS_B200_L2_PREFETCH=1 PETALS_B200_L2_PREFETCH_AHEAD=2 PETALS_B200_L2_PREFETCH_CTAS=8
run l2pf_a1_c16 PETALS_B200_L2_PREFETCH=1 PETALS_B200_L2_PREFETCH_AHEAD=1 PETALS_B200_L2_PREFETCH_CTAS=16
# kernels are 3-19 us at tp8: programmatic dependent launch on the GEMVs (slower at 1 GPU, profiles/r1_pdl_sweep.txt) may pay here
run pdl15 PETALS_B200_PDL_MASK=15
run pdl15_l2pf PETALS_B200_PDL_MASK=15 PETALS_B200_L2_PREFETCH=1
