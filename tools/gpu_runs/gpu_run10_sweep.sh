#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/summary10.txt; : > $S
run() {  # name, env...
  name=$1; shift
  for model in llama-3-70b; do
    env "$@" timeout 300 python bench.py --model $model --steps 64 --warmup 4 --skip-prefill --skip-fp8 > gpurun_out/b10_${name}_${model}.log 2>&1
    line=$(grep -E "^\{" gpurun_out/b10_${name}_${model}.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['clocks']['sm_mhz'], d['clocks']['reasons'])" 2>&1 | tail -1)
    echo "$name $model $line" | tee -a $S
  done
}
run m14      PETALS_B200_PDL_MASK=14
run m31_w0   PETALS_B200_PDL_MASK=31
run m31_w0_pf0 PETALS_B200_PDL_MASK=31 PETALS_B200_PF_LINES=0
run m31_wall PETALS_B200_PDL_MASK=31 PETALS_B200_PDL_WAIT_ALL=1
run m31_w0_late PETALS_B200_PDL_MASK=31 PETALS_B200_PDL_LATE=1
run m15_w0   PETALS_B200_PDL_MASK=15
run m30_w0   PETALS_B200_PDL_MASK=30
run m0       PETALS_B200_PDL_MASK=0
