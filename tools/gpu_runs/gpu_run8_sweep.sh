#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/summary8.txt; : > $S
run() {  # name, env...
  name=$1; shift
  for model in llama-3-70b llama-3-8b; do
    env "$@" timeout 300 python bench.py --model $model --steps 64 --warmup 4 --skip-prefill --skip-fp8 > gpurun_out/b8_${name}_${model}.log 2>&1
    line=$(grep -E "^\{" gpurun_out/b8_${name}_${model}.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['clocks']['sm_mhz'], d['clocks']['reasons'])" 2>&1 | tail -1)
    echo "$name $model $line" | tee -a $S
  done
}
run base      PETALS_B200_PDL=0 PETALS_B200_PF_LINES=0
run pf16      PETALS_B200_PDL=0 PETALS_B200_PF_LINES=16
run pdl       PETALS_B200_PDL=1 PETALS_B200_PF_LINES=0
run pdl_pf8   PETALS_B200_PDL=1 PETALS_B200_PF_LINES=8
run pdl_pf16  PETALS_B200_PDL=1 PETALS_B200_PF_LINES=16
run pdl_pf64  PETALS_B200_PDL=1 PETALS_B200_PF_LINES=64
run late      PETALS_B200_PDL=1 PETALS_B200_PF_LINES=0 PETALS_B200_PDL_LATE=1
run late_pf16 PETALS_B200_PDL=1 PETALS_B200_PF_LINES=16 PETALS_B200_PDL_LATE=1
run late_pf64 PETALS_B200_PDL=1 PETALS_B200_PF_LINES=64 PETALS_B200_PDL_LATE=1
run base2     PETALS_B200_PDL=0 PETALS_B200_PF_LINES=0
