#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/summary9.txt; : > $S
run() {  # name, env...
  name=$1; shift
  for model in llama-3-70b; do
    env "$@" timeout 300 python bench.py --model $model --steps 64 --warmup 4 --skip-prefill --skip-fp8 > gpurun_out/b9_${name}_${model}.log 2>&1
    line=$(grep -E "^\{" gpurun_out/b9_${name}_${model}.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['clocks']['sm_mhz'], d['clocks']['reasons'])" 2>&1 | tail -1)
    echo "$name $model $line" | tee -a $S
  done
}
run m0   PETALS_B200_PDL_MASK=0
run m1   PETALS_B200_PDL_MASK=1
run m16  PETALS_B200_PDL_MASK=16
run m17  PETALS_B200_PDL_MASK=17
run m2   PETALS_B200_PDL_MASK=2
run m4   PETALS_B200_PDL_MASK=4
run m8   PETALS_B200_PDL_MASK=8
run m14  PETALS_B200_PDL_MASK=14
run m15  PETALS_B200_PDL_MASK=15
run m31  PETALS_B200_PDL_MASK=31
