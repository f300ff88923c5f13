#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/summary20.txt; : > $S
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x > gpurun_out/t20_k.log 2>&1; echo "kernel tests exit=$?" | tee -a $S
tail -3 gpurun_out/t20_k.log | cut -c1-400 | tee -a $S
run() { name=$1; shift; model=$1; shift; extra=$1; shift
  env "$@" timeout 300 python bench.py --model $model --steps 64 --warmup 4 --skip-prefill --skip-fp8 $extra > gpurun_out/b20_${name}_${model}.log 2>&1
  echo "$name $model $extra $(grep -E '^\{' gpurun_out/b20_${name}_${model}.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['gpu_launches'], d['clocks']['reasons'])" 2>&1 | tail -1)" | tee -a $S
}
run nochain llama-3-70b "" PETALS_B200_CHAIN=0
run chain_pf16 llama-3-70b "" PETALS_B200_CHAIN=1
run chain_pf0 llama-3-70b "" PETALS_B200_CHAIN=1 PETALS_B200_CHAIN_PF=0
run chain_pf64 llama-3-70b "" PETALS_B200_CHAIN=1 PETALS_B200_CHAIN_PF=64
run nochain_tp8emu llama-3-70b "--tp-emulate 8" PETALS_B200_CHAIN=0
run chain_tp8emu llama-3-70b "--tp-emulate 8" PETALS_B200_CHAIN=1
run chain_pf0_tp8emu llama-3-70b "--tp-emulate 8" PETALS_B200_CHAIN=1 PETALS_B200_CHAIN_PF=0
for w in gemv_fp8 gemv_o; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:linear_decode -s 3 -c 1 -o gpurun_out/ncu_$w -f python tools/profile_kernels.py $w > gpurun_out/ncu_$w.log 2>&1; echo "ncu $w exit=$?" | tee -a $S
done
timeout 600 python bench.py --steps 16 --warmup 3 > gpurun_out/b20_full_70b.log 2>&1; echo "bench full exit=$?" | tee -a $S
grep -E "^\{" gpurun_out/b20_full_70b.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['prefill'], d['fp8_weights'])" | tee -a $S
