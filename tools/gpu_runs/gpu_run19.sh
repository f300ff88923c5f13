#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/summary19.txt; : > $S
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "chain or fp8 or linear_decode" > gpurun_out/t19_k.log 2>&1; echo "kernel tests exit=$?" | tee -a $S
tail -5 gpurun_out/t19_k.log | cut -c1-400 | tee -a $S
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_fp8_gpu.py -q -m gpu -x > gpurun_out/t19_engine.log 2>&1; echo "engine exit=$?" | tee -a $S
tail -5 gpurun_out/t19_engine.log | cut -c1-400 | tee -a $S
run() { name=$1; shift; model=$1; shift; extra=$1; shift
  env "$@" timeout 300 python bench.py --model $model --steps 64 --warmup 4 --skip-prefill --skip-fp8 $extra > gpurun_out/b19_${name}_${model}.log 2>&1
  echo "$name $model $extra $(grep -E '^\{' gpurun_out/b19_${name}_${model}.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['gpu_launches'], d['clocks']['reasons'])" 2>&1 | tail -1)" | tee -a $S
}
run chain llama-3-70b "" PETALS_B200_CHAIN=1
run nochain llama-3-70b "" PETALS_B200_CHAIN=0
run chain llama-3-8b "" PETALS_B200_CHAIN=1
run nochain llama-3-8b "" PETALS_B200_CHAIN=0
run chain_tp8emu llama-3-70b "--tp-emulate 8" PETALS_B200_CHAIN=1
run nochain_tp8emu llama-3-70b "--tp-emulate 8" PETALS_B200_CHAIN=0
run chain2 llama-3-70b "" PETALS_B200_CHAIN=1
timeout 300 python bench.py --steps 32 --warmup 4 --skip-prefill > gpurun_out/b19_fp8_70b.log 2>&1; echo "fp8 appendix: $(grep -E '^\{' gpurun_out/b19_fp8_70b.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['fp8_weights'])")" | tee -a $S
python tools/kernel_bench.py --fp8 > gpurun_out/kbench_fp8.log 2>&1; echo "kernel_bench fp8 exit=$?" | tee -a $S
tail -12 gpurun_out/kbench_fp8.log | cut -c1-200 | tee -a $S
