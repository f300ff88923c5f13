#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/summary24.txt; : > $S
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "pipelined or linear_decode_plain or chain or loopback" > gpurun_out/t24_k.log 2>&1; echo "pipe tests exit=$?" | tee -a $S
tail -3 gpurun_out/t24_k.log | cut -c1-400 | tee -a $S
run() { name=$1; shift; model=$1; shift; extra=$1; shift
  env "$@" timeout 300 python bench.py --model $model --steps 64 --warmup 4 --skip-prefill --skip-fp8 $extra > gpurun_out/b24_${name}_${model}.log 2>&1
  echo "$name $model $extra $(grep -E '^\{' gpurun_out/b24_${name}_${model}.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['clocks']['reasons'])" 2>&1 | tail -1)" | tee -a $S
}
run base llama-3-70b "" PETALS_B200_GEMV_PIPE=0
run pipe llama-3-70b "" PETALS_B200_GEMV_PIPE=1
run base llama-3-8b "" PETALS_B200_GEMV_PIPE=0
run pipe llama-3-8b "" PETALS_B200_GEMV_PIPE=1
run base_tp8emu llama-3-70b "--tp-emulate 8" PETALS_B200_GEMV_PIPE=0
run pipe_tp8emu llama-3-70b "--tp-emulate 8" PETALS_B200_GEMV_PIPE=1
run base_tp4emu llama-3-70b "--tp-emulate 4" PETALS_B200_GEMV_PIPE=0
run pipe_tp4emu llama-3-70b "--tp-emulate 4" PETALS_B200_GEMV_PIPE=1
