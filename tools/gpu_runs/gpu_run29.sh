#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/summary29.txt; : > $S
timeout 400 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "linear_decode or chain" > gpurun_out/t29_k.log 2>&1; echo "gemv tests exit=$?" | tee -a $S
tail -4 gpurun_out/t29_k.log | cut -c1-400 | tee -a $S
run() { name=$1; shift; extra=$1; shift
  env "$@" timeout 300 python bench.py --model llama-3-70b --steps 64 --warmup 4 --skip-prefill --skip-fp8 $extra > gpurun_out/b29_${name}.log 2>&1
  echo "$name $extra $(grep -E '^\{' gpurun_out/b29_${name}.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['clocks']['reasons'])" 2>&1 | tail -1)" | tee -a $S
}
run splitk_tp8emu "--tp-emulate 8" PETALS_B200_GEMV_SPLITK=1
run nosplit_tp8emu "--tp-emulate 8" PETALS_B200_GEMV_SPLITK=0
run splitk_tp4emu "--tp-emulate 4" PETALS_B200_GEMV_SPLITK=1
run nosplit_tp4emu "--tp-emulate 4" PETALS_B200_GEMV_SPLITK=0
timeout 300 python -m pytest tests/test_engine_gpu.py -q -m gpu -x > gpurun_out/t29_e.log 2>&1; echo "engine tests exit=$?" | tee -a $S
