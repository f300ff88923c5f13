#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/summary15.txt; : > $S
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x > gpurun_out/t15_kernels.log 2>&1; echo "kernels exit=$?" | tee -a $S
tail -5 gpurun_out/t15_kernels.log | cut -c1-400
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_fp8_gpu.py -q -m gpu -x > gpurun_out/t15_engine.log 2>&1; echo "engine exit=$?" | tee -a $S
tail -5 gpurun_out/t15_engine.log | cut -c1-400
run() { name=$1; shift; model=$1; shift
  env "$@" timeout 300 python bench.py --model $model --steps 64 --warmup 4 --skip-prefill --skip-fp8 > gpurun_out/b15_${name}_${model}.log 2>&1
  echo "$name $model $(grep -E '^\{' gpurun_out/b15_${name}_${model}.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['gpu_launches'], d['clocks']['reasons'])" 2>&1 | tail -1)" | tee -a $S
}
run fused llama-3-70b PETALS_B200_FUSE_ROPE=1
run unfused llama-3-70b PETALS_B200_FUSE_ROPE=0
run fused llama-3-8b PETALS_B200_FUSE_ROPE=1
run unfused llama-3-8b PETALS_B200_FUSE_ROPE=0
env PETALS_B200_FUSE_ROPE=1 timeout 200 python bench.py --tp-emulate 8 --steps 64 --warmup 4 --skip-prefill --skip-fp8 > gpurun_out/emul15_f.log 2>&1; echo "emulate tp8 fused: $(grep -E '^\{' gpurun_out/emul15_f.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'])")" | tee -a $S
env PETALS_B200_FUSE_ROPE=0 timeout 200 python bench.py --tp-emulate 8 --steps 64 --warmup 4 --skip-prefill --skip-fp8 > gpurun_out/emul15_u.log 2>&1; echo "emulate tp8 unfused: $(grep -E '^\{' gpurun_out/emul15_u.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'])")" | tee -a $S
