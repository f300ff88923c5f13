#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/summary12.txt; : > $S
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "attention" > gpurun_out/t12_attn.log 2>&1; echo "attention tests exit=$?" | tee -a $S
tail -15 gpurun_out/t12_attn.log | cut -c1-400
timeout 120 python tools/profile_kernels.py attn_time > gpurun_out/attn_time.log 2>&1; echo "attn_time exit=$?" | tee -a $S
tail -3 gpurun_out/attn_time.log | cut -c1-600 | tee -a $S
timeout 600 python -m pytest tests/test_engine_gpu.py -q -m gpu -x > gpurun_out/t12_engine.log 2>&1; echo "engine tests exit=$?" | tee -a $S
tail -5 gpurun_out/t12_engine.log | cut -c1-400
timeout 300 python bench.py --steps 64 --warmup 4 --skip-fp8 > gpurun_out/b12_70b.log 2>&1; echo "bench 70b exit=$?" | tee -a $S
grep -E "^\{" gpurun_out/b12_70b.log | tail -1 | cut -c1-2500 | tee -a $S
for R in 8; do
  timeout 200 python bench.py --tp-emulate $R --steps 64 --warmup 4 --skip-prefill --skip-fp8 > gpurun_out/emul12_tp$R.log 2>&1; echo "emulate tp$R (auto pdl): $(grep -E '^\{' gpurun_out/emul12_tp$R.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'])")" | tee -a $S
  PETALS_B200_PDL_MASK=14 timeout 200 python bench.py --tp-emulate $R --steps 64 --warmup 4 --skip-prefill --skip-fp8 > gpurun_out/emul12_tp${R}_m14.log 2>&1; echo "emulate tp$R (mask 14): $(grep -E '^\{' gpurun_out/emul12_tp${R}_m14.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'])")" | tee -a $S
done
timeout 200 python bench.py --model llama-3-8b --steps 64 --warmup 4 --skip-prefill --skip-fp8 > gpurun_out/b12_8b.log 2>&1; echo "8b: $(grep -E '^\{' gpurun_out/b12_8b.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'])")" | tee -a $S
