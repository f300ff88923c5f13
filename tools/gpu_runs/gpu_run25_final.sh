#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/summary25.txt; : > $S
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/t25_all.log 2>&1; echo "pytest -m gpu exit=$?" | tee -a $S
tail -4 gpurun_out/t25_all.log | cut -c1-400 | tee -a $S
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke25.log 2>&1; echo "smoke exit=$?" | tee -a $S
timeout 600 python bench.py > gpurun_out/b25_default.log 2>&1; echo "bench default exit=$?" | tee -a $S
grep -E "^\{" gpurun_out/b25_default.log | tail -1 | cut -c1-3000 | tee -a $S
timeout 400 python bench.py --model mixtral-8x7b --steps 32 --warmup 4 --skip-prefill --skip-fp8 > gpurun_out/b25_mixtral.log 2>&1; echo "bench mixtral-8x7b exit=$?" | tee -a $S
grep -E "^\{" gpurun_out/b25_mixtral.log | tail -1 | cut -c1-1500 | tee -a $S
tail -3 gpurun_out/b25_mixtral.log | cut -c1-300
