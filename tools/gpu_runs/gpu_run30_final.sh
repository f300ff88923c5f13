#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/summary30.txt; : > $S
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/t30_all.log 2>&1; echo "pytest -m gpu exit=$?" | tee -a $S
tail -4 gpurun_out/t30_all.log | cut -c1-400 | tee -a $S
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke30.log 2>&1; echo "smoke exit=$?" | tee -a $S
