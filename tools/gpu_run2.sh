#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_engine_gpu.py -q -m gpu > gpurun_out/test_engine.log 2>&1; echo "engine tests exit=$?" | tee gpurun_out/summary2.txt
tail -25 gpurun_out/test_engine.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?" | tee -a gpurun_out/summary2.txt
tail -5 gpurun_out/smoke.log
timeout 900 python bench.py --model llama-3-8b --steps 32 --warmup 4 --prefill-steps 2 > gpurun_out/bench_8b.log 2>&1; echo "bench8b exit=$?" | tee -a gpurun_out/summary2.txt
tail -3 gpurun_out/bench_8b.log
timeout 1200 python bench.py --steps 32 --warmup 4 --prefill-steps 1 > gpurun_out/bench_70b.log 2>&1; echo "bench70b exit=$?" | tee -a gpurun_out/summary2.txt
tail -3 gpurun_out/bench_70b.log
