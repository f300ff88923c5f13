#!/bin/bash
mkdir -p gpurun_out
python -c "from petals_b200.ops import native; native.lib(); print('native lib loads')" 2>&1 | tail -1 | tee gpurun_out/summary4.txt
timeout 420 python -m pytest tests/test_engine_gpu.py -q -m gpu > gpurun_out/test_engine.log 2>&1; echo "engine tests exit=$?" | tee -a gpurun_out/summary4.txt
tail -15 gpurun_out/test_engine.log | cut -c1-300
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?" | tee -a gpurun_out/summary4.txt
tail -5 gpurun_out/smoke.log | cut -c1-300
timeout 600 python bench.py --model llama-3-8b --steps 32 --warmup 4 --prefill-steps 2 > gpurun_out/bench_8b.log 2>&1; echo "bench8b exit=$?" | tee -a gpurun_out/summary4.txt
tail -2 gpurun_out/bench_8b.log | cut -c1-2000
timeout 900 python bench.py --steps 32 --warmup 4 --prefill-steps 1 > gpurun_out/bench_70b.log 2>&1; echo "bench70b exit=$?" | tee -a gpurun_out/summary4.txt
tail -2 gpurun_out/bench_70b.log | cut -c1-2000
for k in gemm gemv attn; do
  pat=gemm_tcgen05; [ $k = gemv ] && pat=linear_decode; [ $k = attn ] && pat=attn_fwd
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$pat -s 2 -c 1 -f -o gpurun_out/prof_$k python tools/profile_kernels.py $k > gpurun_out/ncu_$k.log 2>&1
  echo "ncu $k exit=$?" | tee -a gpurun_out/summary4.txt
done
ls -la gpurun_out/*.ncu-rep
