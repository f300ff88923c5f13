#!/bin/bash
# Round 2, call 19 (1 GPU): the final-tree checks the driver will repeat — full GPU suite, smoke(), bench.py at N = 1 (2-CTA GEMM on by
# default now) — then the ncu captures of the round's new kernels.
mkdir -p gpurun_out
S=gpurun_out/r2_19_summary.txt; : > $S
PETALS_B200_RUN_UNVALIDATED=1 timeout 300 python -m pytest tests/test_kernels_gpu.py -q --timeout=120 -x -k "2cta" > gpurun_out/r2_19_mn_test.log 2>&1; echo "2cta incl. MN-major exit=$?" | tee -a $S; tail -3 gpurun_out/r2_19_mn_test.log | cut -c1-300 | tee -a $S
timeout 1500 python -m pytest tests -m gpu -q --timeout=300 -x > gpurun_out/r2_19_gpu_suite.log 2>&1; echo "gpu suite exit=$?" | tee -a $S
tail -6 gpurun_out/r2_19_gpu_suite.log | cut -c1-300 | tee -a $S
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_19_smoke.log 2>&1; echo "smoke exit=$?" | tee -a $S
tail -2 gpurun_out/r2_19_smoke.log | cut -c1-300 | tee -a $S
timeout 900 python bench.py --steps 24 --warmup 4 > gpurun_out/r2_19_bench.log 2>&1; echo "bench exit=$?" | tee -a $S
grep '^{' gpurun_out/r2_19_bench.log | python -c "import sys,json; [print({k:d[k] for k in ('value','ms_per_step','gpu_launches','vs_baseline') if k in d}, 'e2e', d.get('e2e'), 'clocks', d.get('clocks'), 'prefill', json.dumps(d.get('prefill'))[:300], 'fp8', json.dumps(d.get('fp8_weights'))[:700]) for d in map(json.loads, sys.stdin)]" | tee -a $S
timeout 600 python benchmarks/benchmark_training.py --model llama-3-8b --n_steps 8 --warmup_steps 3 --batch_size 8 --seq_len 128 > gpurun_out/r2_19_training.log 2>&1; echo "training bench exit=$?" | tee -a $S
grep -iE "tokens/s|tok/s|forward|backward" gpurun_out/r2_19_training.log | tail -6 | cut -c1-300 | tee -a $S
bash tools/gpu_runs/r2_run18_ncu.sh > /dev/null 2>&1; cat gpurun_out/r2_18_summary.txt | cut -c1-200 | tee -a $S
PETALS_B200_FP8_2CTA=1 timeout 600 python tools/kernel_bench.py --only gemm_fp8 > gpurun_out/r2_19_kernel_bench_fp8_2cta.log 2>&1; echo "fp8 2cta (grouped raster) kernel bench exit=$?" | tee -a $S
grep "gemm_mxfp8" gpurun_out/r2_19_kernel_bench_fp8_2cta.log | cut -c1-300 | tee -a $S
for g in 2 8 16; do
  PETALS_B200_GEMM_GROUP_M=$g timeout 300 python tools/kernel_bench.py --only gemm_2cta > gpurun_out/r2_19_kb_group$g.log 2>&1; echo "2cta group_m=$g exit=$?" | tee -a $S
  grep "gemm2" gpurun_out/r2_19_kb_group$g.log | cut -c1-130 | tee -a $S
done
PETALS_B200_GEMM_2CTA_MN=1 timeout 600 python benchmarks/benchmark_training.py --model llama-3-8b --n_steps 8 --warmup_steps 3 --batch_size 8 --seq_len 128 > gpurun_out/r2_19_training_mn.log 2>&1; echo "training bench (MN dgrad on the pair kernel) exit=$?" | tee -a $S
grep -iE "tokens/s|tok/s|forward|backward" gpurun_out/r2_19_training_mn.log | tail -6 | cut -c1-300 | tee -a $S
