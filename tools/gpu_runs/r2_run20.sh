#!/bin/bash
# Round 2, call 20 (1 GPU): FP8 GEMM with the scale-factor copies software-pipelined one K block ahead (A/B), MN-major B on the pair
# kernel (numerics + training benchmark), bf16 pair kernel with the new default raster group.
mkdir -p gpurun_out
S=gpurun_out/r2_20_summary.txt; : > $S
for pipe in 1 0; do
  PETALS_B200_FP8_SFPIPE=$pipe timeout 300 python -m pytest tests/test_fp8_gpu.py -q --timeout=120 -x > gpurun_out/r2_20_fp8_tests_pipe$pipe.log 2>&1; echo "fp8 tests SFPIPE=$pipe exit=$?" | tee -a $S
  tail -2 gpurun_out/r2_20_fp8_tests_pipe$pipe.log | cut -c1-200 | tee -a $S
  PETALS_B200_FP8_SFPIPE=$pipe timeout 300 python tools/kernel_bench.py --only gemm_fp8 > gpurun_out/r2_20_kb_fp8_pipe$pipe.log 2>&1; echo "kernel bench SFPIPE=$pipe exit=$?" | tee -a $S
  grep "gemm_mxfp8" gpurun_out/r2_20_kb_fp8_pipe$pipe.log | cut -c1-110 | tee -a $S
done
PETALS_B200_RUN_UNVALIDATED=1 timeout 300 python -m pytest tests/test_kernels_gpu.py -q --timeout=120 -x -k "2cta" > gpurun_out/r2_20_mn_test.log 2>&1; echo "2cta incl. MN-major exit=$?" | tee -a $S
tail -3 gpurun_out/r2_20_mn_test.log | cut -c1-300 | tee -a $S
if grep -q "passed" gpurun_out/r2_20_mn_test.log && ! grep -q "failed" gpurun_out/r2_20_mn_test.log; then
  for mn in 0 1; do
    PETALS_B200_GEMM_2CTA_MN=$mn timeout 600 python benchmarks/benchmark_training.py --model llama-3-8b --n_steps 8 --warmup_steps 3 --batch_size 8 --seq_len 128 > gpurun_out/r2_20_training_mn$mn.log 2>&1
    echo "training bench MN=$mn exit=$?" | tee -a $S; grep "Final result" gpurun_out/r2_20_training_mn$mn.log | cut -c1-200 | tee -a $S
  done
  PETALS_B200_GEMM_2CTA_MN=1 timeout 600 python -m pytest tests/test_engine_gpu.py -q --timeout=300 -k "backward" > gpurun_out/r2_20_bwd_tests_mn.log 2>&1; echo "engine backward tests with MN pair dgrad exit=$?" | tee -a $S
  tail -2 gpurun_out/r2_20_bwd_tests_mn.log | cut -c1-200 | tee -a $S
fi
timeout 300 python tools/kernel_bench.py --only gemm_2cta > gpurun_out/r2_20_kb_2cta.log 2>&1; echo "2cta kernel bench (group 8 default) exit=$?" | tee -a $S
grep "gemm2" gpurun_out/r2_20_kb_2cta.log | cut -c1-200 | tee -a $S
