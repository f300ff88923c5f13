#!/bin/bash
# Round 2, call 18 (1 GPU): ncu captures (one kernel each, --set full) of the round's new hot kernels, summarised into gpurun_out/*.txt:
# the span kernel at tp8 shapes (2 layers), the block-scaled FP8 GEMM, the 1-CTA and 2-CTA bf16 GEMM. Reports stay in gpurun_out (scratch).
mkdir -p gpurun_out
S=gpurun_out/r2_18_summary.txt; : > $S
cap() { name=$1; regex=$2; skip=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -f -o gpurun_out/r2_18_$name "$@" > gpurun_out/r2_18_$name.log 2>&1
  echo "$name ncu exit=$?" | tee -a $S
  python tools/summarize_ncu.py gpurun_out/r2_18_$name.ncu-rep gpurun_out/r2_18_ncu_$name.txt 2>&1 | head -22 | cut -c1-200 | tee -a $S
}
cap span_tp8 decode_span_kernel 2 python tools/span_probe.py --shape 70b-tp8 --layers 2
cap gemm_mxfp8 gemm_mxfp8_kernel 2 python tools/profile_kernels.py gemm_mxfp8
cap gemm_2cta gemm_2cta_kernel 2 python tools/profile_kernels.py gemm_2cta
cap gemm_mxfp8_2cta gemm_mxfp8_2cta_kernel 2 python tools/profile_kernels.py gemm_mxfp8_2cta
rm -f gpurun_out/r2_18_span_tp8.ncu-rep gpurun_out/r2_18_gemm_2cta.ncu-rep gpurun_out/r2_18_gemm_mxfp8_2cta.ncu-rep   # keep one report (size limit of the merge-back)
