#!/usr/bin/env python
"""Per-kernel SASS listings of the built library (evidence that the hot paths are the Blackwell ones).

For every selected kernel: `profiles/sass/<name>.sass` = its instruction stream (addresses + mnemonics + operands, encodings
stripped) and one line in `profiles/sass/INDEX.md` with the register count and the counts of the mnemonics that matter
(UTC*MMA = tcgen05.mma (UTCQMMA = block-scaled FP8), UTCCP = tcgen05.cp, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG/UBLKCP = TMA, UCGABAR = cluster
barrier, HMMA = mma.sync, LDGMC = multimem.ld_reduce (multimem.st lowers to a plain STG on the multicast address), SYNCS = mbarrier,
ACQBULK/PREEXIT = programmatic dependent launch). Runs on the CPU-only box: `python tools/sass_listing.py`."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "petals_b200", "_native", "obj")
OUT = os.path.join(ROOT, "profiles", "sass")
# object file -> substrings selecting ONE representative instantiation per kernel family
SELECT = {
    "gemm_tcgen05.o": ["gemm_tcgen05_kernelILi256ELb0ELb0", "gemm_tcgen05_kernelILi256ELb1ELb0"], "attention_tc.o": ["attn_fwd_tc_kernelILi128"], "attention.o": ["attn_fwd_kernelILi128"],
    "attention_bwd.o": ["attn_bwd_dq_kernelILi128", "attn_bwd_dkdv_kernelILi128"], "decode_span.o": ["decode_span_kernel"],
    "linear_decode.o": ["linear_decode_kernelILi1ELb1ELb1ELb0ELb0", "linear_decode_kernelILi1ELb0ELb1ELb1ELb0"],
    "linear_decode_fp8.o": ["linear_decode_fp8"], "linear_decode_mma.o": ["linear_decode_mma"], "moe.o": ["moe_gemv", "moe_router", "moe_plan"],
    "seq_parallel.o": ["norm_reduce_gather"], "train_kernels.o": ["rmsnorm_bwd_kernelILi4", "swiglu_bwd"], "gemm_mxfp8.o": ["gemm_mxfp8_kernelILb0", "gemm_mxfp8_kernelILb1", "quant_mxfp8"],
    "gemm_tcgen05_2cta.o": ["gemm_2cta_kernelILb0"], "ll_collectives.o": ["ll_reduce", "ll_push"],
    "elementwise.o": ["norm_kernel"], "rope_kv.o": ["rope_kv_kernel"], "ipc.o": ["push_rows"],
}
KEY = ["UTCHMMA", "UTCQMMA", "UTCOMMA", "UTCMXQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCCP", "UTCBAR", "UCGABAR_ARV", "HMMA", "MULTIMEM", "LDGMC", "SYNCS", "LDGSTS",
       "ACQBULK", "PREEXIT", "LDG", "STG", "LDS", "STS", "FFMA", "BAR"]


def main():
    os.makedirs(OUT, exist_ok=True)
    index = ["# Per-kernel SASS listings (sm_100a)", "", "| kernel | object | instructions | " + " | ".join(KEY) + " |", "|---|---|---|" + "---|" * len(KEY)]
    for obj, wanted in sorted(SELECT.items()):
        path = os.path.join(OBJ, obj)
        if not os.path.exists(path):
            continue
        text = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
        funcs = re.split(r"\n\s*Function : ", text)[1:]
        done = set()
        for f in funcs:
            name = f.split("\n", 1)[0].strip()
            hit = next((w for w in wanted if w in name and w not in done), None)
            if hit is None:
                continue
            done.add(hit)
            lines = []
            counts = collections.Counter()
            for ln in f.split("\n"):
                m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
                if not m:
                    continue
                ins = m.group(2).strip()
                lines.append(f"{m.group(1)}  {ins}")
                mn = re.sub(r"^@!?U?P\d+\s+", "", ins).split()[0].split(".")[0]
                counts[mn] += 1
            short = re.sub(r"[^A-Za-z0-9_]+", "_", hit)[:60]
            with open(os.path.join(OUT, f"{obj[:-2]}__{short}.sass"), "w") as fh:
                fh.write(f"// {name}\n// from petals_b200/csrc/{obj[:-2]}.cu, cuobjdump -sass (encodings stripped)\n" + "\n".join(lines) + "\n")
            index.append(f"| `{hit}` | {obj[:-2]}.cu | {len(lines)} | " + " | ".join(str(counts.get(k, 0) or "") for k in KEY) + " |")
    with open(os.path.join(OUT, "INDEX.md"), "w") as fh:
        fh.write("\n".join(index) + "\n")
    print("\n".join(index))


if __name__ == "__main__":
    sys.exit(main())
