"""Device-timed micro-benchmarks of the hot kernels at Llama-3-70B / 8B shapes.

Timing hygiene (B200_PROFILING.md): CUDA events on the launching stream, >= 3 warm-ups, inputs rotated
through a pool larger than the 126 MB L2 so every timed launch streams from HBM.
Writes gpurun_out/kernel_bench.json and prints a table with roofline fractions against MEASURED_PEAKS.json.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from petals_b200.ops import functional as Fn  # noqa: E402
from petals_b200.utils.peaks import measured_peaks  # noqa: E402


_MN_DEFAULT = os.environ.get("PETALS_B200_GEMM_2CTA_MN", "1") != "0"


def time_fn(fns, iters=20, warmup=3):
    """fns: list of callables rotated round-robin (distinct buffers => cold L2). Returns ms per call."""
    for i in range(warmup):
        fns[i % len(fns)]()
    torch.cuda.synchronize()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for i in range(iters):
        fns[i % len(fns)]()
    end.record()
    torch.cuda.synchronize()
    return start.elapsed_time(end) / iters


def bench_gemv(results, peaks):
    shapes = [
        ("70b.qkv", 10240, 8192, dict(norm=True)),
        ("70b.o", 8192, 8192, dict(residual=True)),
        ("70b.gate_up", 28672, 8192, dict(norm=True, dual=True)),
        ("70b.down", 8192, 28672, dict(residual=True)),
        ("8b.qkv", 6144, 4096, dict(norm=True)),
        ("8b.gate_up", 14336, 4096, dict(norm=True, dual=True)),
        ("8b.down", 4096, 14336, dict(residual=True)),
    ]
    for name, N, K, opt in shapes:
        for M in (1, 4):
            nbuf = max(2, int(300e6 // (N * K * 2 * (2 if opt.get("dual") else 1))) + 1)
            ws = [torch.randn(N, K, device="cuda", dtype=torch.bfloat16) * K ** -0.5 for _ in range(nbuf)]
            w2s = [torch.randn(N, K, device="cuda", dtype=torch.bfloat16) * K ** -0.5 for _ in range(nbuf)] if opt.get("dual") else None
            x = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
            g = torch.ones(K, device="cuda", dtype=torch.bfloat16)
            res = torch.randn(M, N, device="cuda", dtype=torch.bfloat16)
            out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)

            def mk(i):
                kw = dict(out=out)
                if opt.get("norm"):
                    kw.update(norm_weight=g, norm_kind=Fn.NORM_RMS, eps=1e-5)
                if opt.get("residual"):
                    kw.update(residual=res)
                if opt.get("dual"):
                    kw.update(w2=w2s[i], act=Fn.ACT_SWIGLU)
                return lambda: Fn.linear_decode(x, ws[i], **kw)

            ms = time_fn([mk(i) for i in range(nbuf)], iters=40)
            nbytes = N * K * 2 * (2 if opt.get("dual") else 1)
            gbs = nbytes / ms / 1e6
            # cuBLAS baseline for the bare GEMV (no fusion)
            ref_ms = time_fn([(lambda i=i: torch.nn.functional.linear(x, ws[i])) for i in range(nbuf)], iters=40)
            row = dict(kernel="linear_decode", shape=name, M=M, N=N, K=K, ms=ms, GBps=gbs, frac_hbm=gbs / peaks["hbm_gbs"],
                       cublas_ms=ref_ms * (2 if opt.get("dual") else 1))
            results.append(row)
            print(f"gemv {name:12s} M={M} {ms * 1e3:8.1f} us  {gbs:7.0f} GB/s  {row['frac_hbm'] * 100:5.1f}% of measured HBM   (cuBLAS {row['cublas_ms'] * 1e3:.1f} us)", flush=True)
            del ws, w2s


def bench_gemm(results, peaks):
    shapes = [
        ("70b.qkv", 8192, 10240, 8192, {}),
        ("70b.o", 8192, 8192, 8192, dict(residual=True)),
        ("70b.gate_up", 8192, 28672, 8192, dict(dual=True)),
        ("70b.down", 8192, 8192, 28672, dict(residual=True)),
        ("8b.qkv", 8192, 6144, 4096, {}),
        ("sq4096", 4096, 4096, 4096, {}),
        ("sq8192", 8192, 8192, 8192, {}),
    ]
    for name, M, N, K, opt in shapes:
        nbuf = 3
        As = [torch.randn(M, K, device="cuda", dtype=torch.bfloat16) for _ in range(nbuf)]
        Bs = [torch.randn(N, K, device="cuda", dtype=torch.bfloat16) * K ** -0.5 for _ in range(nbuf)]
        B2 = [torch.randn(N, K, device="cuda", dtype=torch.bfloat16) * K ** -0.5 for _ in range(nbuf)] if opt.get("dual") else None
        res = torch.randn(M, N, device="cuda", dtype=torch.bfloat16) if opt.get("residual") else None
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)

        def mk(i):
            kw = dict(out=out)
            if opt.get("dual"):
                kw.update(b2=B2[i], act=Fn.ACT_SWIGLU)
            if res is not None:
                kw.update(residual=res)
            return lambda: Fn.gemm(As[i], Bs[i], **kw)

        ms = time_fn([mk(i) for i in range(nbuf)], iters=10)
        flops = 2.0 * M * N * K * (2 if opt.get("dual") else 1)
        tf = flops / ms / 1e9
        ref_ms = time_fn([(lambda i=i: torch.matmul(As[i], Bs[i].T)) for i in range(nbuf)], iters=10) * (2 if opt.get("dual") else 1)
        row = dict(kernel="gemm_tcgen05", shape=name, M=M, N=N, K=K, ms=ms, TFLOPs=tf, frac_bf16=tf / peaks["bf16_tflops"], cublas_ms=ref_ms,
                   cublas_TFLOPs=flops / ref_ms / 1e9)
        results.append(row)
        print(f"gemm {name:12s} {M}x{N}x{K} {ms:8.3f} ms {tf:7.0f} TFLOP/s {row['frac_bf16'] * 100:5.1f}% of measured cuBLAS peak (cuBLAS here: {row['cublas_TFLOPs']:.0f})", flush=True)
        del As, Bs, B2


def bench_skinny(results, peaks):
    """M = 2 / 4 / 8 rows: FMA kernel vs the mma.sync skinny-GEMM kernel on the same shapes."""
    shapes = [("70b.qkv", 10240, 8192, dict(norm=True)), ("70b.o", 8192, 8192, dict(residual=True)),
              ("70b.gate_up", 28672, 8192, dict(norm=True, dual=True)), ("70b.down", 8192, 28672, dict(residual=True))]
    for name, N, K, opt in shapes:
        nbuf = max(2, int(300e6 // (N * K * 2 * (2 if opt.get("dual") else 1))) + 1)
        ws = [torch.randn(N, K, device="cuda", dtype=torch.bfloat16) * K ** -0.5 for _ in range(nbuf)]
        w2s = [torch.randn(N, K, device="cuda", dtype=torch.bfloat16) * K ** -0.5 for _ in range(nbuf)] if opt.get("dual") else None
        g = torch.ones(K, device="cuda", dtype=torch.bfloat16)
        for M in (2, 4, 8):
            x = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
            res = torch.randn(M, N, device="cuda", dtype=torch.bfloat16)
            out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)

            def mk(i):
                kw = dict(out=out)
                if opt.get("norm"):
                    kw.update(norm_weight=g, norm_kind=Fn.NORM_RMS, eps=1e-5)
                if opt.get("residual"):
                    kw.update(residual=res)
                if opt.get("dual"):
                    kw.update(w2=w2s[i], act=Fn.ACT_SWIGLU)
                return lambda: Fn.linear_decode(x, ws[i], **kw)

            row = dict(kernel="linear_decode vs linear_decode_mma", shape=name, M=M, N=N, K=K)
            nbytes = N * K * 2 * (2 if opt.get("dual") else 1)
            for label, on in (("fma", False), ("mma", True)):
                Fn.set_skinny_gemm(on)
                try:
                    ms = time_fn([mk(i) for i in range(nbuf)], iters=30)
                    row[label + "_ms"], row[label + "_frac_hbm"] = ms, nbytes / ms / 1e6 / peaks["hbm_gbs"]
                except Exception as e:  # noqa: BLE001 - e.g. the FMA kernel cannot stage M*K activations
                    row[label + "_ms"], row[label + "_frac_hbm"] = None, None
                    print(f"  {label} {name} M={M}: {e!r}"[:200])
                finally:
                    Fn.set_skinny_gemm(False)
            results.append(row)
            f = lambda v: "   n/a" if v is None else f"{v * 100:5.1f}%"
            t = lambda v: "    n/a" if v is None else f"{v * 1e3:7.1f}"
            print(f"gemv.skinny {name:12s} M={M}  fma {t(row['fma_ms'])} us {f(row['fma_frac_hbm'])}   mma {t(row['mma_ms'])} us {f(row['mma_frac_hbm'])} of measured HBM", flush=True)
        del ws, w2s


def bench_gemv_fp8(results, peaks):
    """Decode linears over MXFP8 weights (1 byte/weight + 1 scale byte per 32)."""
    from petals_b200.ops.quant import quantize_mxfp8

    shapes = [("70b.qkv", 10240, 8192, dict(norm=True)), ("70b.o", 8192, 8192, dict(residual=True)),
              ("70b.gate_up", 28672, 8192, dict(norm=True, dual=True)), ("70b.down", 8192, 28672, dict(residual=True))]
    for name, N, K, opt in shapes:
        for M in (1, 4):
            if M * K * 2 > 200 * 1024:
                continue  # activations staged in shared memory
            nbuf = max(2, int(300e6 // (N * K * (2 if opt.get("dual") else 1))) + 1)
            qs = [quantize_mxfp8(torch.randn(N, K, device="cuda", dtype=torch.bfloat16) * K ** -0.5) for _ in range(nbuf)]
            q2s = [quantize_mxfp8(torch.randn(N, K, device="cuda", dtype=torch.bfloat16) * K ** -0.5) for _ in range(nbuf)] if opt.get("dual") else None
            x = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
            g = torch.ones(K, device="cuda", dtype=torch.bfloat16)
            res = torch.randn(M, N, device="cuda", dtype=torch.bfloat16)
            out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)

            def mk(i):
                kw = dict(out=out)
                if opt.get("norm"):
                    kw.update(norm_weight=g, norm_kind=Fn.NORM_RMS, eps=1e-5)
                if opt.get("residual"):
                    kw.update(residual=res)
                if opt.get("dual"):
                    kw.update(w2_q=q2s[i][0], w2_scale=q2s[i][1], act=Fn.ACT_SWIGLU)
                return lambda: Fn.linear_decode_fp8(x, qs[i][0], qs[i][1], **kw)

            ms = time_fn([mk(i) for i in range(nbuf)], iters=40)
            nbytes = N * K * (1 + 1 / 32) * (2 if opt.get("dual") else 1)
            gbs = nbytes / ms / 1e6
            row = dict(kernel="linear_decode_fp8", shape=name, M=M, N=N, K=K, ms=ms, GBps=gbs, frac_hbm=gbs / peaks["hbm_gbs"])
            results.append(row)
            print(f"gemv.fp8 {name:12s} M={M} {ms * 1e3:8.1f} us  {gbs:7.0f} GB/s  {row['frac_hbm'] * 100:5.1f}% of measured HBM", flush=True)
            del qs, q2s


def bench_gemm_2cta(results, peaks):
    """1-CTA vs 2-CTA (cta_group::2) tcgen05 GEMM vs cuBLAS on the 70B prefill shapes."""
    shapes = [("70b.qkv", 8192, 10240, 8192, {}), ("70b.o", 8192, 8192, 8192, dict(residual=True)), ("70b.gate_up", 8192, 28672, 8192, dict(dual=True)),
              ("70b.down", 8192, 8192, 28672, dict(residual=True)), ("8b.qkv", 8192, 6144, 4096, {}), ("sq4096", 4096, 4096, 4096, {})]
    for name, M, N, K, opt in shapes:
        nbuf = 2
        As = [torch.randn(M, K, device="cuda", dtype=torch.bfloat16) for _ in range(nbuf)]
        Bs = [torch.randn(N, K, device="cuda", dtype=torch.bfloat16) * K ** -0.5 for _ in range(nbuf)]
        B2 = [torch.randn(N, K, device="cuda", dtype=torch.bfloat16) * K ** -0.5 for _ in range(nbuf)] if opt.get("dual") else None
        res = torch.randn(M, N, device="cuda", dtype=torch.bfloat16) if opt.get("residual") else None
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)

        def mk(i):
            kw = dict(out=out)
            if opt.get("dual"):
                kw.update(b2=B2[i], act=Fn.ACT_SWIGLU)
            if res is not None:
                kw.update(residual=res)
            return lambda: Fn.gemm(As[i], Bs[i], **kw)

        flops = 2.0 * M * N * K * (2 if opt.get("dual") else 1)
        ms = {}
        for on in (False, True):
            Fn.set_gemm_2cta(on)
            ms[on] = time_fn([mk(i) for i in range(nbuf)], iters=10)
        Fn.set_gemm_2cta(True, fp8=False, mn=_MN_DEFAULT)
        ref_ms = time_fn([(lambda i=i: torch.matmul(As[i], Bs[i].T)) for i in range(nbuf)], iters=10) * (2 if opt.get("dual") else 1)
        row = dict(kernel="gemm_tcgen05_2cta", shape=name, M=M, N=N, K=K, ms_1cta=ms[False], ms_2cta=ms[True], TFLOPs_1cta=flops / ms[False] / 1e9,
                   TFLOPs_2cta=flops / ms[True] / 1e9, cublas_TFLOPs=flops / ref_ms / 1e9)
        results.append(row)
        print(f"gemm2 {name:12s} {M}x{N}x{K}  1-CTA {ms[False]:7.3f} ms {row['TFLOPs_1cta']:6.0f} TFLOP/s | 2-CTA {ms[True]:7.3f} ms {row['TFLOPs_2cta']:6.0f} TFLOP/s | "
              f"cuBLAS {row['cublas_TFLOPs']:6.0f} | 2-CTA / cuBLAS {row['TFLOPs_2cta'] / row['cublas_TFLOPs']:.2f}", flush=True)
        del As, Bs, B2


def bench_gemm_fp8(results, peaks):
    """Block-scaled MXFP8 GEMM (tcgen05.mma kind::mxf8f6f4.block_scale) on the 70B prefill shapes, the bf16 tcgen05 GEMM alongside;
    the activation quantiser (what the fp8 path pays extra per projection) is timed separately."""
    from petals_b200.ops.quant import pack_scales, quantize_mxfp8

    shapes = [("70b.qkv", 8192, 10240, 8192, False), ("70b.o", 8192, 8192, 8192, False), ("70b.gate_up", 8192, 28672, 8192, True),
              ("70b.down", 8192, 8192, 28672, False), ("sq4096", 4096, 4096, 4096, False)]
    for name, M, N, K, dual in shapes:
        nbuf = 2
        A = [torch.randn(M, K, device="cuda", dtype=torch.bfloat16) for _ in range(nbuf)]
        W = [torch.randn(N, K, device="cuda", dtype=torch.bfloat16) * K ** -0.5 for _ in range(nbuf * (2 if dual else 1))]
        Aq = [Fn.quant_mxfp8(a) for a in A]
        Wq = [(lambda qe: (qe[0].view(torch.uint8), pack_scales(qe[1])))(quantize_mxfp8(w)) for w in W]
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)

        def mk(i):
            if dual:
                return lambda: Fn.gemm_mxfp8(Aq[i][0], Aq[i][1], Wq[2 * i][0], Wq[2 * i][1], b2_q=Wq[2 * i + 1][0], b2_sf=Wq[2 * i + 1][1], out=out)
            return lambda: Fn.gemm_mxfp8(Aq[i][0], Aq[i][1], Wq[i][0], Wq[i][1], out=out)

        def mk16(i):
            if dual:
                return lambda: Fn.gemm(A[i], W[2 * i], b2=W[2 * i + 1], act=Fn.ACT_SWIGLU, out=out)
            return lambda: Fn.gemm(A[i], W[i], out=out)

        ms = time_fn([mk(i) for i in range(nbuf)], iters=10)
        ms16 = time_fn([mk16(i) for i in range(nbuf)], iters=10)
        qms = time_fn([(lambda i=i: Fn.quant_mxfp8(A[i], q=Aq[i][0], sf=Aq[i][1])) for i in range(nbuf)], iters=10)
        flops = 2.0 * M * N * K * (2 if dual else 1)
        row = dict(kernel="gemm_mxfp8", shape=name, M=M, N=N, K=K, ms=ms, TFLOPs=flops / ms / 1e9, bf16_ms=ms16, bf16_TFLOPs=flops / ms16 / 1e9,
                   quant_ms=qms, quant_GBps=M * K * 3 / qms / 1e6, speedup_vs_bf16=ms16 / ms, speedup_incl_quant=ms16 / (ms + qms))
        results.append(row)
        print(f"gemm_mxfp8 {name:12s} {M}x{N}x{K} {ms:8.3f} ms {row['TFLOPs']:7.0f} TFLOP/s | bf16 tcgen05 {ms16:8.3f} ms {row['bf16_TFLOPs']:7.0f} | "
              f"quantise A {qms:6.3f} ms | x{row['speedup_vs_bf16']:.2f} (x{row['speedup_incl_quant']:.2f} with the quantiser)", flush=True)
        del A, W, Aq, Wq


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--fp8", action="store_true", help="only the MXFP8 decode linears")
    args = ap.parse_args()
    peaks = measured_peaks()
    results = []
    if args.fp8:
        args.only = "fp8"
        bench_gemv_fp8(results, peaks)
    if args.only == "skinny":
        bench_skinny(results, peaks)
    if args.only in ("", "gemv"):
        bench_gemv(results, peaks)
    if args.only in ("", "gemm"):
        bench_gemm(results, peaks)
    if args.only in ("", "gemm_2cta"):
        bench_gemm_2cta(results, peaks)
    if args.only in ("", "gemm_fp8"):
        bench_gemm_fp8(results, peaks)
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/kernel_bench{('_' + args.only) if args.only else ''}.json", "w") as f:
        json.dump(dict(peaks=peaks, results=results), f, indent=1)


if __name__ == "__main__":
    main()
