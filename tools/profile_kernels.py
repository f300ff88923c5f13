"""Launch one hot kernel a few times at a Llama-3-70B shape (for `ncu -k regex:... -s N -c 1`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from petals_b200.ops import functional as Fn  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "gemm"
torch.manual_seed(0)
if which == "gemm":  # gate/up projection with fused SwiGLU, 8192 tokens
    a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
    wg = torch.randn(28672, 8192, device="cuda", dtype=torch.bfloat16) * 0.01
    wu = torch.randn(28672, 8192, device="cuda", dtype=torch.bfloat16) * 0.01
    for _ in range(4):
        Fn.gemm(a, wg, b2=wu, act=Fn.ACT_SWIGLU)
elif which == "gemm_plain":
    a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
    w = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16) * 0.01
    for _ in range(4):
        Fn.gemm(a, w)
elif which == "gemm_2cta":  # the same gate/up GEMM on the cta_group::2 kernel
    Fn.set_gemm_2cta(True)
    a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
    wg = torch.randn(28672, 8192, device="cuda", dtype=torch.bfloat16) * 0.01
    wu = torch.randn(28672, 8192, device="cuda", dtype=torch.bfloat16) * 0.01
    for _ in range(4):
        Fn.gemm(a, wg, b2=wu, act=Fn.ACT_SWIGLU)
elif which in ("gemm_mxfp8", "gemm_mxfp8_2cta"):  # block-scaled FP8 down projection (K = 28672), 8192 tokens
    from petals_b200.ops.quant import pack_scales, quantize_mxfp8

    Fn.set_gemm_2cta(which.endswith("2cta"))
    a = torch.randn(8192, 28672, device="cuda", dtype=torch.bfloat16)
    q, e = quantize_mxfp8(torch.randn(8192, 28672, device="cuda", dtype=torch.bfloat16) * 0.01)
    aq, asf = Fn.quant_mxfp8(a)
    wq, wsf = q.view(torch.uint8), pack_scales(e)
    for _ in range(4):
        Fn.gemm_mxfp8(aq, asf, wq, wsf)
elif which == "gemv":  # decode gate/up with fused RMSNorm + SwiGLU, 1 token
    x = torch.randn(1, 8192, device="cuda", dtype=torch.bfloat16)
    g = torch.ones(8192, device="cuda", dtype=torch.bfloat16)
    ws = [(torch.randn(28672, 8192, device="cuda", dtype=torch.bfloat16) * 0.01, torch.randn(28672, 8192, device="cuda", dtype=torch.bfloat16) * 0.01) for _ in range(2)]
    for i in range(4):
        Fn.linear_decode(x, ws[i % 2][0], w2=ws[i % 2][1], act=Fn.ACT_SWIGLU, norm_weight=g, norm_kind=Fn.NORM_RMS, eps=1e-5)
elif which == "gemv_fp8":  # decode O-projection over MXFP8 weights, 1 token
    from petals_b200.ops.quant import quantize_mxfp8

    x = torch.randn(1, 8192, device="cuda", dtype=torch.bfloat16)
    res = torch.randn(1, 8192, device="cuda", dtype=torch.bfloat16)
    qs = [quantize_mxfp8(torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16) * 0.01) for _ in range(4)]
    for i in range(6):
        Fn.linear_decode_fp8(x, qs[i % 4][0], qs[i % 4][1], residual=res)
elif which == "gemv_o":  # bf16 twin of gemv_fp8 (same shape)
    x = torch.randn(1, 8192, device="cuda", dtype=torch.bfloat16)
    res = torch.randn(1, 8192, device="cuda", dtype=torch.bfloat16)
    ws = [torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16) * 0.01 for _ in range(4)]
    for i in range(6):
        Fn.linear_decode(x, ws[i % 4], residual=res)
elif which == "attn":  # prefill attention, 4096 tokens, GQA 64/8
    B, T, Hq, Hkv, D = 1, 4096, 64, 8, 128
    pages = T // Fn.PAGE
    k_pool = torch.randn(pages, Hkv, Fn.PAGE, D, device="cuda", dtype=torch.bfloat16)
    v_pool = torch.randn_like(k_pool)
    table = torch.arange(pages, dtype=torch.int32, device="cuda").view(1, pages)
    q = torch.randn(B * T, Hq * D, device="cuda", dtype=torch.bfloat16)
    out = torch.empty_like(q)
    impl = int(os.environ.get("ATTN_IMPL", "0"))
    for _ in range(3):
        Fn.paged_attention(q, k_pool, v_pool, table, None, out, B=B, T=T, Hq=Hq, Hkv=Hkv, D=D, scale=D ** -0.5, pos_static=0, impl=impl)
elif which == "attn_time":  # CUDA-event timing of both attention kernels at the 70B prefill shape
    import json

    B, T, Hq, Hkv, D = 1, 4096, 64, 8, 128
    pages = T // Fn.PAGE
    k_pool = torch.randn(pages, Hkv, Fn.PAGE, D, device="cuda", dtype=torch.bfloat16)
    v_pool = torch.randn_like(k_pool)
    table = torch.arange(pages, dtype=torch.int32, device="cuda").view(1, pages)
    q = torch.randn(B * T, Hq * D, device="cuda", dtype=torch.bfloat16)
    flops = 4.0 * T * T * Hq * D / 2
    res = {}
    outs = {}
    for impl in (1, 2):
        out = torch.empty_like(q)
        for _ in range(3):
            Fn.paged_attention(q, k_pool, v_pool, table, None, out, B=B, T=T, Hq=Hq, Hkv=Hkv, D=D, scale=D ** -0.5, pos_static=0, impl=impl)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(10):
            Fn.paged_attention(q, k_pool, v_pool, table, None, out, B=B, T=T, Hq=Hq, Hkv=Hkv, D=D, scale=D ** -0.5, pos_static=0, impl=impl)
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 10
        res["mma_sync" if impl == 1 else "tcgen05"] = {"ms": round(ms, 4), "TFLOPs": round(flops / ms / 1e9, 1)}
        outs[impl] = out.float()
    res["max_abs_diff"] = round((outs[1] - outs[2]).abs().max().item(), 5)
    print(json.dumps({"attention_prefill_4096x64h_d128_causal": res}))
torch.cuda.synchronize()
print("done", which)
