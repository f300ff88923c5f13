"""Block-scaled FP8 serving (--quant_type fp8) through the engine: close to the bf16 model, exact vs its own dequantised oracle."""
import pytest
import torch

from petals_b200.parallel.swarm import Swarm
from petals_b200.utils.convert_block import QuantType
from petals_b200.utils.random_model import launch_random_stage, random_client_model, write_config_only

pytestmark = pytest.mark.gpu
_MN_DEFAULT = __import__("os").environ.get("PETALS_B200_GEMM_2CTA_MN", "1") != "0"
DEV = "cuda:0"


@pytest.mark.parametrize("prefill", ["w8a8", "dequant"])
def test_fp8_stage_matches_dequantised_oracle(prefill, tmp_path, monkeypatch):
    """Decode streams the MXFP8 weights (bf16 activations). Prefill: "w8a8" quantises the activations too and multiplies on the
    block-scaled tensor-core path (format error of both operands: looser bound, arg-max agreement required); "dequant" feeds the bf16
    GEMM with dequantised weights (tight bound)."""
    monkeypatch.setenv("PETALS_B200_FP8_PREFILL", prefill)
    path = write_config_only("llama-tiny", {}, str(tmp_path / "m"))
    swarm = Swarm(f"t-fp8-{prefill}")
    stage = launch_random_stage(path, range(4), swarm, DEV, quant_type=QuantType.FP8)
    try:
        eng = stage.stage.engine
        assert eng is not None and eng.fp8 is not None, "fp8 stages must run on the sm_100a engine"
        assert eng.fp8_w8a8 == (prefill == "w8a8")
        assert all(getattr(stage.stage.blocks[0], n).numel() == 0 for n in eng.fp8[0]), "bf16 copies must be released"
        model = random_client_model(path, swarm, DEV)
        ids = torch.randint(0, 4000, (1, 24), device=DEV)
        with torch.inference_mode():
            with model.inference_session(max_length=64):
                a = model(ids[:, :20]).logits  # prefill: dequant scratch + tcgen05 GEMM
                b = torch.cat([model(ids[:, t: t + 1]).logits for t in range(20, 24)], 1)  # decode: fp8 weight streamer
            sess = torch.cat([a, b], 1).float()
            # oracle on the dequantised weights
            h = model.model.embed(ids)
            for i, blk in enumerate(stage.stage.blocks):
                stage.stage._materialize(i)
                h = blk.forward_cached(h, None, None, 0)
                stage.stage._dematerialize(i)
            ref = model.lm_head(model.model.final_norm(h)).float()
        scale = ref.abs().mean().item()
        assert (sess - ref).abs().mean().item() < (0.10 if prefill == "w8a8" else 0.05) * scale + 1e-3
        assert (sess.argmax(-1) == ref.argmax(-1)).float().mean().item() >= 0.9
        # gradients still flow (weights are dequantised block by block for the backward pass)
        emb = model.model.embed(ids).float().requires_grad_(True)
        model.model.layers(emb.to(torch.bfloat16)).float().pow(2).mean().backward()
        assert torch.isfinite(emb.grad).all() and emb.grad.abs().sum() > 0
    finally:
        stage.shutdown()


# ---- the block-scaled tensor-core GEMM (csrc/gemm_mxfp8.cu) and its activation quantiser ------------------------------------------
def _deq(q, e):
    from petals_b200.ops.quant import dequantize_mxfp8

    return dequantize_mxfp8(q.view(torch.float8_e4m3fn), e, torch.float32)


@pytest.mark.parametrize("M,K", [(1, 128), (200, 1024), (300, 4096)])
def test_activation_quantiser_matches_the_format_definition(M, K):
    """pb_quant_mxfp8 == ops/quant.py:quantize_mxfp8 + pack_scales, bit for bit (payload and scale exponents)."""
    from petals_b200.ops import functional as Fn
    from petals_b200.ops.quant import pack_scales, quantize_mxfp8, unpack_scales

    torch.manual_seed(3)
    x = (torch.randn(M, K, device=DEV) * torch.rand(M, 1, device=DEV) * 3).to(torch.bfloat16)
    x[0, :32] = 0  # an all-zero group: exponent -127, payload 0
    q, sf = Fn.quant_mxfp8(x)
    q_ref, e_ref = quantize_mxfp8(x)
    e = unpack_scales(sf, M, K)
    nz = x.float().view(M, K // 32, 32).abs().amax(-1) > 0
    assert torch.equal(e[nz], e_ref[nz])
    assert torch.equal(_deq(q, e), _deq(q_ref.view(torch.uint8), e_ref).where(nz.repeat_interleave(32, 1), torch.zeros((), device=DEV)))
    assert torch.equal(sf, pack_scales(torch.where(nz, e_ref, torch.zeros_like(e_ref))))  # the block layout, padding rows included
    # fused RMSNorm: same as quantising the bf16 norm output
    w = (1 + 0.1 * torch.randn(K, device=DEV)).to(torch.bfloat16)
    qn, sfn = Fn.quant_mxfp8(x, w, 1e-5)
    xn = Fn.norm_ref(x, w, None, Fn.NORM_RMS, 1e-5)
    q2, e2 = quantize_mxfp8(xn)
    got, want = _deq(qn, unpack_scales(sfn, M, K)), _deq(q2.view(torch.uint8), e2)
    assert (got - want).abs().mean().item() < 0.01 * want.abs().mean().item() + 1e-6  # bf16 rounding of the norm may flip a payload step


@pytest.mark.parametrize("M,N,K", [(128, 256, 128), (200, 384, 1024), (777, 1000, 512), (1024, 4096, 4096)])
def test_block_scaled_gemm_matches_dequantised_matmul(M, N, K):
    """tcgen05.mma kind::mxf8f6f4.block_scale applies both operands' per-32 scales in hardware: the result must equal the fp32 matmul
    of the dequantised operands up to accumulation order and the bf16 output rounding."""
    from petals_b200.ops import functional as Fn
    from petals_b200.ops.quant import pack_scales, quantize_mxfp8

    torch.manual_seed(4)
    a = (torch.randn(M, K, device=DEV) * (0.2 + torch.rand(M, 1, device=DEV))).to(torch.bfloat16)
    w = (torch.randn(N, K, device=DEV) * 0.05 * (0.2 + torch.rand(N, 1, device=DEV))).to(torch.bfloat16)
    aq, asf = Fn.quant_mxfp8(a)
    wq, we = quantize_mxfp8(w)
    wq = wq.view(torch.uint8)
    from petals_b200.ops.quant import unpack_scales

    ref = _deq(aq, unpack_scales(asf, M, K)) @ _deq(wq, we).T
    out = Fn.gemm_mxfp8(aq, asf, wq, pack_scales(we)).float()
    err = (out - ref).abs()
    assert err.max().item() < 2e-2 * ref.abs().max().item() and err.mean().item() < 4e-3 * ref.abs().mean().item(), (err.max().item(), err.mean().item())
    # residual epilogue
    res = torch.randn(M, N, device=DEV).to(torch.bfloat16)
    out = Fn.gemm_mxfp8(aq, asf, wq, pack_scales(we), residual=res).float()
    assert ((out - (ref.to(torch.bfloat16).float() + res.float())).abs().mean().item()) < 6e-3 * (ref.abs().mean().item() + 1)
    # close to the unquantised product too (format error, not kernel error)
    full = a.float() @ w.float().T
    assert (ref - full).abs().mean().item() < 0.06 * full.abs().mean().item()


def test_block_scaled_gemm_swiglu_epilogue():
    from petals_b200.ops import functional as Fn
    from petals_b200.ops.quant import pack_scales, quantize_mxfp8, unpack_scales

    torch.manual_seed(5)
    M, I, K = 300, 640, 1024
    a = torch.randn(M, K, device=DEV).to(torch.bfloat16)
    wg, wu = [(torch.randn(I, K, device=DEV) * 0.04).to(torch.bfloat16) for _ in range(2)]
    aq, asf = Fn.quant_mxfp8(a)
    (gq, ge), (uq, ue) = quantize_mxfp8(wg), quantize_mxfp8(wu)
    ad = _deq(aq, unpack_scales(asf, M, K))
    g, u = (ad @ _deq(gq.view(torch.uint8), ge).T).to(torch.bfloat16), (ad @ _deq(uq.view(torch.uint8), ue).T).to(torch.bfloat16)
    ref = (torch.nn.functional.silu(g.float()).to(torch.bfloat16).float() * u.float())
    out = Fn.gemm_mxfp8(aq, asf, gq.view(torch.uint8), pack_scales(ge), b2_q=uq.view(torch.uint8), b2_sf=pack_scales(ue)).float()
    assert (out - ref).abs().mean().item() < 1e-2 * ref.abs().mean().item() + 1e-4


@pytest.mark.parametrize("M,N,K", [(1024, 1024, 512), (1100, 2048 + 128, 1024), (4096, 4096, 4096)])
def test_block_scaled_gemm_2cta_matches_the_1cta_kernel(M, N, K):
    """The cta_group::2 variant (one 256 x 256 tile per SM pair, each SM stages half of the weight tile, scale factors of all 256 weight rows
    in both SMs' tensor memory) against the 1-CTA kernel: plain, residual and SwiGLU epilogues, ragged M / N tails."""
    from petals_b200.ops import functional as Fn
    from petals_b200.ops.quant import pack_scales, quantize_mxfp8

    torch.manual_seed(6)
    a = (torch.randn(M, K, device=DEV) * (0.2 + torch.rand(M, 1, device=DEV))).to(torch.bfloat16)
    ws = [(torch.randn(N, K, device=DEV) * 0.05 * (0.2 + torch.rand(N, 1, device=DEV))).to(torch.bfloat16) for _ in range(2)]
    res = torch.randn(M, N, device=DEV).to(torch.bfloat16)
    aq, asf = Fn.quant_mxfp8(a)
    (q1, e1), (q2, e2) = [quantize_mxfp8(w) for w in ws]
    q1, q2, s1, s2 = q1.view(torch.uint8), q2.view(torch.uint8), pack_scales(e1), pack_scales(e2)

    def run():
        return [Fn.gemm_mxfp8(aq, asf, q1, s1).float(), Fn.gemm_mxfp8(aq, asf, q1, s1, residual=res).float(),
                Fn.gemm_mxfp8(aq, asf, q1, s1, b2_q=q2, b2_sf=s2).float()]

    try:
        Fn.set_gemm_2cta(True, fp8=False, mn=_MN_DEFAULT)
        want = run()
        Fn.set_gemm_2cta(True, fp8=True)
        got = run()
    finally:
        Fn.set_gemm_2cta(True, fp8=False, mn=_MN_DEFAULT)  # the defaults
    for g, w, what in zip(got, want, ("plain", "residual", "swiglu")):
        assert (g - w).abs().max().item() <= 2e-2 * w.abs().max().item(), what
        assert (g - w).abs().mean().item() <= 2e-3 * w.abs().mean().item() + 1e-6, what
