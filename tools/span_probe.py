#!/usr/bin/env python
"""Where does a step of the persistent span kernel (csrc/decode_span.cu) spend its time?

Builds random shards of the given shape directly (no server), runs the kernel for a few steps and prints
  * the step time, the bytes streamed and the fraction of the measured HBM copy bandwidth;
  * CTA 0's %globaltimer stamps at every phase boundary of a middle block (consumer and producer side);
and, with PETALS_B200_SPAN_DEBUG=1/2/3 in the environment (set by the caller, one process per mode), the same with polls
short-circuited and/or the math skipped — which separates streaming, computing and waiting."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from petals_b200.ops import functional as Fn  # noqa: E402
from petals_b200.utils.peaks import measured_peaks  # noqa: E402

SHAPES = {  # H, Hq, Hkv, D, I, layers
    "70b": (8192, 64, 8, 128, 28672, 20), "70b-tp8": (8192, 8, 1, 128, 3584, 80), "70b-tp4": (8192, 16, 2, 128, 7168, 80),
    "8b": (4096, 32, 8, 128, 14336, 32),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="70b-tp8")
    ap.add_argument("--pos", type=int, default=130)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--layers", type=int, default=0, help="override the number of blocks (0: the shape's default)")
    args = ap.parse_args()
    H, Hq, Hkv, D, I, L = SHAPES[args.shape]
    L = args.layers or L
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(0)
    rnd = lambda *s: (torch.randn(*s, device=dev, generator=g) * 0.02).to(torch.bfloat16)
    pages = args.pos // 64 + 2
    layers = []
    for _ in range(L):
        layers.append(dict(wqkv=rnd((Hq + 2 * Hkv) * D, H), wo=rnd(H, Hq * D), w_gate=rnd(I, H), w_up=rnd(I, H), w_down=rnd(H, I),
                           ln1_w=torch.ones(H, dtype=torch.bfloat16, device=dev), ln2_w=torch.ones(H, dtype=torch.bfloat16, device=dev),
                           k_pool=rnd(pages, Hkv, 64, D), v_pool=rnd(pages, Hkv, 64, D)))
    plan = Fn.DecodeSpanPlan(layers, H=H, Hq=Hq, Hkv=Hkv, D=D, I=I, eps=1e-5, attn_scale=D ** -0.5, max_chunks=pages, device=dev)
    plan.timing = torch.zeros(L, 24, dtype=torch.int64, device=dev)
    cos, sin = Fn.rope_tables(D, 4096, 500000.0, None, device=dev)
    table = torch.arange(pages, dtype=torch.int32, device=dev).view(1, pages)
    pos = torch.tensor([args.pos], dtype=torch.int32, device=dev)
    x, y = rnd(1, H), torch.empty(1, H, dtype=torch.bfloat16, device=dev)
    for _ in range(3):
        Fn.decode_span(plan, x, y, table, pos.data_ptr(), cos, sin)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        Fn.decode_span(plan, x, y, table, pos.data_ptr(), cos, sin)
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / args.steps
    nbytes = L * 2 * ((Hq + 2 * Hkv) * D * H + H * Hq * D + 3 * I * H)
    peaks = measured_peaks()
    t = plan.timing.cpu().numpy()
    mid = L // 2
    c = [int(v) for v in t[mid, :13]]
    names = ["norm1", "qkv", "attn units", "combine", "gather attn", "o-proj", "reduce1", "gather+norm2", "gate/up", "gather act", "down", "reduce2+gather"]
    phases = {n: round((c[i + 1] - c[i]) / 1e3, 2) for i, n in enumerate(names)}
    prod = {"all_issued": round((int(t[mid, 20]) - c[0]) / 1e3, 2)}
    print(json.dumps({"shape": args.shape, "debug": int(os.environ.get("PETALS_B200_SPAN_DEBUG", "0")), "ms_per_step": round(ms, 3),
                      "us_per_layer": round(1e3 * ms / L, 1), "GBps": round(nbytes / ms / 1e6, 1), "frac_hbm": round(nbytes / ms / 1e6 / peaks["hbm_gbs"], 3),
                      "layer_us": round((int(t[mid + 1, 0]) - c[0]) / 1e3, 2), "consumer_phase_us": phases,
                      "producer_done_issuing_at_us": prod, "error_flag": int(plan.err.item())}))


if __name__ == "__main__":
    main()
