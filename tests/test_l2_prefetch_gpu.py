"""Layer-ahead L2 prefetcher (csrc/l2_prefetch.cu). The kernel was written after this round's GPU budget was spent: these
tests run only with PETALS_B200_RUN_UNVALIDATED=1 until its first hardware run has been looked at."""
import os
import time

import pytest
import torch

from petals_b200.ops import functional as Fn

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(os.environ.get("PETALS_B200_RUN_UNVALIDATED", "0") != "1",
                                                  reason="first hardware run pending: opt in with PETALS_B200_RUN_UNVALIDATED=1")]
DEV = "cuda"


def _layers(n_layers=4, per_layer=3, elems=1 << 20):
    g = torch.Generator(device=DEV).manual_seed(0)
    return [[torch.randn(elems + 13 * i, device=DEV, generator=g).to(torch.bfloat16) for i in range(per_layer)] for _ in range(n_layers)]


def test_unpaced_walk_touches_nothing():
    layers = _layers()
    before = [[t.clone() for t in ts] for ts in layers]
    plan = Fn.L2PrefetchPlan(layers)
    assert plan.n_layers == 4 and plan.per_layer == 3 and plan.total_bytes == sum(t.numel() * 2 for ts in layers for t in ts)
    Fn.l2_prefetch(plan, ctas=4)
    torch.cuda.synchronize()
    assert all(torch.equal(a, b) for ta, tb in zip(layers, before) for a, b in zip(ta, tb))


def test_paced_walk_follows_the_tag_and_times_out_without_hanging():
    layers = _layers(n_layers=6)
    plan = Fn.L2PrefetchPlan(layers)
    L = 6
    epoch = torch.tensor([7], dtype=torch.int64, device=DEV)
    unit = torch.zeros(2, dtype=torch.int32, device=DEV)  # {payload, tag}
    # the producer is already at the last layer: the walk never waits
    unit[1] = 7 * L + (L - 1)
    Fn.l2_prefetch(plan, progress_ptr=unit.data_ptr(), epoch_ptr=epoch.data_ptr(), tag_mul=L, lookahead=1, ctas=4)
    torch.cuda.synchronize()
    # a stale tag from the previous step never releases layer 2: the watchdog ends the kernel after ~wait_us
    unit[1] = 6 * L + (L - 1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    Fn.l2_prefetch(plan, progress_ptr=unit.data_ptr(), epoch_ptr=epoch.data_ptr(), tag_mul=L, lookahead=1, ctas=4, wait_us=500)
    torch.cuda.synchronize()
    assert time.perf_counter() - t0 < 1.0
    # a side stream can be released mid-walk by the "step" advancing the tag
    side = torch.cuda.Stream()
    unit[1] = 7 * L + 0
    torch.cuda.synchronize()
    with torch.cuda.stream(side):
        Fn.l2_prefetch(plan, progress_ptr=unit.data_ptr(), epoch_ptr=epoch.data_ptr(), tag_mul=L, lookahead=1, ctas=4, wait_us=2_000_000)
    for l in range(1, L):
        unit[1] = 7 * L + l
        torch.cuda.current_stream().synchronize()
    side.synchronize()


def test_bad_arguments_are_rejected():
    plan = Fn.L2PrefetchPlan(_layers(2, 2, 1024))
    with pytest.raises(Exception):
        Fn.l2_prefetch(plan, ctas=0)
    with pytest.raises(Exception):
        Fn.l2_prefetch(plan, progress_ptr=plan.ranges.data_ptr(), epoch_ptr=0, tag_mul=2)
