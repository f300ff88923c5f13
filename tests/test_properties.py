"""Property-based checks (hypothesis) of the small pure functions everything else leans on: argument packing, the uid grammar,
the wire codecs' shape/dtype contract, sampling helpers. Kept small and deterministic (derandomised, few examples)."""
import msgpack
import torch
from hypothesis import given, settings, strategies as st

from petals_b200.data_structures import join_uids, make_uid, parse_uid, split_uids
from petals_b200.utils.compression import CompressionType, roundtrip
from petals_b200.utils.packaging import pack_args_kwargs, unpack_args_kwargs
from petals_b200.utils.random import sample_up_to

FAST = settings(max_examples=60, deadline=None, derandomize=True)

leaves = st.one_of(st.none(), st.booleans(), st.integers(-2 ** 40, 2 ** 40), st.floats(allow_nan=False, allow_infinity=False, width=32),
                   st.text(max_size=8), st.binary(max_size=8), st.sampled_from(["__T0", "__T", b"__T0", b"__T7", b"__E__T1", b"__E"]),
                   st.builds(lambda n: torch.arange(n, dtype=torch.float32), st.integers(0, 5)))
trees = st.recursive(leaves, lambda kids: st.one_of(st.lists(kids, max_size=4), st.dictionaries(st.text(max_size=4), kids, max_size=4)), max_leaves=12)


def _same(a, b) -> bool:
    if isinstance(a, torch.Tensor) or isinstance(b, torch.Tensor):
        return isinstance(a, torch.Tensor) and isinstance(b, torch.Tensor) and a is b
    if isinstance(a, (list, tuple)) and isinstance(b, (list, tuple)):
        return len(a) == len(b) and all(_same(x, y) for x, y in zip(a, b))
    if isinstance(a, dict) and isinstance(b, dict):
        return a.keys() == b.keys() and all(_same(a[k], b[k]) for k in a)
    return type(a) is type(b) and a == b


@FAST
@given(st.lists(trees, max_size=3), st.dictionaries(st.text(min_size=1, max_size=4), trees, max_size=3))
def test_pack_unpack_round_trips_any_argument_tree(args, kwargs):
    """Also through msgpack (how the structure travels), and for text / bytes arguments that look like tensor placeholders."""
    tensors, structure = pack_args_kwargs(*args, **kwargs)
    assert all(isinstance(t, torch.Tensor) for t in tensors)
    wire = msgpack.unpackb(msgpack.packb(structure, use_bin_type=True), raw=False)
    got_args, got_kwargs = unpack_args_kwargs(tensors, wire)
    assert _same(got_args, list(args)) and _same(got_kwargs, kwargs)


@FAST
@given(st.text(alphabet=st.characters(blacklist_characters=" ", blacklist_categories=("Cs",)), min_size=1, max_size=12), st.integers(0, 10 ** 6),
       st.integers(1, 6))
def test_uid_grammar_round_trips(prefix, index, n):
    uid = make_uid(prefix, index)
    assert parse_uid(uid) == (prefix, index)
    chain = [make_uid(prefix, index + i) for i in range(n)]
    assert split_uids(join_uids(chain)) == chain


@FAST
@given(st.sampled_from(list(CompressionType)), st.lists(st.integers(1, 7), min_size=1, max_size=3), st.sampled_from([torch.float32, torch.bfloat16, torch.float16]),
       st.integers(0, 2 ** 31 - 1))
def test_every_codec_preserves_shape_dtype_and_finiteness(codec, shape, dtype, seed):
    x = (torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * 3).to(dtype)
    y = roundtrip(x, codec)
    assert y.shape == x.shape and y.dtype == x.dtype and torch.isfinite(y.float()).all()
    if codec == CompressionType.NONE:
        assert torch.equal(x, y)


@FAST
@given(st.lists(st.integers(), max_size=20, unique=True), st.integers(-2, 25))
def test_sample_up_to_is_an_ordered_subset_of_the_right_size(population, k):
    chosen = sample_up_to(population, k)
    assert len(chosen) == min(len(population), max(k, 0)) and len(set(chosen)) == len(chosen)
    positions = [population.index(c) for c in chosen]
    assert positions == sorted(positions)


@given(st.integers(1, 5), st.integers(1, 7), st.sets(st.tuples(st.integers(0, 4), st.integers(0, 6)), max_size=6), st.booleans())
@settings(max_examples=40, deadline=None)
def test_wavefront_keeps_item_order_on_every_stage_and_skips_failed_items(n_stages, n_items, failures, threaded):
    """client/pipeline.py:run_wave — whatever fails where: every stage sees the items it processes in item order, an item visits the
    stages in stage order, an item whose work raised is skipped by all later stages and carries the error, nothing is lost or duplicated."""
    import threading

    from petals_b200.client.pipeline import run_wave

    class Item:
        def __init__(self, index):
            self.index, self.error, self.detached, self.visited = index, None, False, []

    items = [Item(i) for i in range(n_items)]
    seen = [[] for _ in range(n_stages)]
    lock = threading.Lock()

    def stage_work(s):
        def work(it):
            with lock:
                seen[s].append(it.index)
            it.visited.append(s)
            if (s, it.index) in failures:
                raise RuntimeError(f"stage {s} refuses item {it.index}")
        return work

    run_wave(items, [stage_work(s) for s in range(n_stages)], threaded=threaded)
    for s in range(n_stages):
        assert seen[s] == sorted(seen[s]) and len(set(seen[s])) == len(seen[s])  # item order, no duplicates
    for it in items:
        first_fail = min((s for s in range(n_stages) if (s, it.index) in failures), default=None)
        expect = list(range(n_stages)) if first_fail is None else list(range(first_fail + 1))
        assert it.visited == expect and (it.error is None) == (first_fail is None)


@given(st.integers(1, 300), st.integers(1, 6))
@settings(max_examples=30, deadline=None)
def test_mxfp8_scale_layout_round_trips_for_any_shape(rows, kblocks128):
    from petals_b200.ops.quant import pack_scales, unpack_scales

    K = 128 * kblocks128
    e = torch.randint(0, 256, (rows, K // 32), dtype=torch.uint8)
    packed = pack_scales(e)
    assert packed.numel() == kblocks128 * ((rows + 127) // 128) * 512
    assert torch.equal(unpack_scales(packed, rows, K), e)
