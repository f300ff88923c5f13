"""Small utility modules of the reference's inventory that no other test touches by name: disk cache housekeeping, asyncio
shielding, reachability, spending policy, deprecated aliases, auth plumbing, dtype map."""
import asyncio
import os
import time
import warnings

import pytest
import torch


def test_disk_cache_lru_eviction_and_locks(tmp_path):
    from petals_b200.utils.disk_cache import allow_cache_reads, allow_cache_writes, free_disk_space_for

    cache = tmp_path / "cache"
    for i, name in enumerate(["old-model", "mid-model", "new-model"]):
        d = cache / name
        d.mkdir(parents=True)
        (d / "weights.bin").write_bytes(b"x" * 1000)
        t = time.time() - 1000 + 100 * i
        os.utime(d, (t, t))
    with allow_cache_reads(str(cache)), allow_cache_reads(str(cache)):  # shared locks stack
        pass
    with allow_cache_writes(str(cache)):
        free_disk_space_for(1500, cache_dir=str(cache), max_disk_space=3500, os_quota=0)  # 3000 used: evict just the LRU entry
    assert not (cache / "old-model").exists() and (cache / "mid-model").exists() and (cache / "new-model").exists()
    free_disk_space_for(10, cache_dir=str(cache), max_disk_space=10**9, os_quota=0)      # fits: nothing evicted
    assert (cache / "mid-model").exists()


def test_shield_and_wait_finishes_the_task_before_propagating_cancellation():
    from petals_b200.utils.asyncio import shield_and_wait

    async def scenario():
        done = []

        async def critical():
            await asyncio.sleep(0.05)
            done.append(True)
            return 7

        assert await shield_and_wait(critical()) == 7
        waiter = asyncio.create_task(shield_and_wait(critical()))
        await asyncio.sleep(0.01)
        waiter.cancel()
        with pytest.raises(asyncio.CancelledError):
            await waiter
        return done

    assert asyncio.run(scenario()) == [True, True]  # the shielded task ran to completion despite the cancellation


def test_reachability_spending_policy_aliases_and_constants():
    from petals_b200.client.routing.spending_policy import NoSpendingPolicy, SpendingPolicyBase
    from petals_b200.constants import DTYPE_MAP
    from petals_b200.server.reachability import check_direct_reachability, check_p2p_access
    from petals_b200.utils.hf_auth import always_needs_auth, resolve_token

    assert check_direct_reachability() is True
    assert isinstance(check_p2p_access(), dict)
    assert isinstance(NoSpendingPolicy(), SpendingPolicyBase) and NoSpendingPolicy().get_points("rpc_inference", 1, 2) == 0
    assert DTYPE_MAP["bfloat16"] is torch.bfloat16 and DTYPE_MAP["float16"] is torch.float16 and DTYPE_MAP["auto"] == "auto"
    assert not always_needs_auth("any/model") and resolve_token("hf_x") == "hf_x" and resolve_token(True) is None
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        import importlib

        import petals_b200.dht_utils as alias

        importlib.reload(alias)
        assert any(issubclass(x.category, DeprecationWarning) for x in w)
        assert hasattr(alias, "declare_active_modules") and hasattr(alias, "get_remote_module_infos")


def test_graphed_callable_falls_back_on_cpu():
    from petals_b200.utils.cuda_graphs import make_inference_graphed_callable

    fn = make_inference_graphed_callable(lambda a, b: a * 2 + b, (torch.ones(3), torch.zeros(3)))
    assert torch.equal(fn(torch.full((3,), 2.0), torch.ones(3)), torch.full((3,), 5.0))


def test_task_prioritizers_and_spending_policies():
    from petals_b200.client.routing.spending_policy import ConstantSpendingPolicy, NoSpendingPolicy, SpendingPolicyBase
    from petals_b200.server.task_prioritizer import DummyTaskPrioritizer, PointsTaskPrioritizer, TaskPrioritizerBase

    dummy = DummyTaskPrioritizer()
    assert dummy.prioritize(points=0.0, type="inference") == 1.0 < dummy.prioritize(points=5.0, type="forward") == 2.0
    assert dummy.prioritize(points=0.0, type="backward") == 2.0
    paid = PointsTaskPrioritizer()
    assert paid.prioritize(points=0, type="inference") == 1.0 and paid.prioritize(points=3, type="inference") < 1.0
    assert 1.0 < paid.prioritize(points=1e9, type="forward") < paid.prioritize(points=1, type="forward") < 2.0  # never overtakes inference
    with pytest.raises(NotImplementedError):
        TaskPrioritizerBase().prioritize(points=0.0)
    assert NoSpendingPolicy().get_points("rpc_inference") == 0.0 and ConstantSpendingPolicy(2).get_points("rpc_forward", 1, x=2) == 2.0
    with pytest.raises(ValueError):
        ConstantSpendingPolicy(-1)
    with pytest.raises(NotImplementedError):
        SpendingPolicyBase().get_points("rpc_forward")


@pytest.mark.parametrize("scaling", [
    dict(rope_type="linear", factor=4.0),
    dict(rope_type="llama3", factor=8.0, low_freq_factor=1.0, high_freq_factor=4.0, original_max_position_embeddings=512),
    dict(rope_type="yarn", factor=4.0, original_max_position_embeddings=512),
    dict(rope_type="yarn", factor=16.0, original_max_position_embeddings=256, beta_fast=16, beta_slow=2, mscale=1.0, mscale_all_dim=0.5),
    dict(rope_type="yarn", factor=2.0, original_max_position_embeddings=1024, attention_factor=1.25, truncate=False),
])
def test_rope_tables_match_the_library_definitions(scaling):
    """cos/sin tables of every supported rope_scaling type against transformers' own initialisers (used as an oracle only)."""
    from transformers import LlamaConfig
    from transformers.modeling_rope_utils import ROPE_INIT_FUNCTIONS

    from petals_b200.ops.functional import rope_tables

    head_dim, theta, max_pos = 64, 10000.0, 2048
    cfg = LlamaConfig(hidden_size=head_dim * 4, num_attention_heads=4, max_position_embeddings=max_pos,
                      rope_parameters=dict(scaling, rope_theta=theta))
    inv_freq, attention_factor = ROPE_INIT_FUNCTIONS[scaling["rope_type"]](cfg, "cpu")
    ang = torch.arange(max_pos, dtype=torch.float32)[:, None] * inv_freq[None, :].float()
    cos, sin = rope_tables(head_dim, max_pos, theta, scaling)
    assert cos.shape == (max_pos, head_dim // 2)
    assert torch.allclose(cos, ang.cos() * attention_factor, atol=2e-4) and torch.allclose(sin, ang.sin() * attention_factor, atol=2e-4)


def test_rope_tables_reject_length_dependent_scaling():
    from petals_b200.ops.functional import rope_tables

    for kind in ("dynamic", "longrope", "made-up"):
        with pytest.raises(NotImplementedError, match="rope_scaling"):
            rope_tables(64, 128, 10000.0, {"rope_type": kind, "factor": 2.0})
    plain = rope_tables(64, 128, 10000.0, None)
    same = rope_tables(64, 128, 10000.0, {"rope_type": "default"})
    assert torch.equal(plain[0], same[0]) and torch.equal(plain[1], same[1])


def test_safetensors_reader_refuses_inconsistent_files(tmp_path):
    """The header of a checkpoint file is untrusted input for the C++ reader: truncated / absurd / inconsistent records are errors,
    never crashes, over-long copies or half-initialised tensors."""
    import json
    import struct

    from petals_b200.utils.safetensors_io import SafetensorsFile, save_file

    good = tmp_path / "good.safetensors"
    save_file({"a": torch.arange(12, dtype=torch.float32).reshape(3, 4), "b": torch.ones(5, dtype=torch.bfloat16)}, str(good))
    raw = good.read_bytes()
    (hl,) = struct.unpack("<Q", raw[:8])
    body = raw[8 + hl:]

    def with_header(h):
        hb = json.dumps(h).encode()
        return struct.pack("<Q", len(hb)) + hb + body

    rec = lambda **kw: {"a": dict(dict(dtype="F32", shape=[3, 4], data_offsets=[0, 48]), **kw)}  # noqa: E731
    broken = {
        "empty": b"", "short": raw[:5], "truncated_header": raw[:8 + hl // 2], "truncated_body": raw[:-7],
        "huge_header_len": struct.pack("<Q", 1 << 60) + raw[8:], "garbage_header": struct.pack("<Q", 16) + b"\\xff" * 16 + body,
        "not_an_object": with_header([1, 2, 3]), "offsets_past_eof": with_header(rec(data_offsets=[0, 10 ** 9])),
        "negative_offsets": with_header(rec(data_offsets=[-8, 40])), "reversed_offsets": with_header(rec(data_offsets=[48, 0])),
        "shape_larger_than_data": with_header(rec(shape=[300, 400])), "shape_smaller_than_data": with_header(rec(shape=[2, 2])),
        "missing_fields": with_header({"a": {"dtype": "F32"}}), "huge_shape": with_header(rec(shape=[2 ** 40, 2 ** 40])),
        "deep_nesting": struct.pack("<Q", 20000) + b"[" * 20000 + body,
    }
    for name, data in broken.items():
        path = tmp_path / f"{name}.safetensors"
        path.write_bytes(data)
        with pytest.raises(IOError):
            SafetensorsFile(str(path))
    # a dtype this reader does not know makes only that tensor unavailable
    odd = tmp_path / "odd.safetensors"
    odd.write_bytes(with_header({"a": {"dtype": "F32", "shape": [3, 4], "data_offsets": [0, 48]}, "z": {"dtype": "Q7", "shape": [2], "data_offsets": [48, 58]}}))
    with SafetensorsFile(str(odd)) as f:
        assert torch.equal(f.get_tensor("a"), torch.arange(12, dtype=torch.float32).reshape(3, 4))
        with pytest.raises(IOError, match="unsupported dtype"):
            f.get_tensor("z")


def test_mxfp8_scale_block_layout_round_trip():
    """ops/quant.py:pack_scales lays the UE8M0 exponents out the way tcgen05.cp copies them into tensor memory (csrc/gemm_mxfp8.cu):
    [K/128][rows/128][512 B], scale of (row r, K slice c) at (r % 32) * 16 + (r // 32) * 4 + c; unpack_scales inverts it."""
    import torch

    from petals_b200.ops.quant import pack_scales, unpack_scales

    e = torch.randint(1, 255, (300, 8), dtype=torch.uint8)
    p = pack_scales(e)
    assert p.numel() == 2 * 3 * 512 and torch.equal(unpack_scales(p, 300, 256), e)
    for r, j in [(0, 0), (33, 1), (127, 7), (128, 4), (299, 3)]:
        off = ((j // 4) * 3 + r // 128) * 512 + (r % 128 % 32) * 16 + (r % 128 // 32) * 4 + j % 4
        assert p[off] == e[r, j]
    assert int(p.view(2, 3, 32, 4, 4)[:, 2, 12:, 1].sum()) == 0  # rows 300..383 are padding


def test_fabric_request_metadata_is_validated_before_it_indexes_device_memory(monkeypatch):
    """server/handler.py:fabric_endpoints — landing-slot coordinates come from the wire: sizes, ranks, slots and kinds are checked against
    the fabric of this process before any kernel sees them; without a fabric such a request is refused."""
    import types

    import pytest

    import petals_b200.parallel.fabric as fabric_mod
    from petals_b200.server.handler import fabric_endpoints

    assert fabric_endpoints({}) == (None, None)
    monkeypatch.setattr(fabric_mod, "_fabric", None)
    with pytest.raises(RuntimeError):
        fabric_endpoints({"fabric_in": {"B": 1, "T": 1, "src_rank": 0}})
    fab = types.SimpleNamespace(max_tokens=64, n_slots=4, world=4, rank=1, hidden_size=8)
    monkeypatch.setattr(fabric_mod, "_fabric", fab)
    take, push = fabric_endpoints({"fabric_in": {"B": 2, "T": 16, "src_rank": 0, "slot": 3}, "fabric_out": {"kind": "g_in", "rank": 2, "slot": 5}})
    assert take == (fab, 0, 2, 16, 3) and push == (fab, "g_in", 2, 1)  # the output slot wraps around the ring
    for bad in ({"B": 0, "T": 4, "src_rank": 0}, {"B": 2, "T": 64, "src_rank": 0}, {"B": 1, "T": 1, "src_rank": 4}, {"B": 1, "T": 1, "src_rank": -1},
                {"B": 1, "T": 1, "src_rank": 0, "slot": 4}, {"B": 1, "T": 1, "src_rank": 0, "slot": -1}):
        with pytest.raises(ValueError):
            fabric_endpoints({"fabric_in": bad})
    for bad in ({"kind": "weights", "rank": 0}, {"kind": "x_in", "rank": 4}, {"kind": "x_in", "rank": -1}):
        with pytest.raises(ValueError):
            fabric_endpoints({"fabric_out": bad})


def test_stage_activation_stash_is_bounded_and_single_use():
    """server/backend.py:Stage.stash_put / stash_pop — the span inputs a stage keeps between a fabric forward and its backward: at most
    STASH_ENTRIES micro-batches, expired entries dropped, an entry can be consumed once, unknown keys are an error (not silence)."""
    import collections
    import threading
    import time

    import pytest
    import torch

    from petals_b200.server.backend import Stage

    st = Stage.__new__(Stage)
    st._stash, st._stash_lock = collections.OrderedDict(), threading.Lock()
    for i in range(Stage.STASH_ENTRIES + 5):
        st.stash_put(f"k{i}", torch.full((1,), float(i)))
    assert len(st._stash) == Stage.STASH_ENTRIES and "k0" not in st._stash and "k5" in st._stash
    assert st.stash_pop("k7").item() == 7.0
    with pytest.raises(KeyError):
        st.stash_pop("k7")
    with pytest.raises(KeyError):
        st.stash_pop("k0")
    st._stash["old"] = (torch.zeros(1), time.monotonic() - Stage.STASH_TTL - 1)
    st._stash.move_to_end("old", last=False)
    st.stash_put("fresh", torch.zeros(1))
    assert "old" not in st._stash and "fresh" in st._stash


def test_failed_fabric_request_drains_its_landing_slot(monkeypatch):
    """server/handler.py: rpc_forward / rpc_backward that fail before the stage consumed the announced transfer (fault injection, unknown
    adapter, missing stash) still take + acknowledge the landing slot exactly once; requests that fail later do not drain twice."""
    import types

    import pytest
    import torch

    import petals_b200.parallel.fabric as fabric_mod
    import petals_b200.server.handler as handler_mod

    drained = []

    class FakeFabric:
        max_tokens, n_slots, world, rank, hidden_size = 64, 4, 2, 1, 8

        def recv(self, rows, kind, src, slot=0):
            drained.append((rows, kind, src, slot))
            return torch.zeros(rows, 8)

    fab = FakeFabric()
    monkeypatch.setattr(fabric_mod, "_fabric", fab)
    stage = types.SimpleNamespace(device=torch.device("cpu"), dtype=torch.float32, spec=types.SimpleNamespace(hidden_size=8), _stash={})
    h = types.SimpleNamespace(peer_id="s1", stage=stage, module_backends={}, _check_uids=lambda uids: (_ for _ in ()).throw(ValueError("unknown uid")),
                              check_adapter=lambda a: None, _stash_key=handler_mod.TransformerConnectionHandler._stash_key)
    meta = {"fabric_in": {"B": 2, "T": 3, "src_rank": 0, "slot": 2}}
    with pytest.raises(ValueError):
        handler_mod.TransformerConnectionHandler._rpc_forward(h, ["x"], torch.empty(0), None, dict(meta))
    assert drained == [(6, "x_in", 0, 2)]
    with pytest.raises(ValueError):
        handler_mod.TransformerConnectionHandler._rpc_backward(h, ["x"], torch.empty(0), torch.empty(0), None, dict(meta))
    assert drained[-1] == (6, "g_in", 0, 2) and len(drained) == 2
    # a backward whose stash is gone: refused before the stage runs, gradient slot drained
    h._check_uids = lambda uids: list(uids)
    h.module_backends = {"x": object()}
    with pytest.raises(KeyError):
        handler_mod.TransformerConnectionHandler._rpc_backward(h, ["x"], torch.empty(0), torch.empty(0), None, dict(meta, stash="gone"))
    assert len(drained) == 3
    # no fabric metadata: nothing to drain
    h._check_uids = lambda uids: (_ for _ in ()).throw(ValueError("unknown uid"))
    with pytest.raises(ValueError):
        handler_mod.TransformerConnectionHandler._rpc_forward(h, ["x"], torch.zeros(1, 1, 8), None, {})
    assert len(drained) == 3


def test_version_check_looks_at_the_peers_of_the_swarm():
    """utils/version.py: offline there is no package index to ask for updates, but the peers of the swarm announce their versions —
    a process warns (and returns the version) when peers serving the same blocks run a newer release; malformed versions sort lowest."""
    import time

    import petals_b200
    from petals_b200.data_structures import ServerInfo, ServerState, make_uid
    from petals_b200.parallel.swarm import Swarm
    from petals_b200.utils.dht import declare_active_modules
    from petals_b200.utils.version import newest_version, parse_version, validate_version

    assert parse_version("2.3.0.dev2") == (2, 3, 0) and parse_version("v10.1") == (10, 1) and parse_version("garbage") == () and parse_version(None) == ()
    assert newest_version(["1.9.9", None, "1.10.0", "x"]) == "1.10.0" and newest_version([]) is None
    swarm = Swarm("t-version")
    uids = [make_uid("m", i) for i in range(2)]
    mine = parse_version(petals_b200.__version__)
    newer = ".".join(str(p) for p in (mine[0] + 1,) + tuple(mine[1:])) if mine else "999.0"
    assert validate_version() is None and validate_version(swarm, uids) is None  # nobody else around
    info = ServerInfo(state=ServerState.ONLINE, throughput=1.0, version=petals_b200.__version__)
    declare_active_modules(swarm, uids, info, time.time() + 60, peer_id="same")
    assert validate_version(swarm, uids) is None
    declare_active_modules(swarm, uids[:1], ServerInfo(state=ServerState.ONLINE, throughput=1.0, version=newer), time.time() + 60, peer_id="newer")
    assert validate_version(swarm, uids) == newer


def test_stages_without_fused_hops_use_host_issued_fabric_copies():
    """server/backend.py: a tensor-parallel leader (``whole_span_only`` engine) has no epilogue that stores into a peer's landing slot, so
    Stage.inference_step / Stage.forward take the input with a fabric receive and deliver the output with a fabric send around the plain
    engine call — the protocol the fused engines implement inside their kernels."""
    import collections
    import threading

    import torch

    from petals_b200.server.backend import Stage

    log = []

    class FakeFabric:
        def recv(self, rows, kind, src, slot=0):
            log.append(("recv", rows, kind, src, slot))
            return torch.full((rows, 4), 2.0)

        def send(self, rows, rank, kind, slot=0):
            log.append(("send", tuple(rows.shape), rank, kind, slot, float(rows.sum())))

    class LeaderEngine:
        whole_span_only = True

        def inference_step(self, session, hidden, prompts, hypo_ids, block_range):
            return hidden + 1

        def forward(self, hidden, prompts, block_range):
            return hidden * 3

    st = Stage.__new__(Stage)
    st.engine, st.blocks, st.device, st.dtype, st.start_block, st.active_adapter = LeaderEngine(), [None, None], torch.device("cpu"), torch.float32, 0, None
    st._stash, st._stash_lock = collections.OrderedDict(), threading.Lock()
    fab = FakeFabric()
    out = st.inference_step(None, torch.empty(1, 2, 4), None, None, 0, 2, take_from=(fab, 0, 1, 2, 3), push_to=(fab, "x_in", 2, 1))
    assert out.shape == (1, 0, 4) and log == [("recv", 2, "x_in", 0, 3), ("send", (2, 4), 2, "x_in", 1, 24.0)]
    log.clear()
    out = st.forward(torch.empty(1, 2, 4), None, 0, 2, take_from=(fab, 1, 1, 2, 0), push_to=(fab, "y_ret", 0, 2), stash="k")
    assert out.shape == (1, 0, 4) and log == [("recv", 2, "x_in", 1, 0), ("send", (2, 4), 0, "y_ret", 2, 48.0)]
    assert torch.equal(st.stash_pop("k"), torch.full((1, 2, 4), 2.0))  # the stage kept its input for the backward
    log.clear()
    assert torch.equal(st.inference_step(None, torch.ones(1, 1, 4), None, None, 0, 2), torch.full((1, 1, 4), 2.0)) and not log  # no fabric: plain call
