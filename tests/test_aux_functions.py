"""Unit tests of auxiliary pieces (reference tests/test_aux_functions.py, test_dtype.py): packaging round trip, compute
throughput measurement, dtype resolution, block size accounting, data structures, safetensors reader, block selection."""
import os

import pytest
import torch

from petals_b200.data_structures import ServerInfo, ServerState, RemoteModuleInfo, make_uid, parse_uid
from petals_b200.server.block_utils import get_block_size, resolve_block_dtype
from petals_b200.server.from_pretrained import load_pretrained_block
from petals_b200.server.throughput import measure_compute_rps
from petals_b200.utils.auto_config import AutoDistributedConfig
from petals_b200.utils.convert_block import QuantType
from petals_b200.utils.misc import DUMMY, is_dummy
from petals_b200.utils.packaging import pack_args_kwargs, unpack_args_kwargs
from tests.utils import checkpoint


def test_pack_unpack_roundtrip():
    x, y = torch.randn(3), torch.randn(2, 2)
    args = (x, [y, {"k": x, "n": 3}], "s", None)
    kwargs = dict(a=y, b=[1, 2, (x,)], c=DUMMY)
    flat, structure = pack_args_kwargs(*args, **kwargs)
    assert len(flat) == 6 and all(isinstance(t, torch.Tensor) for t in flat)
    import msgpack

    structure = msgpack.unpackb(msgpack.packb(structure), raw=False)  # survives the control channel encoding
    rargs, rkwargs = unpack_args_kwargs(flat, structure)
    assert torch.equal(rargs[0], x) and torch.equal(rargs[1][0], y) and rargs[1][1]["n"] == 3 and rargs[2] == "s" and rargs[3] is None
    assert torch.equal(rkwargs["a"], y) and torch.equal(rkwargs["b"][2][0], x) and is_dummy(rkwargs["c"])


def test_uids_and_server_info():
    assert parse_uid(make_uid("llama-hf", 17)) == ("llama-hf", 17)
    with pytest.raises(ValueError):
        parse_uid("a.1 a.2")
    info = ServerInfo(state=ServerState.ONLINE, throughput=12.5, start_block=1, end_block=4, adapters=("x",), next_pings={"p": 0.1})
    assert ServerInfo.from_tuple(info.to_tuple()) == info
    fwd = ServerInfo.from_tuple((2, 1.0, {"unknown_future_field": 1, "version": "9"}))  # forward compatible
    assert fwd.version == "9" and fwd.state == ServerState.ONLINE


@pytest.mark.parametrize("inference", [True, False])
def test_measure_compute_rps(inference):
    config = AutoDistributedConfig.from_pretrained(checkpoint("llama"))
    rps = measure_compute_rps(config, torch.device("cpu"), torch.float32, n_tokens=2, n_steps=2, inference=inference)
    assert isinstance(rps, float) and rps > 0


def test_dtype_resolution_and_block_loading():
    path = checkpoint("llama")
    config = AutoDistributedConfig.from_pretrained(path)
    assert resolve_block_dtype(config, torch.float16) == torch.float16
    assert resolve_block_dtype(config, "auto") == torch.bfloat16  # fp32 checkpoints are served in bf16 by default
    for dtype in (torch.float32, torch.bfloat16):
        block = load_pretrained_block(path, 0, torch_dtype=dtype)
        assert all(p.dtype == dtype for p in block.parameters())
        assert all(not p.requires_grad for p in block.parameters())
    with pytest.raises(KeyError):
        load_pretrained_block(path, 99, torch_dtype=torch.float32)
    n = config.block_spec().num_params()
    assert abs(get_block_size(config, "memory", dtype=torch.bfloat16) - 2 * n * 1.01) < 8
    assert get_block_size(config, "memory", quant_type=QuantType.FP8) < get_block_size(config, "memory", dtype=torch.bfloat16) * 0.6


def test_native_safetensors_reader_matches_python_package(tmp_path):
    from petals_b200.utils.safetensors_io import SafetensorsFile, save_file

    tensors = {"a.weight": torch.randn(5, 7), "b": torch.arange(10, dtype=torch.int64), "c.bf16": torch.randn(3, 3).to(torch.bfloat16),
               "empty": torch.zeros(0, 4)}
    p = str(tmp_path / "t.safetensors")
    save_file(tensors, p, metadata={"format": "pt"})
    with SafetensorsFile(p) as f:
        assert sorted(f.keys()) == sorted(tensors)
        for k, t in tensors.items():
            assert f.info(k) == (t.dtype, tuple(t.shape))
            assert torch.equal(f.get_tensor(k), t)
    st = pytest.importorskip("safetensors.torch")
    theirs = st.load_file(p)  # files we write are readable by the reference's loader ...
    assert all(torch.equal(theirs[k], tensors[k]) for k in tensors)
    p2 = str(tmp_path / "t2.safetensors")
    st.save_file({k: v for k, v in tensors.items()}, p2)  # ... and vice versa
    with SafetensorsFile(p2) as f:
        assert all(torch.equal(f.get_tensor(k), tensors[k]) for k in tensors)
    with pytest.raises(IOError):
        SafetensorsFile(str(tmp_path / "missing.safetensors"))


def test_block_selection_policy():
    from petals_b200.server.block_selection import choose_best_blocks, should_choose_other_blocks

    def infos(assign):  # assign: {peer: (start, end, throughput)}
        out = []
        for i in range(8):
            servers = {p: ServerInfo(state=ServerState.ONLINE, throughput=t, start_block=s, end_block=e) for p, (s, e, t) in assign.items() if s <= i < e}
            out.append(RemoteModuleInfo(uid=f"m.{i}", servers=servers))
        return out

    # blocks 4..7 are unserved -> a new 4-block server should take them
    assert choose_best_blocks(4, infos({"a": (0, 4, 1.0)})) == [4, 5, 6, 7]
    # two servers on the same half, nobody on the other: the swarm is badly balanced -> move
    assert should_choose_other_blocks("b", infos({"a": (0, 4, 1.0), "b": (0, 4, 1.0), "c": (4, 8, 0.1)}), balance_quality=0.75)
    # evenly covered -> stay
    assert not should_choose_other_blocks("b", infos({"a": (0, 4, 1.0), "b": (4, 8, 1.0)}), balance_quality=0.75)
    assert should_choose_other_blocks("b", infos({"a": (0, 4, 1.0), "b": (4, 8, 1.0)}), balance_quality=1.5)  # forced


def test_fault_plan_parsing_and_counters():
    from petals_b200.utils import fault_injection as fi

    try:
        fi.set_fault_plan("rpc=rpc_forward,peer=a,after=1,times=2,error=boom; rpc=rpc_info")
        fi.maybe_fail("rpc_forward", "b")          # other peer: rule does not apply
        fi.maybe_fail("rpc_forward", "a")          # call 1 <= after
        for _ in range(2):                         # calls 2 and 3 fire
            with pytest.raises(fi.InjectedFault, match="boom"):
                fi.maybe_fail("rpc_forward", "a")
        fi.maybe_fail("rpc_forward", "a")          # exhausted
        with pytest.raises(fi.InjectedFault):
            fi.maybe_fail("rpc_info", "whoever")   # peer-less rule matches any peer
        assert fi.fired_count() == 3
        with pytest.raises(ValueError):
            fi.set_fault_plan("peer=a,after=1")
    finally:
        fi.set_fault_plan(None)
    fi.maybe_fail("rpc_forward", "a")              # no plan: no-op


def test_step_timer_and_nvtx_noop():
    from petals_b200.utils.tracing import StepTimer, nvtx_range

    t = StepTimer()
    with t.measure("x"), nvtx_range("cpu-range-is-a-no-op"):
        pass
    with t.measure("x"):
        pass
    snap = t.snapshot()
    assert snap["x"]["count"] == 2 and snap["x"]["seconds"] >= 0 and t.snapshot() == {}


def test_importing_the_package_pulls_in_no_heavy_optional_dependency():
    """reference tests/test_aux_functions.py:16-29 checks that bitsandbytes is not imported by `import petals` (slow, noisy, optional).
    The equivalents here: Hugging Face transformers (only the tokenizer of the HTTP front-end and test oracles use it), triton and the
    library kernels in the image (flash_attn / flashinfer / vllm), pydantic — none is needed to serve or to run a client."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys, petals, petals.client, petals.server.server, petals.utils.peft\n"
            "heavy = [m for m in ('transformers', 'triton', 'flash_attn', 'flashinfer', 'vllm', 'pydantic', 'scipy', 'pandas') if m in sys.modules]\n"
            "print(','.join(heavy))")
    out = subprocess.run([sys.executable, "-c", code], cwd=root, env=dict(os.environ, PYTHONPATH=root), capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.strip() == "", f"imported on `import petals`: {out.stdout.strip()}"


def test_checkpoints_that_disagree_with_their_config_are_refused(tmp_path):
    """Kernel launches are sized from config.json (BlockSpec, vocab, hidden size). A checkpoint whose tensors have other shapes must fail
    at load time with a readable message — on a GPU it would otherwise be an out-of-bounds access."""
    import json
    import os
    import shutil

    from petals_b200.parallel.swarm import Swarm
    from petals_b200.server.from_pretrained import load_pretrained_block
    from petals_b200.utils.auto_config import AutoDistributedModelForCausalLM
    from tests.utils import checkpoint

    def variant(**overrides):
        d = str(tmp_path / ("v" + "-".join(overrides)))
        shutil.copytree(checkpoint("llama"), d)
        cfg_path = os.path.join(d, "config.json")
        cfg = json.load(open(cfg_path))
        cfg.update({k: (cfg[k] * 2 if v == "double" else v) for k, v in overrides.items()})
        json.dump(cfg, open(cfg_path, "w"))
        return d

    with pytest.raises(RuntimeError, match="do not match config.json"):
        load_pretrained_block(variant(intermediate_size="double"), 1, torch_dtype=torch.float32)
    with pytest.raises(RuntimeError, match="do not match config.json"):
        load_pretrained_block(variant(num_key_value_heads=4), 0, torch_dtype=torch.float32)
    with pytest.raises(ValueError, match="config.json implies"):
        AutoDistributedModelForCausalLM.from_pretrained(variant(vocab_size="double"), initial_peers=Swarm("shape-check"))
    for index in (99, -1):
        with pytest.raises(KeyError, match="no tensors for block"):
            load_pretrained_block(checkpoint("llama"), index, torch_dtype=torch.float32)


def test_throughput_cache_keys_and_corrupt_cache(tmp_path, monkeypatch):
    """`--throughput auto`: measured once per (model, device, dtype, quantisation, TP width) and cached on disk; an unreadable cache
    file is re-measured, not fatal (reference server/throughput.py:56-92)."""
    from petals_b200.server import throughput as tp_mod

    config = AutoDistributedConfig.from_pretrained(checkpoint("llama"))
    calls = []

    def fake_measure(cfg, device, dtype, *, quant_type, tensor_parallel_devices=()):
        calls.append(len(tensor_parallel_devices))
        return dict(inference_rps=100.0, forward_rps=1000.0, network_rps=1e6)

    monkeypatch.setattr(tp_mod, "measure_throughput_info", fake_measure)
    kw = dict(num_blocks=3, cache_dir=str(tmp_path))
    a = tp_mod.get_server_throughput("m", config, torch.device("cpu"), torch.float32, **kw)
    b = tp_mod.get_server_throughput("m", config, torch.device("cpu"), torch.float32, **kw)  # served from the cache
    assert a == b and calls == [0] and a["throughput"] == pytest.approx(1000.0 / 2)  # (3 + 1) / 2 blocks of compute per request
    tp_mod.get_server_throughput("m", config, torch.device("cpu"), torch.float32, tensor_parallel_devices=(torch.device("cpu"),) * 2, **kw)
    tp_mod.get_server_throughput("m", config, torch.device("cpu"), torch.bfloat16, **kw)
    assert calls == [0, 2, 0]  # other TP width / dtype = other cache entries
    (tmp_path / "throughput_v1.json").write_text("{ not json")
    tp_mod.get_server_throughput("m", config, torch.device("cpu"), torch.float32, **kw)
    assert calls == [0, 2, 0, 0]
    tp_mod.get_server_throughput("m", config, torch.device("cpu"), torch.float32, force_eval=True, **kw)
    assert len(calls) == 5
