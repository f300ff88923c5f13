"""``bench.py --impl reference``: the UNMODIFIED reference (installed under ``baseline/_ref``) measured on the same metric/config.

Nothing from ``petals_b200`` is imported here. The arm

1. imports ``hivemind`` and the reference's ``petals`` package from ``baseline/_ref`` — if either is missing the arm reports
   ``{"impl": "reference", "unavailable": "<why>"}`` (the offline image has no ``hivemind`` wheel: DESIGN.md §6);
2. writes a random-init Hugging Face checkpoint of the named architecture (``config.json`` + one safetensors shard per block +
   embeddings/head), because there is no network to download one;
3. starts the reference's own swarm as subprocesses — ``python -m petals.cli.run_dht`` and one ``python -m petals.cli.run_server``
   per GPU with an equal share of the blocks (the reference's only multi-GPU layout: pipeline stages joined by libp2p);
4. runs the loop of the reference's ``benchmarks/benchmark_inference.py:44-68`` (one ``inference_session``,
   ``generate(max_new_tokens=1, session=sess)`` per step) for W warm-up + K timed steps and prints the same JSON line as ours.
   The reference's client holds tokens on the host and its stages are reached through RPCs, so its only timing is end to end:
   ``value`` and ``e2e.value`` are the same host-clock number.
"""
from __future__ import annotations

import json
import os
import re
import subprocess
import sys
import tempfile
import time
from typing import List, Optional

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")

ARCH = {  # public HF configs of the benchmarked architectures
    "llama-3-70b": dict(architectures=["LlamaForCausalLM"], model_type="llama", vocab_size=128256, hidden_size=8192, intermediate_size=28672,
                        num_hidden_layers=80, num_attention_heads=64, num_key_value_heads=8, max_position_embeddings=8192, rms_norm_eps=1e-5,
                        rope_theta=500000.0, hidden_act="silu", tie_word_embeddings=False, torch_dtype="bfloat16", bos_token_id=128000, eos_token_id=128001),
    "llama-3-8b": dict(architectures=["LlamaForCausalLM"], model_type="llama", vocab_size=128256, hidden_size=4096, intermediate_size=14336,
                       num_hidden_layers=32, num_attention_heads=32, num_key_value_heads=8, max_position_embeddings=8192, rms_norm_eps=1e-5,
                       rope_theta=500000.0, hidden_act="silu", tie_word_embeddings=False, torch_dtype="bfloat16", bos_token_id=128000, eos_token_id=128001),
}


def _why_unavailable() -> Optional[str]:
    if not os.path.isdir(os.path.join(REF, "petals")):
        return "reference not installed under baseline/_ref (pip --no-index install fails: hivemind/tensor_parallel/bitsandbytes wheels absent offline)"
    sys.path.insert(0, REF)
    os.environ.setdefault("PETALS_IGNORE_DEPENDENCY_VERSION", "1")
    try:
        import importlib

        importlib.import_module("hivemind")
        importlib.import_module("petals")
    except Exception as e:  # noqa: BLE001
        return f"reference import fails offline: {type(e).__name__}: {str(e)[:140]}"
    return None


def write_random_checkpoint(name: str, root: str) -> str:
    """HF-layout checkpoint with random bf16 weights: what the reference's per-block loader (server/from_pretrained.py) reads."""
    import torch
    from safetensors.torch import save_file

    cfg = ARCH[name]
    path = os.path.join(root, name)
    if os.path.exists(os.path.join(path, "model.safetensors.index.json")):
        return path
    os.makedirs(path, exist_ok=True)
    with open(os.path.join(path, "config.json"), "w") as f:
        json.dump(cfg, f)
    H, I, V = cfg["hidden_size"], cfg["intermediate_size"], cfg["vocab_size"]
    D = H // cfg["num_attention_heads"]
    kv = cfg["num_key_value_heads"] * D
    index = {}
    g = torch.Generator().manual_seed(0)

    def rnd(*shape):
        return (torch.randn(*shape, generator=g) * 0.02).to(torch.bfloat16)

    for i in range(cfg["num_hidden_layers"]):
        p = f"model.layers.{i}."
        tensors = {p + "self_attn.q_proj.weight": rnd(H, H), p + "self_attn.k_proj.weight": rnd(kv, H), p + "self_attn.v_proj.weight": rnd(kv, H),
                   p + "self_attn.o_proj.weight": rnd(H, H), p + "mlp.gate_proj.weight": rnd(I, H), p + "mlp.up_proj.weight": rnd(I, H),
                   p + "mlp.down_proj.weight": rnd(H, I), p + "input_layernorm.weight": torch.ones(H, dtype=torch.bfloat16),
                   p + "post_attention_layernorm.weight": torch.ones(H, dtype=torch.bfloat16)}
        fn = f"model-{i + 1:05d}.safetensors"
        save_file(tensors, os.path.join(path, fn))
        index.update({k: fn for k in tensors})
    shell = {"model.embed_tokens.weight": rnd(V, H), "model.norm.weight": torch.ones(H, dtype=torch.bfloat16), "lm_head.weight": rnd(V, H)}
    save_file(shell, os.path.join(path, "model-shell.safetensors"))
    index.update({k: "model-shell.safetensors" for k in shell})
    with open(os.path.join(path, "model.safetensors.index.json"), "w") as f:
        json.dump({"metadata": {}, "weight_map": index}, f)
    return path


def _spawn(cmd: List[str], log: str) -> subprocess.Popen:
    env = dict(os.environ, PYTHONPATH=REF + os.pathsep + os.environ.get("PYTHONPATH", ""), PETALS_IGNORE_DEPENDENCY_VERSION="1")
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    return subprocess.Popen(cmd, stdout=open(log, "w"), stderr=subprocess.STDOUT, env=env, start_new_session=True)


def _wait_for(log: str, pattern: str, timeout: float) -> str:
    deadline = time.time() + timeout
    while time.time() < deadline:
        if os.path.exists(log):
            m = re.search(pattern, open(log, errors="replace").read())
            if m:
                return m.group(1) if m.groups() else m.group(0)
        time.sleep(1.0)
    raise TimeoutError(f"{pattern!r} never appeared in {log}")


def run_reference(args) -> dict:
    why = _why_unavailable()
    if why is not None:
        return {"impl": "reference", "unavailable": why}
    if int(os.environ.get("RANK", "0")) != 0:
        return {"impl": "reference", "rank": int(os.environ["RANK"]), "note": "rank 0 drives the reference swarm"}
    if args.model not in ARCH:
        return {"impl": "reference", "unavailable": f"no random checkpoint recipe for {args.model}"}
    import torch

    n_gpus = max(1, args.gpus)
    root = tempfile.mkdtemp(prefix="petals-ref-")
    procs: List[subprocess.Popen] = []
    try:
        ckpt = write_random_checkpoint(args.model, os.environ.get("PETALS_REF_CKPT_DIR", "/tmp/petals-ref-ckpt"))
        n_blocks = ARCH[args.model]["num_hidden_layers"]
        dht_log = os.path.join(root, "dht.log")
        procs.append(_spawn([sys.executable, "-m", "petals.cli.run_dht", "--host_maddrs", "/ip4/127.0.0.1/tcp/31337",
                             "--identity_path", os.path.join(root, "dht.id")], dht_log))
        peer = _wait_for(dht_log, r"(/ip4/127\.0\.0\.1/tcp/31337/p2p/\w+)", 120)
        bounds = [round(i * n_blocks / n_gpus) for i in range(n_gpus + 1)]
        for g in range(n_gpus):
            log = os.path.join(root, f"server{g}.log")
            procs.append(_spawn([sys.executable, "-m", "petals.cli.run_server", ckpt, "--initial_peers", peer, "--block_indices",
                                 f"{bounds[g]}:{bounds[g + 1]}", "--torch_dtype", "bfloat16", "--device", f"cuda:{g}", "--throughput", "1",
                                 "--attn_cache_tokens", str(args.seq_len + 256), "--new_swarm"], log))
        for g in range(n_gpus):
            _wait_for(os.path.join(root, f"server{g}.log"), r"Started", 3600)
        from petals import AutoDistributedModelForCausalLM  # the reference's public API

        model = AutoDistributedModelForCausalLM.from_pretrained(ckpt, initial_peers=[peer], torch_dtype=torch.bfloat16)
        K, W = args.steps, max(args.warmup, 3)
        prompt = torch.randint(0, ARCH[args.model]["vocab_size"], (1, args.prompt_len))
        with model.transformer.h.inference_session(max_length=args.seq_len) as sess:
            model.generate(prompt, max_new_tokens=1, session=sess)
            for _ in range(W):
                model.generate(max_new_tokens=1, session=sess)
            t0 = time.perf_counter()
            for _ in range(K):
                out = model.generate(max_new_tokens=1, session=sess)
                int(out[0, -1])  # the result is read on the host, like the reference benchmark's tokenizer.decode
            dt = time.perf_counter() - t0
        value = K / dt
        return {"impl": "reference", "metric": f"{args.model} single-stream decode tokens/s (reference swarm on this box, host-timed end to end)",
                "value": round(value, 3), "unit": "tokens/s", "n_gpus": n_gpus, "steps": K, "warmup": W, "ms_per_step": round(1e3 * dt / K, 3),
                "higher_is_better": True, "scaling": "strong", "vs_baseline": round(value / 6.0, 3), "dtype": "bf16",
                "data": "synthetic token ids; random-init weights of the named architecture",
                "config": {"model": args.model, "global_batch": 1, "seq_len": args.seq_len, "parallelism": f"pp{n_gpus} (reference servers, libp2p hops)"},
                "e2e": {"value": round(value, 3), "unit": "tokens/s", "h2d_bytes_per_step": 8, "d2h_bytes_per_step": 8}}
    except Exception as e:  # noqa: BLE001 - the driver needs a line, not a traceback
        return {"impl": "reference", "unavailable": f"reference swarm failed to run here: {type(e).__name__}: {str(e)[:160]}"}
    finally:
        for p in procs:
            try:
                os.killpg(p.pid, 15)  # exactly the process groups started above
            except Exception:  # noqa: BLE001
                pass
