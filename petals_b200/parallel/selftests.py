"""Numerics self-tests of the multi-GPU paths, as functions: run by ``tools/tp_selftest.py`` / ``tools/pp_selftest.py``, by
``tests/test_multi_gpu.py`` and — so that every multi-GPU benchmark run also proves the numbers come from a correct engine —
at the start of ``bench.py --gpus N`` (parallel/multi_gpu_bench.py), which fails the run on a mismatch."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def tp_selftest(dev: torch.device, preset: str = "llama-tiny") -> dict:
    """Collective over the default process group (>= 2 ranks, one GPU each). Every rank builds the SAME random tiny-Llama blocks, keeps
    its shard, and the group runs a multi-step session (prompt ingestion, single-token steps, a rollback, a long chunked prompt)
    through the public client API on rank 0; the result is compared with the oracle blocks evaluated on rank 0. Returns the report on
    rank 0 (``{"tp_selftest": "ok" | "FAILED", ...}``) and ``{}`` elsewhere. ``preset="mixtral-tiny"``: sparse-MoE blocks, every expert's
    FFN columns split over the ranks, router replicated (BASELINE config #4 is this layout on Mixtral-8x7B)."""
    from petals_b200.data_structures import ModelInfo, ServerInfo, ServerState
    from petals_b200.parallel.swarm import Swarm
    from petals_b200.parallel.symmetric import host_barrier, measure_hop_latency, measure_peer_bandwidth
    from petals_b200.parallel.tp_worker import TPLeaderEngine, build_tp_engine, follower_loop, make_ring
    from petals_b200.server.backend import Stage
    from petals_b200.server.server import ModuleContainer
    from petals_b200.utils.auto_config import AutoDistributedConfig
    from petals_b200.utils.random_model import random_blocks, random_client_model, write_config_only
    import petals_b200

    rank, world = dist.get_rank(), dist.get_world_size()
    overrides = dict(num_attention_heads=8, num_key_value_heads=max(2, world), num_hidden_layers=3)
    if os.environ.get("TP_SELFTEST_HIDDEN"):  # e.g. 2048: wide enough for the split-K decode path of small QKV shards
        h = int(os.environ["TP_SELFTEST_HIDDEN"])
        overrides.update(hidden_size=h, intermediate_size=2 * h, head_dim=128 if h >= 1024 else 64)
    path = write_config_only(preset, overrides)
    config = AutoDistributedConfig.from_pretrained(path)
    n = config.num_hidden_layers
    blocks = random_blocks(config, range(n), dev, seed=3)  # identical on every rank (seeded)
    engine, cache, heap = build_tp_engine(config, n, blocks=blocks, attn_cache_tokens=512, inference_max_length=256)
    ring = make_ring()
    probe = heap.alloc(8)
    bw = measure_peer_bandwidth(heap, 0, 1)
    lat = measure_hop_latency(heap, probe, 0, 1)
    if rank != 0:
        follower_loop(engine, cache, ring, rank - 1)
        host_barrier()
        heap.close()
        return {}
    swarm = Swarm("tp-selftest")
    leader = TPLeaderEngine(engine, ring)
    stage = Stage(config, blocks, 0, device=dev, memory_cache=cache, torch_dtype=torch.bfloat16, engine=leader)
    info = ServerInfo(state=ServerState.JOINING, throughput=1.0, version=petals_b200.__version__, torch_dtype="bfloat16", quant_type="none")
    container = ModuleContainer.from_stage(dht=swarm, dht_prefix=config.dht_prefix, block_config=config, stage=stage, server_info=info,
                                           model_info=ModelInfo(num_blocks=n, repository=path), peer_id="tp-leader", inference_max_length=256)
    try:
        model = random_client_model(path, swarm, dev)
        torch.manual_seed(0)
        ids = torch.randint(0, 4000, (1, 21), device=dev)
        with torch.inference_mode():
            h = model.model.embed(ids)
            for b in blocks:
                h = b.forward_cached(h, None, None, 0)
            ref = model.lm_head(model.model.final_norm(h)).float()
            with model.inference_session(max_length=64) as sess:
                a = model(ids[:, :13]).logits  # 13 rows -> one sequence-parallel prefill chunk with a ragged row split
                b_ = model(ids[:, 13:14]).logits
                junk = model(torch.randint(0, 4000, (1, 3), device=dev)).logits  # will be rolled back
                sess.position = 14
                c = model(ids[:, 14:15]).logits
                d = model(ids[:, 15:]).logits
            got = torch.cat([a, b_, c, d], 1).float()
            out = model.generate(ids[:, :8], max_new_tokens=6)
            # long prompt: 150 rows span two 128-row GEMM tiles and several KV pages; then two more chunks on top of the cache
            ids2 = torch.randint(0, 4000, (1, 200), device=dev)
            h = model.model.embed(ids2)
            for b in blocks:
                h = b.forward_cached(h, None, None, 0)
            ref2 = model.lm_head(model.model.final_norm(h)).float()
            with model.inference_session(max_length=256):
                got2 = torch.cat([model(ids2[:, :150]).logits, model(ids2[:, 150:199]).logits, model(ids2[:, 199:]).logits], 1).float()
        engine.check_errors()
        err = (got - ref).abs().mean().item() / (ref.abs().mean().item() + 1e-9)
        err2 = (got2 - ref2).abs().mean().item() / (ref2.abs().mean().item() + 1e-9)
        agree = (got.argmax(-1) == ref.argmax(-1)).float().mean().item()
        agree2 = (got2.argmax(-1) == ref2.argmax(-1)).float().mean().item()
    finally:  # the followers leave their command loops only when the leader says so
        leader.shutdown()
        container.shutdown()
        host_barrier()
        heap.close()
    # Bounds with head-room over what B200 boxes measured (profiles/): dense 0.9-1.0 % at 2-8 ranks; sparse MoE 0.8 % on decode steps and
    # 3.7-4.3 % on the chunked prompt (a routing decision that flips at a near-tie changes a whole row), arg-max agreement 92-97 % there
    moe = config.block_spec().mlp == "moe"
    ok = err < 0.05 and agree > 0.9 and err2 < (0.08 if moe else 0.05) and agree2 > (0.85 if moe else 0.9)
    return {"tp_selftest": "ok" if ok else "FAILED", "model": preset, "world": world, "rel_err": round(err, 5), "argmax_agreement": round(agree, 4),
                      "prefill_rel_err": round(err2, 5), "prefill_argmax_agreement": round(agree2, 4),
                      "generated": out[0, 8:].tolist(), "peer_store_GBps": bw, "flag_latency_us": lat}


def pp_selftest(dev: torch.device) -> dict:
    """Collective over the default process group. Every rank serves one pipeline stage of a tiny Llama; rank 0 also runs the client.
    Inference sessions (prefill, decode capture + replay, generate) and a training pass with and without deep prompts go through the
    public API; between stages activations and gradients travel through the NVLink landing rings (parallel/fabric.py). Compared with
    the oracle blocks (fp32 autograd for the training pass) on rank 0. Creates and closes its own fabric. Returns the report on rank 0
    (``{"pp_selftest": "ok" | "FAILED", ...}``) and ``{}`` elsewhere."""
    import tempfile

    import petals_b200.parallel.fabric as fabric_mod
    from petals_b200.parallel.fabric import init_fabric
    from petals_b200.parallel.swarm import FileSwarm
    from petals_b200.parallel.symmetric import host_barrier
    from petals_b200.utils.auto_config import AutoDistributedConfig
    from petals_b200.utils.random_model import launch_random_stage, random_blocks, random_client_model, write_config_only

    rank, world = dist.get_rank(), dist.get_world_size()
    n_layers = 2 * world
    path = write_config_only("llama-tiny", dict(num_hidden_layers=n_layers))
    config = AutoDistributedConfig.from_pretrained(path)
    fabric = init_fabric(config.hidden_size, max_tokens=1024)
    if fabric is None:
        raise RuntimeError("the pipeline self-test needs >= 2 ranks")
    dirs = [tempfile.mkdtemp(prefix="pb200-pp-") if rank == 0 else None]
    dist.broadcast_object_list(dirs, src=0)
    swarm = FileSwarm(dirs[0])
    per = n_layers // world
    stage = launch_random_stage(path, range(rank * per, (rank + 1) * per), swarm, dev, seed=5, peer_id=f"stage{rank}", attn_cache_tokens=1024,
                                inference_max_length=512)
    host_barrier()
    ok, report = True, {}
    if rank == 0:
        try:
            model = random_client_model(path, swarm, dev)
            blocks = random_blocks(config, range(n_layers), dev, seed=5)  # the same weights every stage drew (seeded per layer)
            torch.manual_seed(0)
            ids = torch.randint(0, 4000, (2, 40), device=dev)
            with torch.inference_mode():
                h = model.model.embed(ids)
                for b in blocks:
                    h = b.forward_cached(h, None, None, 0)
                ref = model.lm_head(model.model.final_norm(h)).float()
                with model.inference_session(max_length=64) as sess:
                    a = model(ids[:, :33]).logits  # prefill: tcgen05 GEMM epilogue pushes the tiles
                    b_ = model(ids[:, 33:34]).logits  # decode: GEMV epilogue pushes (graph capture)
                    c = model(ids[:, 34:35]).logits  # decode: graph replay
                    d = model(ids[:, 35:]).logits
                    used_fabric = [s.no_history for s in sess._server_sessions]
                    peers = [s.span.peer_id for s in sess._server_sessions]
                got = torch.cat([a, b_, c, d], 1).float()
                out = model.generate(ids[:1, :8], max_new_tokens=6)
            fabric.check_errors()
            err = (got - ref).abs().mean().item() / (ref.abs().mean().item() + 1e-9)
            agree = (got.argmax(-1) == ref.argmax(-1)).float().mean().item()
            ok = err < 0.05 and agree > 0.9 and all(used_fabric[1:]) and len(peers) == world
            report = {"pp_selftest": "ok" if ok else "FAILED", "world": world, "rel_err": round(err, 5), "argmax_agreement": round(agree, 4),
                      "stages": peers, "inputs_over_fabric": used_fabric, "generated": out[0, 8:].tolist()}
            # training over the fabric (BASELINE config #5): 3 micro-batches hop forward through the x_in rings (GEMM-epilogue pushes), the
            # gradients hop back through the g_in rings (stored by the last kernel of each stage's backward), every stage stashes its
            # input; compared with fp32 autograd through the oracle blocks, deep prompts included
            import copy

            import petals_b200.client.sequential_autograd as sa

            saved_mb, sa.MAX_TOKENS_IN_BATCH = sa.MAX_TOKENS_IN_BATCH, 2 * 48
            H = config.hidden_size
            blocks32 = [copy.deepcopy(b).float() for b in blocks]
            rel = lambda a, b: ((a.float() - b).abs().mean() / (b.abs().mean() + 1e-9)).item()
            t_err, hops = {}, {}
            for tag, use_prompts in (("prompts", True), ("plain", False)):  # without prompts the gradient hop is the fused one (last kernel stores to the peer)
                torch.manual_seed(1)
                x = (0.7 * torch.randn(6, 48, H, device=dev)).to(torch.bfloat16).requires_grad_(True)
                prompts = (0.1 * torch.randn(n_layers, 1, 4, H, device=dev)).to(torch.bfloat16).requires_grad_(True) if use_prompts else None
                before = dict(sa.FabricPlan.hops_done)
                y = model.model.layers(x, prompts=prompts)
                w = (0.1 * torch.randn_like(y)).float()
                (y.float() * w).sum().backward()
                hops[tag] = {k: sa.FabricPlan.hops_done[k] - before[k] for k in before}
                fabric.check_errors()
                x2 = x.detach().float().requires_grad_(True)
                p2 = prompts.detach().float().requires_grad_(True) if use_prompts else None
                h = x2
                for i, b32 in enumerate(blocks32):
                    if use_prompts:
                        h = torch.cat([h[:, :4] + p2[i], h[:, 4:]], 1)
                    h = b32.forward_cached(h, None, None, 0)
                (h * w).sum().backward()
                t_err[tag] = {"y": rel(y, h.detach()), "grad_x": rel(x.grad, x2.grad)}
                if use_prompts:
                    t_err[tag]["grad_prompts"] = rel(prompts.grad, p2.grad)
            # bf16 rounding accumulates with depth (2 blocks per stage): 0.7 % at 2 stages, 1.5 % at 8 on B200 boxes — 2 % per 8 blocks
            t_ok = (max(v for e in t_err.values() for v in e.values()) < 2e-2 * max(1.0, n_layers / 8)
                    and all(h_ == {"forward": 3 * world, "backward": 3 * world} for h_ in hops.values()))
            ok = ok and t_ok
            t_err = {k: {kk: round(vv, 5) for kk, vv in v.items()} for k, v in t_err.items()}
            sa.MAX_TOKENS_IN_BATCH = saved_mb
            report.update(pp_selftest="ok" if ok else "FAILED", training_rel_err=t_err, training_fabric_hops=hops)
        except Exception as e:  # noqa: BLE001 - the other ranks are waiting at the barrier below: report instead of raising here
            ok, report = False, {"pp_selftest": "FAILED", "world": world, "error": repr(e)[:300]}
    host_barrier()
    stage.shutdown()
    try:
        fabric.check_errors()
    except Exception as e:  # noqa: BLE001
        if rank == 0:
            report.update(pp_selftest="FAILED", error=repr(e)[:300])
    host_barrier()
    fabric.close()
    fabric_mod._fabric = None
    return report
