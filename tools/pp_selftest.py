"""Numerics self-test of the fused NVLink stage hop (run with torch.distributed.run on >= 2 GPUs).

Every rank serves one pipeline stage of a tiny Llama (a `Server`-like container registered in a rendezvous directory);
rank 0 also runs the client. Sessions go through the public API; between stages the activations are stored by the
producing stage's last kernel straight into the next stage's landing zone (parallel/fabric.py) and the control RPCs
carry no tensor bytes. The result is compared with the oracle blocks evaluated on rank 0."""
import json
import os
import sys
import tempfile

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from petals_b200.parallel.symmetric import host_barrier
    from petals_b200.parallel.fabric import init_fabric
    from petals_b200.parallel.swarm import FileSwarm
    from petals_b200.utils.auto_config import AutoDistributedConfig
    from petals_b200.utils.random_model import launch_random_stage, random_blocks, random_client_model, write_config_only

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group(backend="cpu:gloo,cuda:nccl", device_id=dev)
    n_layers = 2 * world
    path = write_config_only("llama-tiny", dict(num_hidden_layers=n_layers))
    config = AutoDistributedConfig.from_pretrained(path)
    fabric = init_fabric(config.hidden_size, max_tokens=1024)
    dirs = [tempfile.mkdtemp(prefix="pb200-pp-") if rank == 0 else None]
    dist.broadcast_object_list(dirs, src=0)
    swarm = FileSwarm(dirs[0])
    per = n_layers // world
    stage = launch_random_stage(path, range(rank * per, (rank + 1) * per), swarm, dev, seed=5, peer_id=f"stage{rank}", attn_cache_tokens=1024,
                                inference_max_length=512)
    host_barrier()
    ok, report = True, {}
    if rank == 0:
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

        sa.MAX_TOKENS_IN_BATCH = 2 * 48
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
        t_ok = (max(v for e in t_err.values() for v in e.values()) < 2e-2
                and all(h_ == {"forward": 3 * world, "backward": 3 * world} for h_ in hops.values()))
        ok = ok and t_ok
        t_err = {k: {kk: round(vv, 5) for kk, vv in v.items()} for k, v in t_err.items()}
        report.update(pp_selftest="ok" if ok else "FAILED", training_rel_err=t_err, training_fabric_hops=hops)
    host_barrier()
    stage.shutdown()
    fabric.check_errors()
    host_barrier()
    if rank == 0:
        print(json.dumps(report))
    fabric.close()
    dist.destroy_process_group()
    if not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
