"""CPU twin of tools/pp_selftest.py (gloo + HostFabric): checks the fabric *protocol* end to end — landing zones, flags,
acknowledgements, tensor-less control RPCs, full-chain replay — with oracle executors. Launched by tests/test_fabric_cpu.py."""
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
    from petals_b200.server.from_pretrained import load_pretrained_block
    from petals_b200.server.server import Server
    from petals_b200.utils.auto_config import AutoDistributedConfig, AutoDistributedModelForCausalLM

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group(backend="gloo")
    path = sys.argv[1]
    config = AutoDistributedConfig.from_pretrained(path)
    n = config.num_hidden_layers
    fabric = init_fabric(config.hidden_size, max_tokens=256)
    dirs = [tempfile.mkdtemp(prefix="pb200-ppcpu-") if rank == 0 else None]
    dist.broadcast_object_list(dirs, src=0)
    per = n // world
    server = Server(initial_peers=dirs[0], converted_model_name_or_path=path, block_indices=f"{rank * per}:{(rank + 1) * per}", torch_dtype="float32",
                    device="cpu", throughput=1.0, update_period=0.5, peer_id=f"stage{rank}")
    server.run_in_background(timeout=120)
    host_barrier()
    ok, report = True, {}
    if rank == 0:
        model = AutoDistributedModelForCausalLM.from_pretrained(path, initial_peers=[dirs[0]], max_retries=20, min_backoff=0.2, max_backoff=0.5)
        torch.manual_seed(0)
        ids = torch.randint(0, config.vocab_size, (2, 12))
        with torch.inference_mode():
            h = model.model.embed(ids)
            for i in range(n):
                h = load_pretrained_block(path, i, torch_dtype=torch.float32)(h)[0]
            ref = model.lm_head(model.model.final_norm(h))
            with model.inference_session(max_length=16) as sess:
                a = model(ids[:, :7]).logits
                b_ = model(ids[:, 7:8]).logits
                junk = model(ids[:, :2]).logits
                sess.position = 8  # rollback travels as start_from_position next to the fabric metadata
                c = model(ids[:, 8:]).logits
                used = [s.no_history for s in sess._server_sessions]
                peers = [s.span.peer_id for s in sess._server_sessions]
            got = torch.cat([a, b_, c], 1)
        err = (got - ref).abs().max().item()
        parts = [(a - ref[:, :7]).abs().max().item(), (b_ - ref[:, 7:8]).abs().max().item(), (c - ref[:, 8:]).abs().max().item()]
        ok = err < 1e-3 and all(used[1:]) and len(peers) == world
        report = {"pp_selftest_cpu": "ok" if ok else "FAILED", "max_err": err, "stages": peers, "inputs_over_fabric": used, "part_errs": parts}
        # training over the fabric: micro-batches hop forward through x_in rings, gradients hop back through g_in rings, every stage
        # stashes its input; compared with local autograd through the same blocks (deep prompts included)
        from petals_b200.client.sequential_autograd import FabricPlan
        import petals_b200.client.sequential_autograd as sa

        sa.MAX_TOKENS_IN_BATCH = 24  # 3 micro-batches of 2 sequences
        torch.manual_seed(1)
        x = torch.randn(6, 12, config.hidden_size, requires_grad=True)
        prompts = (0.1 * torch.randn(n, 1, 3, config.hidden_size)).requires_grad_(True)
        before = dict(FabricPlan.hops_done)
        y = model.model.layers(x, prompts=prompts)
        w = torch.randn_like(y)
        (y * w).sum().backward()
        hops = {k: FabricPlan.hops_done[k] - before[k] for k in before}
        x2, p2 = x.detach().clone().requires_grad_(True), prompts.detach().clone().requires_grad_(True)
        h = x2
        for i in range(n):
            blk = load_pretrained_block(path, i, torch_dtype=torch.float32)
            h = torch.cat([h[:, :3] + p2[i], h[:, 3:]], 1)
            h = blk(h)[0]
        (h * w).sum().backward()
        t_err = {"y": (y - h).abs().max().item(), "grad_x": (x.grad - x2.grad).abs().max().item(), "grad_prompts": (prompts.grad - p2.grad).abs().max().item()}
        t_ok = max(t_err.values()) < 1e-3 and hops == {"forward": 3 * world, "backward": 3 * world}
        ok = ok and t_ok
        report.update(pp_selftest_cpu="ok" if ok else "FAILED", training_err=t_err, training_fabric_hops=hops)
        if os.environ.get("PP_SELFTEST_FAULT") == "1":
            # PETALS_B200_FAULTS makes the LAST stage refuse its 4th rpc_forward (= micro-batch 0 of the next pass) after the previous stage
            # has already pushed into its landing slot: the stage must drain that slot, the client must finish the micro-batch on the
            # tensor-carrying path, pause fabric passes, and the rings must still be in step afterwards
            import time

            def one_pass():
                xa, pa = x.detach().clone().requires_grad_(True), prompts.detach().clone().requires_grad_(True)
                before = dict(FabricPlan.hops_done)
                t0 = time.monotonic()
                ya = model.model.layers(xa, prompts=pa)
                (ya * w).sum().backward()
                err = max((ya - h).abs().max().item(), (xa.grad - x2.grad).abs().max().item(), (pa.grad - p2.grad).abs().max().item())
                return err, {k: FabricPlan.hops_done[k] - before[k] for k in before}, time.monotonic() - t0

            manager = model.model.layers.sequence_manager
            faulted = one_pass()      # fault injected at the last stage: one micro-batch redone with tensors
            cooling = one_pass()      # inside the cool-down: no fabric hops at all
            manager.fabric_broken_until = 0.0
            healed = one_pass()       # fabric again: only works if the refused transfer was drained
            f_ok = (max(faulted[0], cooling[0], healed[0]) < 1e-3 and faulted[1]["forward"] < 3 * world and cooling[1] == {"forward": 0, "backward": 0}
                    and healed[1] == {"forward": 3 * world, "backward": 3 * world} and healed[2] < 20)
            ok = ok and f_ok
            report.update(pp_selftest_cpu="ok" if ok else "FAILED",
                          fault_passes={"faulted": faulted[1], "cooling": cooling[1], "healed": healed[1], "max_err": max(faulted[0], cooling[0], healed[0]), "healed_s": round(healed[2], 2)})
    host_barrier()
    server.shutdown()
    host_barrier()
    if rank == 0:
        print(json.dumps(report))
    fabric.close()
    dist.destroy_process_group()
    if not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
