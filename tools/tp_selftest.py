"""Numerics self-test of the tensor-parallel decode engine (run with torch.distributed.run on >= 2 GPUs).

Every rank builds the SAME random tiny-Llama blocks, keeps its shard, and the group runs a multi-step session
(prompt ingestion in 8-row micro-steps, single-token steps, a rollback) through the public client API on rank 0.
The result is compared with the oracle blocks evaluated on rank 0. Prints one JSON line on rank 0."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from petals_b200.parallel.symmetric import host_barrier
    from petals_b200.data_structures import ModelInfo, ServerInfo, ServerState
    from petals_b200.parallel.swarm import Swarm
    from petals_b200.parallel.symmetric import measure_hop_latency, measure_peer_bandwidth
    from petals_b200.parallel.tp_worker import TPLeaderEngine, build_tp_engine, follower_loop, make_ring
    from petals_b200.server.backend import Stage
    from petals_b200.server.server import ModuleContainer
    from petals_b200.utils.auto_config import AutoDistributedConfig
    from petals_b200.utils.random_model import random_blocks, random_client_model, write_config_only
    import petals_b200

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group(backend="cpu:gloo,cuda:nccl", device_id=dev)
    overrides = dict(num_attention_heads=8, num_key_value_heads=max(2, world), num_hidden_layers=3)
    if os.environ.get("TP_SELFTEST_HIDDEN"):  # e.g. 2048: wide enough for the split-K decode path of small QKV shards
        h = int(os.environ["TP_SELFTEST_HIDDEN"])
        overrides.update(hidden_size=h, intermediate_size=2 * h, head_dim=128 if h >= 1024 else 64)
    path = write_config_only("llama-tiny", overrides)
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
        dist.destroy_process_group()
        return
    swarm = Swarm("tp-selftest")
    leader = TPLeaderEngine(engine, ring)
    stage = Stage(config, blocks, 0, device=dev, memory_cache=cache, torch_dtype=torch.bfloat16, engine=leader)
    info = ServerInfo(state=ServerState.JOINING, throughput=1.0, version=petals_b200.__version__, torch_dtype="bfloat16", quant_type="none")
    container = ModuleContainer.from_stage(dht=swarm, dht_prefix=config.dht_prefix, block_config=config, stage=stage, server_info=info,
                                           model_info=ModelInfo(num_blocks=n, repository=path), peer_id="tp-leader", inference_max_length=256)
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
    leader.shutdown()
    container.shutdown()
    host_barrier()
    heap.close()
    ok = err < 0.05 and agree > 0.9 and err2 < 0.05 and agree2 > 0.9
    print(json.dumps({"tp_selftest": "ok" if ok else "FAILED", "world": world, "rel_err": round(err, 5), "argmax_agreement": round(agree, 4),
                      "prefill_rel_err": round(err2, 5), "prefill_argmax_agreement": round(agree2, 4),
                      "generated": out[0, 8:].tolist(), "peer_store_GBps": bw, "flag_latency_us": lat}))
    dist.destroy_process_group()
    if not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
