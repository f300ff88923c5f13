"""bench.py back-end for N > 1 GPUs (one process per GPU, launched by torch.distributed.run).

Single-stream decode is bound by streaming the weights once per token, so the configuration that scales it is tensor
parallelism over NVLink: every GPU streams 1/N of every block and the two all-reduces per block are fused into the
GEMV epilogues/prologues (parallel/tensor_parallel.py). Rank 0 hosts the client model shell and the TP leader; the
timed loops go through the same public client API as the 1-GPU run (``model(input_ids=...)`` inside an inference
session). Device time is measured on every rank with CUDA events around the same K steps; the reported time is the
maximum over ranks."""
from __future__ import annotations

import json
import os
import time

import torch
import torch.distributed as dist

from petals_b200.parallel.symmetric import host_barrier

from petals_b200.utils.bench_common import (BASELINE_TOKENS_PER_S, ClockSampler, device_timed_decode, e2e_decode, metric_name,
                                            prime_session)


def run_selftests(args, dev, which=("tp", "pp")) -> dict:
    """The numerics self-tests of the multi-GPU paths (parallel/selftests.py: tiny model, public client API, compared with the oracle
    blocks / fp32 autograd), run inside the benchmark job so that every multi-GPU number comes with proof that the engine producing it
    computes the right thing on this box. Collective; the reports land on rank 0. ``--skip-selftests`` turns them off."""
    if getattr(args, "skip_selftests", False):
        return {"skipped": True}
    from petals_b200.parallel import selftests

    out = {}
    for name in which:
        try:
            if name == "tp" and str(getattr(args, "model", "")).startswith("mixtral"):
                out[name] = selftests.tp_selftest(dev, "mixtral-tiny")
                continue
            out[name] = getattr(selftests, f"{name}_selftest")(dev)
        except Exception as e:  # noqa: BLE001 - reported (and fatal for the run) on rank 0
            out[name] = {f"{name}_selftest": "FAILED", "error": repr(e)[:300]}
    return out


def _selftests_ok(reports: dict) -> bool:
    return bool(reports.get("skipped")) or all(r.get(f"{k}_selftest") == "ok" for k, r in reports.items())


def run_pp_tp(args, n_stages: int, tp: int) -> None:
    """``bench.py --parallelism pp<S>xtp<T>`` (S * T = N GPUs): a pipeline of S stages, each stage a tensor-parallel group of T GPUs —
    BASELINE.json config #4 names this layout for Mixtral-8x7B (4 stages x TP 2). Ranks [gT, (g+1)T) form group g and serve blocks
    [bounds[g], bounds[g+1]); the group's first rank is its leader (stage process: handler, KV bookkeeping, command ring to the
    followers), the client runs on rank 0. Inside a group everything is the tensor-parallel engine (NVLink LL all-reduces, sequence-
    parallel prefill); between stages the hidden states hop through an NVLink fabric the LEADERS share (landing rings, one peer copy +
    flag per hop; PETALS_B200_PPTP_FABRIC=0: with the stage-to-stage RPCs instead)."""
    import tempfile

    from petals_b200.data_structures import ModelInfo, ServerInfo, ServerState
    from petals_b200.ops import native
    from petals_b200.parallel.swarm import FileSwarm
    from petals_b200.parallel.tp_worker import TPLeaderEngine, build_tp_engine, follower_loop, make_ring
    from petals_b200.server.backend import Stage
    from petals_b200.server.server import ModuleContainer
    from petals_b200.utils.auto_config import AutoDistributedConfig
    from petals_b200.utils.peaks import measured_peaks
    from petals_b200.utils.random_model import MODEL_PRESETS, random_client_model, write_config_only
    import petals_b200

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    if n_stages * tp != world:
        raise SystemExit(f"--parallelism pp{n_stages}xtp{tp} needs {n_stages * tp} GPUs, the job has {world}")
    os.environ.setdefault("PETALS_B200_SYMM_MEM", "0")  # heaps of sub-groups: plain CUDA IPC (pairs do not need the multicast mapping)
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist.init_process_group(backend="cpu:gloo,cuda:nccl", device_id=dev)
    native.lib()
    groups = [dist.new_group(list(range(g * tp, (g + 1) * tp)), backend="cpu:gloo,cuda:nccl") for g in range(n_stages)]  # collective: every rank, every group
    leaders_group = dist.new_group([i * tp for i in range(n_stages)], backend="cpu:gloo,cuda:nccl")
    g, grank = rank // tp, rank % tp
    group = groups[g]
    path = write_config_only(args.model)
    config = AutoDistributedConfig.from_pretrained(path)
    n_layers = MODEL_PRESETS[args.model]["num_hidden_layers"]
    bounds = [round(i * n_layers / n_stages) for i in range(n_stages + 1)]
    n_mine = bounds[g + 1] - bounds[g]
    do_prefill = not args.skip_prefill
    PB, PT = args.prefill_batch, args.prefill_seq
    max_len = max(args.seq_len, PT) if do_prefill else args.seq_len
    t0 = time.time()
    engine, cache, heap = build_tp_engine(config, n_mine, group=group, seed=g + 1, attn_cache_tokens=(max(args.seq_len, PB * PT) if do_prefill else args.seq_len) + 512,
                                          inference_max_length=max_len, max_prefill_rows=args.tp_prefill_rows)
    ring = make_ring(group)
    dirs = [tempfile.mkdtemp(prefix="pb200-pptp-") if rank == 0 else None]
    dist.broadcast_object_list(dirs, src=0)
    build_s = time.time() - t0
    if grank != 0:
        follower_loop(engine, cache, ring, grank - 1)
        host_barrier()
        heap.close()
        dist.destroy_process_group()
        return
    # the leaders share an NVLink fabric of their own (landing rings, parallel/fabric.py): the hidden state of a token hops from the last
    # kernel's output of stage g to stage g + 1 as a peer copy + flag, and the client issues a step to all four leaders at once
    from petals_b200.parallel.fabric import init_fabric

    fabric = init_fabric(config.hidden_size, max_tokens=max(PB * 256, 1024), group=leaders_group, n_slots=4) if os.environ.get("PETALS_B200_PPTP_FABRIC", "1") != "0" else None
    swarm = FileSwarm(dirs[0])
    leader = TPLeaderEngine(engine, ring)
    stage = Stage(config, [torch.nn.Identity() for _ in range(n_mine)], bounds[g], device=dev, memory_cache=cache, torch_dtype=torch.bfloat16, engine=leader)
    info = ServerInfo(state=ServerState.JOINING, throughput=1.0, version=petals_b200.__version__, torch_dtype="bfloat16", quant_type="none", using_relay=False)
    container = ModuleContainer.from_stage(dht=swarm, dht_prefix=config.dht_prefix, block_config=config, stage=stage, server_info=info,
                                           model_info=ModelInfo(num_blocks=n_layers, repository=path), peer_id=f"stage{g}-tp{tp}", inference_max_length=max_len)
    result = None
    if rank == 0:
        K, W = args.steps, max(args.warmup, 3)
        peaks = measured_peaks()
        model = random_client_model(path, swarm, dev, max_retries=60, min_backoff=0.5, max_backoff=1.0)
        model.model.layers.sequence_manager.update(wait=True)
        vocab = model.config.vocab_size
        spec = config.block_spec()
        prompt = torch.randint(0, vocab, (1, args.prompt_len), device=dev)
        with torch.inference_mode(), model.inference_session(max_length=args.seq_len) as sess:
            prime_session(model, sess, prompt, W)
            sampler = ClockSampler(local_rank)
            sampler.start()
            ms, launches = device_timed_decode(model, sess, K)
            clocks = sampler.stop()
            e2e_s, h2d, d2h = e2e_decode(model, sess, K, dev)
            stages_used = [s_.span.peer_id for s_ in sess._server_sessions]
            over_fabric = [s_.no_history for s_ in sess._server_sessions]
        engine.check_errors()
        value = K / (ms / 1e3)
        prefill = None
        if do_prefill:
            try:
                ids = torch.randint(0, vocab, (PB, PT), device=dev)
                with torch.inference_mode():
                    model.model(input_ids=ids)
                    torch.cuda.synchronize()
                    ps, pe = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    ps.record()
                    for _ in range(args.prefill_steps):
                        model.model(input_ids=ids)
                    pe.record()
                    torch.cuda.synchronize()
                pms = ps.elapsed_time(pe) / args.prefill_steps
                flops = 2.0 * spec.active_params() * n_layers * PB * PT + 4.0 * n_layers * PB * PT * PT * spec.num_heads * spec.head_dim / 2
                prefill = {"tokens_per_s": round(PB * PT / (pms / 1e3), 1), "ms_per_step": round(pms, 2), "batch": PB, "seq_len": PT,
                           "TFLOPs_total": round(flops / pms / 1e9, 1),
                           "path": "micro-batches of <= 1024 tokens as a wavefront over the stages (client/sequential_autograd.py); sequence-parallel tcgen05 GEMMs inside a stage"}
            except Exception as e:  # noqa: BLE001
                prefill = {"error": repr(e)[:300]}
        weight_bytes_rank = (spec.active_params() * n_layers) * 2 / world
        result = {
            "metric": metric_name(args.model), "value": round(value, 3), "unit": "tokens/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": round(ms / K, 4), "higher_is_better": True, "scaling": "strong", "vs_baseline": round(value / BASELINE_TOKENS_PER_S, 3),
            "dtype": "bf16", "data": "synthetic token ids; random-init weights of the named architecture",
            "config": {"model": args.model, "global_batch": 1, "seq_len": args.seq_len, "parallelism": f"pp{n_stages}xtp{tp}",
                       "layout": f"{n_stages} pipeline stages x tensor-parallel groups of {tp} GPUs; blocks per stage {[bounds[i + 1] - bounds[i] for i in range(n_stages)]}",
                       "stages": stages_used, "inputs_over_fabric": over_fabric, "build_s": round(build_s, 1),
                       "l2": "each step streams every rank's full weight shard (>> 126 MB L2): inputs larger than L2",
                       "timing": "CUDA events on rank 0 (client + leader of the first stage), which receives every token from the last stage before the next step starts"},
            "clocks": clocks,
            "e2e": {"value": round(K / e2e_s, 3), "unit": "tokens/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "model.generate(max_new_tokens=1, session=sess) per step; token in from pinned host memory, token out to the host"},
            "gpu_launches": launches, "prefill": prefill,
            "roofline": {"weight_bytes_per_token_per_rank": int(weight_bytes_rank), "note": "single stream: one stage (T ranks) works at a time",
                         "frac_of_measured_hbm_while_active": round(weight_bytes_rank * n_stages * value / 1e9 / peaks["hbm_gbs"], 3)},
        }
    host_barrier_leaders(n_stages, tp)
    leader.shutdown()
    container.shutdown()
    if fabric is not None:
        fabric.check_errors()
        fabric.close()
    host_barrier()
    heap.close()
    if result is not None:
        print(json.dumps(result))
    dist.destroy_process_group()


def host_barrier_leaders(n_stages: int, tp: int) -> None:
    """Leaders wait here until the client (rank 0) is done with every stage; followers are parked in their command loops and reach
    the world barrier only after their leader's shutdown, so this is a file-free rendezvous over the store: a tiny all-gather among
    the leaders would need its own group created collectively — the TCP store of the default group is enough."""
    store = dist.distributed_c10d._get_default_store()
    store.add("pb200_pptp_leaders_done", 1)
    import time as _t

    deadline = _t.time() + 3600
    while int(store.add("pb200_pptp_leaders_done", 0)) < n_stages:
        if _t.time() > deadline:
            raise TimeoutError("leaders of the other stages never finished")
        _t.sleep(0.05)


def run_multi_gpu(args) -> None:
    par = str(args.parallelism)
    if par.startswith("pp") and "xtp" in par:
        s_, t_ = par[2:].split("xtp")
        return run_pp_tp(args, int(s_), int(t_))
    if str(args.parallelism).startswith("pp"):
        return run_pipeline(args)
    from petals_b200.data_structures import ModelInfo, ServerInfo, ServerState
    from petals_b200.ops import native
    from petals_b200.parallel.swarm import Swarm
    from petals_b200.parallel.symmetric import measure_hop_latency, measure_peer_bandwidth
    from petals_b200.parallel.tp_worker import TPLeaderEngine, build_tp_engine, follower_loop, make_ring
    from petals_b200.server.backend import Stage
    from petals_b200.server.server import ModuleContainer
    from petals_b200.utils.auto_config import AutoDistributedConfig
    from petals_b200.utils.peaks import NVLINK_PEER_GBS, measured_peaks
    from petals_b200.utils.random_model import MODEL_PRESETS, random_client_model, write_config_only
    import petals_b200

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist.init_process_group(backend="cpu:gloo,cuda:nccl", device_id=dev)
    native.lib()
    selftests = run_selftests(args, dev, ("tp",) if getattr(args, "skip_pipeline", False) else ("tp", "pp"))
    path = write_config_only(args.model)
    config = AutoDistributedConfig.from_pretrained(path)
    n_layers = MODEL_PRESETS[args.model]["num_hidden_layers"]
    t0 = time.time()
    do_prefill = not args.skip_prefill
    PB, PT = args.prefill_batch, args.prefill_seq
    max_len = max(args.seq_len, PT) if do_prefill else args.seq_len
    engine, cache, heap = build_tp_engine(config, n_layers, attn_cache_tokens=(max(args.seq_len, PB * PT) if do_prefill else args.seq_len) + 512,
                                          inference_max_length=max_len, max_prefill_rows=args.tp_prefill_rows)
    ring = make_ring()
    build_s = time.time() - t0
    # ---- NVLink probes (stage-hop denominators) -------------------------------------------------------------------
    probe_flag = heap.alloc(8)
    peer_gbs = measure_peer_bandwidth(heap, 0, 1)
    hop_us = measure_hop_latency(heap, probe_flag, 0, 1)
    K, W = args.steps, max(args.warmup, 3)

    if rank != 0:
        marks = follower_loop_with_marks(engine, cache, ring, rank - 1)
        ms = marks["start"].elapsed_time(marks["end"]) if "start" in marks and "end" in marks else 0.0
        pms = marks["p_start"].elapsed_time(marks["p_end"]) if "p_start" in marks and "p_end" in marks else 0.0
        gathered = [None] * world
        dist.all_gather_object(gathered, (ms, pms))
        host_barrier()
        heap.close()
        _pipeline_appendix(args, engine, cache)
        dist.destroy_process_group()
        return

    # ---- rank 0: leader + client ----------------------------------------------------------------------------------------
    swarm = Swarm("bench")
    leader = TPLeaderEngine(engine, ring)
    stage = Stage(config, [torch.nn.Identity() for _ in range(n_layers)], 0, device=dev, memory_cache=cache, torch_dtype=torch.bfloat16, engine=leader)
    info = ServerInfo(state=ServerState.JOINING, throughput=1.0, version=petals_b200.__version__, torch_dtype="bfloat16", quant_type="none", using_relay=False)
    container = ModuleContainer.from_stage(dht=swarm, dht_prefix=config.dht_prefix, block_config=config, stage=stage, server_info=info,
                                           model_info=ModelInfo(num_blocks=n_layers, repository=path), peer_id=f"tp{world}-leader",
                                           inference_max_length=max_len)
    model = random_client_model(path, swarm, dev)
    vocab = model.config.vocab_size
    prompt = torch.randint(0, vocab, (1, args.prompt_len), device=dev)
    with torch.inference_mode(), model.inference_session(max_length=args.seq_len) as sess:
        prime_session(model, sess, prompt, W)
        sampler = ClockSampler(local_rank)
        sampler.start()
        ms0, launches = device_timed_decode(model, sess, K, on_start=lambda: ring.send({"op": "mark", "name": "start"}),
                                            on_end=lambda: ring.send({"op": "mark", "name": "end"}))
        clocks = sampler.stop()
        e2e_s, h2d, d2h = e2e_decode(model, sess, K, dev)
    engine.check_errors()
    # ---- prompt ingestion: parallel forward of [PB, PT] tokens through the public client API (rpc_forward) ------------------
    pms0, prefill = 0.0, None
    if do_prefill:
        try:
            ids = torch.randint(0, vocab, (PB, PT), device=dev)
            with torch.inference_mode():
                model.model(input_ids=ids)
                torch.cuda.synchronize()
                ring.send({"op": "mark", "name": "p_start"})
                ps, pe = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ps.record()
                for _ in range(args.prefill_steps):
                    model.model(input_ids=ids)
                pe.record()
                torch.cuda.synchronize()
                ring.send({"op": "mark", "name": "p_end"})
            pms0 = ps.elapsed_time(pe)
            engine.check_errors()
        except Exception as e:  # the decode number stands on its own
            prefill = {"error": repr(e)[:300]}
    leader.shutdown()
    gathered = [None] * world
    dist.all_gather_object(gathered, (ms0, pms0))
    if do_prefill and prefill is None:
        pms = max(float(x[1]) for x in gathered) / args.prefill_steps
        spec_ = config.block_spec()
        flops = 2.0 * spec_.active_params() * n_layers * PB * PT + 4.0 * n_layers * PB * PT * PT * spec_.num_heads * spec_.head_dim / 2
        pk = measured_peaks()
        prefill = {"tokens_per_s": round(PB * PT / (pms / 1e3), 1), "ms_per_step": round(pms, 2), "batch": PB, "seq_len": PT,
                   "TFLOPs_total": round(flops / pms / 1e9, 1), "frac_of_measured_bf16_sustained_per_gpu": round(flops / pms / 1e9 / world / pk["bf16_tflops_sustained"], 3),
                   "path": f"sequence-parallel tcgen05 GEMMs, reduce-scatter in the GEMM epilogue + all-gather in the norm kernel over NVLink, chunks of {engine.max_prefill_rows} rows",
                   "launches_per_chunk_per_rank": getattr(engine, "prefill_launches", None)}
    gathered = [x[0] for x in gathered]
    ms = max(float(x) for x in gathered)
    value = K / (ms / 1e3)
    peaks = measured_peaks()
    spec = config.block_spec()
    weight_bytes_rank = (spec.active_params() * n_layers) * 2 / world + vocab * spec.hidden_size * 2  # LM head is replicated on rank 0
    hop_bytes = spec.hidden_size * 2
    result = {
        "metric": metric_name(args.model),
        "value": round(value, 3), "unit": "tokens/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": round(ms / K, 4),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": round(value / BASELINE_TOKENS_PER_S, 3), "dtype": "bf16",
        "data": "synthetic token ids; random-init weights of the named architecture",
        "config": {"model": args.model, "global_batch": 1, "seq_len": args.seq_len, "parallelism": f"tp{world} (1 stage x {n_layers} blocks, fused GEMV+NVLink all-reduce)",
                   "l2": "each step streams every rank's full weight shard (>> 126 MB L2): inputs larger than L2", "build_s": round(build_s, 1),
                   "per_rank_ms": [round(float(x), 3) for x in gathered]},
        "clocks": clocks,
        "e2e": {"value": round(K / e2e_s, 3), "unit": "tokens/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "model.generate(max_new_tokens=1, session=sess) per step; token in from pinned host memory, token out to the host"},
        "gpu_launches": launches,
        "prefill": prefill,
        "roofline": {"weight_bytes_per_token_per_rank": int(weight_bytes_rank), "achieved_GBps_per_rank": round(weight_bytes_rank * value / 1e9, 1),
                     "frac_of_measured_hbm": round(weight_bytes_rank * value / 1e9 / peaks["hbm_gbs"], 3), "peaks": peaks["source"]},
        "stage_hop": {"peer_store_GBps": None if peer_gbs is None else round(peer_gbs, 1), "frac_of_measured_peer_copy": None if peer_gbs is None else round(peer_gbs / NVLINK_PEER_GBS, 3),
                      "flag_latency_us": None if hop_us is None else round(hop_us, 2), "allreduce_payload_bytes": hop_bytes,
                      "fused_allreduces_per_token": 2 * n_layers},
    }
    container.shutdown()
    host_barrier()
    heap.close()
    result["pipeline"] = _pipeline_appendix(args, engine, cache)
    result["selftests"] = selftests
    if not _selftests_ok(selftests):
        result["invalid"] = "a numerics self-test of the multi-GPU engine failed on this box: the numbers above are not to be trusted"
    print(json.dumps(result))
    dist.destroy_process_group()
    if "invalid" in result:
        raise SystemExit(1)


def _pipeline_appendix(args, engine, cache) -> dict:
    """The same model as N pipeline stages (BASELINE.json configs #2/#3), measured in the same job after the tensor-parallel run has
    released its shards: every rank must call this (collective). A failure here never costs the headline line."""
    if getattr(args, "skip_pipeline", False):
        return {"skipped": True}
    try:
        import gc

        engine.shards.clear()
        engine._graphs.clear()
        engine._span_plan = None
        cache.pool = None
        gc.collect()
        torch.cuda.empty_cache()
        return pipeline_record(args)
    except Exception as e:  # noqa: BLE001
        try:
            host_barrier()
        except Exception:  # noqa: BLE001
            pass
        return {"error": repr(e)[:300]}


def follower_loop_with_marks(engine, cache, ring, consumer: int) -> dict:
    """follower_loop + CUDA-event marks so that every rank times the same K steps on its own device."""
    sessions = {}
    marks = {}
    while True:
        cmd = ring.recv(consumer, timeout=None)
        op = cmd["op"]
        if op in ("step", "prefill"):
            s = sessions[cmd["sid"]]
            if cmd["pos"] != s.position:
                s.set_position(cmd["pos"])
            if "hypo" in cmd:
                s.reorder(torch.tensor(cmd["hypo"], dtype=torch.int64))
            (engine.run_step if op == "step" else engine.run_prefill)(s, cmd["B"], cmd["T"])
        elif op == "open":
            sessions[cmd["sid"]] = cache.open_session(cmd["B"], cmd["max_length"], timeout=None)
        elif op == "close":
            s = sessions.pop(cmd["sid"], None)
            if s is not None:
                s.close()
        elif op == "mark":
            if cmd["name"].endswith("end"):
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                torch.cuda.synchronize()
            else:
                torch.cuda.synchronize()
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
            marks[cmd["name"]] = ev
        elif op == "stop":
            break
    for s in sessions.values():
        s.close()
    torch.cuda.synchronize()
    engine.check_errors()
    return marks


def pipeline_record(args, *, with_decode: bool = True) -> dict:
    """N pipeline stages (one per GPU, equal spans) joined by the NVLink landing rings; the public client API on rank 0.
    Collective over the initialised process group; returns the record on rank 0 and {} elsewhere.

    * decode: single stream, one token per step; every hop is the producing stage's last kernel storing into the next stage's
      landing slot (no tensor in any RPC);
    * prompt ingestion: [prefill_batch, prefill_seq] tokens through an inference session; the client cuts the step into chunks of
      ``pipeline_chunk_tokens`` positions and runs them as a wavefront over the stages (client/inference_session.py), the
      chunks travel through the landing rings. Ideal bubble fraction (S - 1) / (M + S - 1) for M chunks over S stages."""
    import tempfile

    from petals_b200.ops import native
    from petals_b200.parallel.fabric import init_fabric
    from petals_b200.parallel.swarm import FileSwarm
    from petals_b200.parallel.symmetric import measure_hop_latency, measure_peer_bandwidth
    from petals_b200.utils.auto_config import AutoDistributedConfig
    from petals_b200.utils.peaks import NVLINK_PEER_GBS, measured_peaks
    from petals_b200.utils.random_model import MODEL_PRESETS, launch_random_stage, random_client_model, write_config_only

    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device("cuda", torch.cuda.current_device())
    path = write_config_only(args.model)
    config = AutoDistributedConfig.from_pretrained(path)
    n_layers = MODEL_PRESETS[args.model]["num_hidden_layers"]
    PB, PT = args.prefill_batch, args.prefill_seq
    chunk = int(getattr(args, "pp_chunk_tokens", 256))
    do_prefill = not args.skip_prefill
    fabric = init_fabric(config.hidden_size, max_tokens=max(PB * chunk, 64), n_slots=4)
    probe = fabric.heap.alloc(8)
    peer_gbs = measure_peer_bandwidth(fabric.heap, 0, 1)
    hop_us = measure_hop_latency(fabric.heap, probe, 0, 1)
    dirs = [tempfile.mkdtemp(prefix="pb200-bench-") if rank == 0 else None]
    dist.broadcast_object_list(dirs, src=0)
    swarm = FileSwarm(dirs[0])
    bounds = [round(i * n_layers / world) for i in range(world + 1)]
    t0 = time.time()
    cache_tokens = (max(args.seq_len, PB * (PT + 128)) if do_prefill else args.seq_len) + 256  # sessions round every row up to whole pages
    stage = launch_random_stage(path, range(bounds[rank], bounds[rank + 1]), swarm, dev, peer_id=f"stage{rank}", attn_cache_tokens=cache_tokens,
                                inference_max_length=max(args.seq_len, PT), max_batch_size=1 << 20)
    torch.cuda.synchronize()
    host_barrier()
    build_s = time.time() - t0
    K, W = args.steps, max(args.warmup, 3)
    record: dict = {}
    if rank == 0:
        model = random_client_model(path, swarm, dev, pipeline_chunk_tokens=chunk)
        vocab = model.config.vocab_size
        record = {"parallelism": f"pp{world} ({world} stages x {n_layers // world} blocks, NVLink landing rings)", "build_s": round(build_s, 1),
                  "stage_hop": {"peer_store_GBps": None if peer_gbs is None else round(peer_gbs, 1),
                                "frac_of_measured_peer_copy": None if peer_gbs is None else round(peer_gbs / NVLINK_PEER_GBS, 3),
                                "flag_latency_us": None if hop_us is None else round(hop_us, 2), "hop_payload_bytes": config.hidden_size * 2,
                                "hops_per_token": world, "reference_hop_model_ms": 18.0}}
        if with_decode:
            prompt = torch.randint(0, vocab, (1, args.prompt_len), device=dev)
            with torch.inference_mode(), model.inference_session(max_length=args.seq_len) as sess:
                prime_session(model, sess, prompt, W)
                sampler = ClockSampler(dev.index)
                sampler.start()
                # rank 0 waits for the last stage's result every step: its device time IS the max over ranks
                ms, launches = device_timed_decode(model, sess, K)
                record["clocks"] = sampler.stop()
                e2e_s, h2d, d2h = e2e_decode(model, sess, K, dev)
                record["inputs_over_fabric"] = [s.no_history for s in sess._server_sessions]
            record["decode"] = {"tokens_per_s": round(K / (ms / 1e3), 3), "ms_per_step": round(ms / K, 4), "gpu_launches_rank0": launches,
                                "e2e_tokens_per_s": round(K / e2e_s, 3), "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h}
        if do_prefill:
            try:
                ids = torch.randint(0, vocab, (PB, PT), device=dev)
                times = []
                for it in range(1 + args.prefill_steps):
                    with torch.inference_mode(), model.inference_session(max_length=PT) as sess:
                        torch.cuda.synchronize()
                        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        a.record()
                        model.model(input_ids=ids)
                        b.record()
                        torch.cuda.synchronize()
                        if it > 0:
                            times.append(a.elapsed_time(b))
                pms = sum(times) / len(times)
                spec_ = config.block_spec()
                flops = 2.0 * spec_.active_params() * n_layers * PB * PT + 4.0 * n_layers * PB * PT * PT * spec_.num_heads * spec_.head_dim / 2
                pk = measured_peaks()
                M = (PT + chunk - 1) // chunk
                record["prefill"] = {"tokens_per_s": round(PB * PT / (pms / 1e3), 1), "ms_per_step": round(pms, 2), "batch": PB, "seq_len": PT,
                                     "chunk_tokens": chunk, "chunks": M, "stages": world, "ideal_bubble_fraction": round((world - 1) / (M + world - 1), 3),
                                     "TFLOPs_total": round(flops / pms / 1e9, 1),
                                     "frac_of_measured_bf16_sustained_per_gpu": round(flops / pms / 1e9 / world / pk["bf16_tflops_sustained"], 3),
                                     "path": "chunked prompt ingestion as a wavefront over the stages; chunks travel through the landing rings (fused GEMM-epilogue push)"}
            except Exception as e:  # noqa: BLE001 - the decode number stands on its own
                record["prefill"] = {"error": repr(e)[:300]}
        fabric.check_errors()
    host_barrier()
    stage.shutdown()
    host_barrier()
    fabric.close()
    return record


def run_pipeline(args) -> None:
    """``bench.py --parallelism ppN``: the pipeline layout as the headline line."""
    from petals_b200.ops import native

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist.init_process_group(backend="cpu:gloo,cuda:nccl", device_id=dev)
    native.lib()
    selftests = run_selftests(args, dev, ("pp",))
    rec = pipeline_record(args)
    if rank == 0:
        d = rec.get("decode", {})
        value = d.get("tokens_per_s", 0.0)
        result = {
            "metric": metric_name(args.model), "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": d.get("ms_per_step"), "higher_is_better": True, "scaling": "strong", "vs_baseline": round(value / BASELINE_TOKENS_PER_S, 3),
            "dtype": "bf16", "data": "synthetic token ids; random-init weights of the named architecture",
            "config": {"model": args.model, "global_batch": 1, "seq_len": args.seq_len, "parallelism": rec.get("parallelism"),
                       "l2": "each step streams every stage's full weight span (>> 126 MB L2): inputs larger than L2", "build_s": rec.get("build_s"),
                       "inputs_over_fabric": rec.get("inputs_over_fabric")},
            "clocks": rec.get("clocks"),
            "e2e": {"value": d.get("e2e_tokens_per_s"), "unit": "tokens/s", "h2d_bytes_per_step": d.get("h2d_bytes_per_step"), "d2h_bytes_per_step": d.get("d2h_bytes_per_step")},
            "gpu_launches": d.get("gpu_launches_rank0"), "prefill": rec.get("prefill"), "stage_hop": rec.get("stage_hop"), "selftests": selftests,
        }
        if not _selftests_ok(selftests):
            result["invalid"] = "the numerics self-test of the pipeline fabric failed on this box: the numbers above are not to be trusted"
        print(json.dumps(result))
        if "invalid" in result:
            dist.destroy_process_group()
            raise SystemExit(1)
    dist.destroy_process_group()
