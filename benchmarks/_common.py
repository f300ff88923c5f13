"""Shared helpers of the benchmark scripts: start an in-process swarm with random weights when no --initial_peers is given."""
from __future__ import annotations

import contextlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def add_common_args(parser):
    parser.add_argument("--model", type=str, default="llama-tiny", help="preset name (random-init) or a checkpoint directory")
    parser.add_argument("--initial_peers", type=str, nargs="+", default=None, help="rendezvous of a running swarm (default: start one here)")
    parser.add_argument("--torch_dtype", type=str, default="bfloat16")
    parser.add_argument("--device", type=str, default="cuda:0" if torch.cuda.is_available() else "cpu")
    parser.add_argument("--n_stages", type=int, default=1, help="pipeline stages of the private swarm (same device)")
    parser.add_argument("--warmup_steps", type=int, default=1)


@contextlib.contextmanager
def swarm_and_model(args, model_class: str = "model_for_causal_lm", **model_kwargs):
    """Yields a client model connected to ``args.initial_peers`` or to a private swarm started for this run."""
    from petals_b200.constants import DTYPE_MAP
    from petals_b200.parallel.swarm import Swarm, resolve_swarm
    from petals_b200.utils.auto_config import AutoDistributedConfig, detect_model_type, get_model_classes
    from petals_b200.utils.random_model import MODEL_PRESETS, launch_random_stage, random_client_model, write_config_only

    dtype = DTYPE_MAP[args.torch_dtype]
    stages = []
    try:
        if args.model in MODEL_PRESETS:
            path = write_config_only(args.model, {"torch_dtype": args.torch_dtype})
            random = True
        else:
            path, random = args.model, False
        if args.initial_peers:
            swarm = resolve_swarm(args.initial_peers)
        else:
            swarm = Swarm(f"bench-{os.getpid()}")
            n = AutoDistributedConfig.from_pretrained(path).num_hidden_layers
            bounds = [round(i * n / args.n_stages) for i in range(args.n_stages + 1)]
            for lo, hi in zip(bounds[:-1], bounds[1:]):
                if random:
                    stages.append(launch_random_stage(path, range(lo, hi), swarm, args.device, dtype=dtype, attn_cache_tokens=16384,
                                                      inference_max_length=8192, peer_id=f"stage{lo}-{hi}"))
                else:
                    from petals_b200.server.server import Server

                    s = Server(initial_peers=swarm, converted_model_name_or_path=path, block_indices=f"{lo}:{hi}", torch_dtype=args.torch_dtype,
                               device=args.device, throughput=1.0)
                    s.run_in_background()
                    stages.append(s)
        if random:
            model = random_client_model(path, swarm, args.device, dtype=dtype, model_class=model_class, **model_kwargs)
        else:
            cls = get_model_classes(detect_model_type(path))[model_class]
            model = cls.from_pretrained(path, initial_peers=swarm, torch_dtype=dtype, device=args.device, **model_kwargs)
        yield model
    finally:
        for s in stages:
            s.shutdown()


def sync(device) -> None:
    if torch.device(device).type == "cuda":
        torch.cuda.synchronize(device)
