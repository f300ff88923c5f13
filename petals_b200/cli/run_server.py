"""``python -m petals.cli.run_server MODEL [flags]`` — start one stage worker
(reference: src/petals/cli/run_server.py:19-231; flag names are kept, optional ``config.yml`` via ``-c``).

Networking-only flags of the reference (``--public_ip``, relays, identity, reachability ...) are accepted and
ignored: a stage is addressed by the rendezvous location (``--initial_peers``) and its GPU."""
from __future__ import annotations

import argparse
import os
import signal

import torch
import yaml

from petals_b200.constants import DTYPE_MAP
from petals_b200.server.server import Server
from petals_b200.utils.compression import parse_compression
from petals_b200.utils.convert_block import QuantType
from petals_b200.utils.logging import get_logger
from petals_b200.utils.version import validate_version

logger = get_logger(__name__)


def parse_size(value: str) -> int:
    """'300GB' / '4GiB' / '1024' -> bytes (the reference uses humanfriendly)."""
    s = value.strip().lower().replace("ib", "b")
    units = {"kb": 10**3, "mb": 10**6, "gb": 10**9, "tb": 10**12, "b": 1}
    for u in ("kb", "mb", "gb", "tb", "b"):
        if s.endswith(u):
            return int(float(s[: -len(u)]) * units[u])
    return int(float(s))


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    p.add_argument("-c", "--config", default=None, help="yaml file with default values for any of the flags")
    g = p.add_mutually_exclusive_group(required=False)
    g.add_argument("--converted_model_name_or_path", type=str, default=None, help="path or name of a pretrained model")
    g.add_argument("model", nargs="?", type=str, default=None, help="same as --converted_model_name_or_path")
    p.add_argument("--public_name", type=str, default=None)
    g = p.add_mutually_exclusive_group(required=False)
    g.add_argument("--token", type=str, default=None)
    g.add_argument("--use_auth_token", action="store_true", dest="token")
    p.add_argument("--num_blocks", type=int, default=None, help="number of transformer blocks to serve")
    p.add_argument("--block_indices", type=str, default=None, help="specific block indices to serve, e.g. 0:40")
    p.add_argument("--dht_prefix", type=str, default=None)
    p.add_argument("--port", type=int, default=None, help="(ignored: RPC ports are picked by the OS and announced through the registry)")
    p.add_argument("--host_maddrs", nargs="+", default=None,
                   help="multi-box swarms (tcp:// or /ip4/ initial peers): interface to serve RPCs on, e.g. /ip4/0.0.0.0/tcp/0 (default)")
    p.add_argument("--announce_maddrs", nargs="+", default=None,
                   help="multi-box swarms: address other peers should use to reach this server, e.g. /ip4/10.0.0.5/tcp/0")
    p.add_argument("--public_ip", type=str, default=None, help="multi-box swarms: shorthand for --announce_maddrs /ip4/<ip>/tcp/0")
    p.add_argument("--no_auto_relay", action="store_false", dest="use_auto_relay")
    p.add_argument("--daemon_startup_timeout", type=float, default=60)
    p.add_argument("--fabric_address", type=str, default=None, help="HOST:PORT where the stage processes of this NVLink box rendezvous to form a "
                        "landing-ring fabric: between members, hidden states / micro-batches / gradients hop GPU to GPU instead of travelling with the RPCs")
    p.add_argument("--fabric_rank", type=int, default=None, help="this process's index among the fabric members (0 .. fabric_world - 1)")
    p.add_argument("--fabric_world", type=int, default=None, help="number of stage processes forming the fabric (all must start; joining is collective)")
    p.add_argument("--fabric_max_tokens", type=int, default=8192, help="rows (batch x positions) one landing slot holds; larger steps travel with the RPCs")
    p.add_argument("--compression", type=str, default="NONE", help="wire codec for the hidden states this server returns over the socket transport: NONE, FLOAT16, MEANSTD_16BIT, "
                        "QUANTILE_8BIT, UNIFORM_8BIT, BLOCKWISE_8BIT or MXFP8 (clients can override per request with output_compression; "
                        "NVLink stage hops are never compressed)")
    p.add_argument("--num_handlers", type=int, default=8)
    p.add_argument("--prefetch_batches", type=int, default=1)
    p.add_argument("--sender_threads", type=int, default=1)
    p.add_argument("--inference_max_length", type=int, default=None)
    p.add_argument("--min_batch_size", type=int, default=1)
    p.add_argument("--max_batch_size", type=int, default=None, help="max tokens per forward/backward/inference task")
    p.add_argument("--max_chunk_size_bytes", type=int, default=256 * 1024 * 1024)
    p.add_argument("--attn_cache_tokens", type=int, default=None, help="KV budget in tokens per block")
    p.add_argument("--cache_dir", type=str, default=None)
    p.add_argument("--max_disk_space", type=str, default=None)
    p.add_argument("--device", type=str, default=None)
    p.add_argument("--torch_dtype", type=str, choices=list(DTYPE_MAP.keys()), default="auto")
    p.add_argument("--max_alloc_timeout", type=float, default=600)
    p.add_argument("--revision", type=str, default=None)
    p.add_argument("--throughput", type=lambda v: v if v in ("auto", "eval", "dry_run") else float(v), default="auto")
    p.add_argument("--update_period", type=float, default=120)
    p.add_argument("--expiration", type=float, default=None)
    p.add_argument("--request_timeout", type=float, default=3 * 60)
    p.add_argument("--session_timeout", type=float, default=30 * 60)
    p.add_argument("--step_timeout", type=float, default=5 * 60)
    g = p.add_mutually_exclusive_group()
    g.add_argument("--initial_peers", type=str, nargs="+", default=None, help="rendezvous location of the swarm")
    g.add_argument("--new_swarm", action="store_true", help="start a private in-process swarm")
    p.add_argument("--increase_file_limit", type=int, default=None, help="(ignored)")
    p.add_argument("--stats_report_interval", type=int, default=None)
    p.add_argument("--custom_module_path", type=str, default=None)
    p.add_argument("--identity_path", type=str, default=None, help="(ignored)")
    p.add_argument("--balance_quality", type=float, default=0.75)
    p.add_argument("--mean_balance_check_period", type=float, default=60)
    p.add_argument("--quant_type", type=str, default=None, choices=[c.name.lower() for c in QuantType])
    p.add_argument("--tensor_parallel_devices", nargs="+", default=None)
    p.add_argument("--skip_reachability_check", action="store_true")
    p.add_argument("--metrics_port", type=int, default=None, help="serve Prometheus metrics (requests, tokens, latency histograms, sessions, "
                   "KV tokens left) on this port; 0 picks a free one")
    p.add_argument("--adapters", nargs="*", default=())
    p.add_argument("--peer_id", type=str, default=None, help="name of this stage in the swarm (default: derived from the device)")
    return p


def main(argv=None) -> None:
    parser = build_parser()
    pre, _ = parser.parse_known_args(argv)
    if pre.config:
        with open(pre.config) as f:
            parser.set_defaults(**(yaml.safe_load(f) or {}))
    args = vars(parser.parse_args(argv))
    args.pop("config", None)
    args["converted_model_name_or_path"] = args.pop("model") or args["converted_model_name_or_path"]
    if not args["converted_model_name_or_path"]:
        parser.error("a model name or path is required")
    for ignored in ("port", "increase_file_limit", "identity_path", "daemon_startup_timeout"):
        args.pop(ignored, None)
    if args.pop("new_swarm"):
        args["initial_peers"] = []
    max_disk_space = args.pop("max_disk_space")
    args["max_disk_space"] = parse_size(max_disk_space) if max_disk_space is not None else None
    args["compression"] = parse_compression(args["compression"])
    if args["quant_type"] is not None:
        args["quant_type"] = QuantType[args["quant_type"].upper()]
    if args["tensor_parallel_devices"]:
        args["tensor_parallel_devices"] = [torch.device(d) for d in args["tensor_parallel_devices"]]
    validate_version()
    server = Server(**args)
    signal.signal(signal.SIGTERM, lambda *_: server.stop.set())
    try:
        server.run()
    except KeyboardInterrupt:
        logger.info("Caught KeyboardInterrupt, shutting down")
    finally:
        server.shutdown()


if __name__ == "__main__":
    main()
