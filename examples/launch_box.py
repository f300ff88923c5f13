#!/usr/bin/env python
"""Serve one model from every GPU of a box: a rendezvous directory, one `run_server` process per GPU with an even split of the blocks,
all of them members of one NVLink landing-ring fabric (hidden states, training micro-batches and gradients hop GPU to GPU between them).

    python examples/launch_box.py /path/to/llama --gpus 8                  # 8 pipeline stages on cuda:0..7
    python examples/launch_box.py /path/to/model --gpus 2 --device cpu     # the same plumbing on CPU (shared-memory fabric)

Prints the `--initial_peers` value clients connect with, then waits; Ctrl-C stops the servers. Extra arguments after `--` go to every
`run_server` (e.g. `-- --quant_type fp8 --attn_cache_tokens 32768`)."""
import argparse
import json
import os
import signal
import socket
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def split_blocks(n_blocks: int, n_stages: int):
    bounds = [round(i * n_blocks / n_stages) for i in range(n_stages + 1)]
    return [(bounds[i], bounds[i + 1]) for i in range(n_stages)]


def free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("model", help="path of a Hugging Face checkpoint directory")
    ap.add_argument("--gpus", type=int, default=None, help="number of stage processes (default: every visible GPU)")
    ap.add_argument("--device", default="cuda", choices=["cuda", "cpu"])
    ap.add_argument("--rendezvous", default=None, help="swarm directory (default: a fresh one under /dev/shm or the temp dir)")
    ap.add_argument("--torch_dtype", default=None)
    ap.add_argument("--no_fabric", action="store_true", help="do not form the landing-ring fabric (every hop carries its tensor with the RPC)")
    ap.add_argument("--fabric_max_tokens", type=int, default=8192)
    argv = list(sys.argv[1:] if argv is None else argv)
    extra = []
    if "--" in argv:  # everything after `--` goes to every run_server
        cut = argv.index("--")
        argv, extra = argv[:cut], argv[cut + 1:]
    args = ap.parse_args(argv)
    n_blocks = json.load(open(os.path.join(args.model, "config.json")))
    n_blocks = n_blocks.get("num_hidden_layers", n_blocks.get("n_layer"))
    if args.gpus is None:
        import torch

        args.gpus = max(1, torch.cuda.device_count()) if args.device == "cuda" else 2
    if not 1 <= args.gpus <= n_blocks:
        ap.error(f"--gpus must be between 1 and the number of blocks ({n_blocks})")
    rendezvous = args.rendezvous or tempfile.mkdtemp(prefix="petals-swarm-", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    subprocess.run([sys.executable, "-m", "petals.cli.run_dht", "--rendezvous", rendezvous, "--once"], check=True, env=env)
    fabric = [] if args.no_fabric or args.gpus < 2 else ["--fabric_address", f"127.0.0.1:{free_port()}", "--fabric_world", str(args.gpus),
                                                         "--fabric_max_tokens", str(args.fabric_max_tokens)]
    procs = []
    for i, (lo, hi) in enumerate(split_blocks(n_blocks, args.gpus)):
        cmd = [sys.executable, "-m", "petals.cli.run_server", args.model, "--initial_peers", rendezvous, "--block_indices", f"{lo}:{hi}",
               "--device", f"cuda:{i}" if args.device == "cuda" else "cpu", "--peer_id", f"stage{i}", *extra]
        if args.torch_dtype:
            cmd += ["--torch_dtype", args.torch_dtype]
        if fabric:
            cmd += [*fabric, "--fabric_rank", str(i)]
        procs.append(subprocess.Popen(cmd, env=env))
    print(f"serving blocks 0:{n_blocks} from {args.gpus} stage processes; clients: --initial_peers {rendezvous}", flush=True)

    def stop(*_):
        for p in procs:
            p.terminate()

    signal.signal(signal.SIGINT, stop)
    signal.signal(signal.SIGTERM, stop)
    code = 0
    for p in procs:
        try:
            code = p.wait() or code
        except KeyboardInterrupt:
            stop()
    return code


if __name__ == "__main__":
    sys.exit(main())
