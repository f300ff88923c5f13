#!/usr/bin/env python3
"""Parallel forward speed (reference: benchmarks/benchmark_forward.py:45-71): random ids through the backbone (no LM head);
tokens/s = batch * seq_len / mean step time."""
import argparse
from time import perf_counter

import numpy as np
import torch

from _common import add_common_args, swarm_and_model, sync


def main():
    parser = argparse.ArgumentParser(formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    add_common_args(parser)
    parser.add_argument("--seq_len", type=int, default=128)
    parser.add_argument("--batch_size", type=int, default=1)
    parser.add_argument("--n_steps", type=int, default=100)
    args = parser.parse_args()
    with swarm_and_model(args, model_class="model") as model, torch.inference_mode():
        times = []
        for step in range(args.warmup_steps + args.n_steps):
            ids = torch.randint(0, model.config.vocab_size, (args.batch_size, args.seq_len), device=args.device)
            sync(args.device)
            start = perf_counter()
            model(input_ids=ids)
            sync(args.device)
            if step >= args.warmup_steps:
                times.append(perf_counter() - start)
        speed = args.batch_size * args.seq_len / np.mean(times)
        print(f"Final result: speed={speed:.2f} tokens/sec ({1e3 * np.mean(times):.3f} ms/step)")


if __name__ == "__main__":
    main()
