#!/usr/bin/env python3
"""Single-stream generation speed (reference: benchmarks/benchmark_inference.py:44-68): one inference session with
``max_length = seq_len``, then ``generate(max_new_tokens=1, session=sess)`` per step; tokens/s = 1 / mean step time."""
import argparse
from time import perf_counter

import numpy as np
import torch

from _common import add_common_args, swarm_and_model, sync


def main():
    parser = argparse.ArgumentParser(formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    add_common_args(parser)
    parser.add_argument("--seq_len", type=int, default=2048)
    parser.add_argument("--prompt_len", type=int, default=16)
    args = parser.parse_args()
    with swarm_and_model(args) as model:
        ids = torch.randint(0, model.config.vocab_size, (1, args.prompt_len), device=args.device)
        step_times = []
        with model.inference_session(max_length=args.seq_len) as sess:
            model.generate(ids, max_new_tokens=1, session=sess)
            for step in range(args.seq_len - args.prompt_len - 1):
                sync(args.device)
                start = perf_counter()
                model.generate(max_new_tokens=1, session=sess)
                sync(args.device)
                if step >= args.warmup_steps:
                    step_times.append(perf_counter() - start)
                if step >= args.warmup_steps and (step + 1) % 128 == 0:
                    print(f"step {step + 1}: {1 / np.mean(step_times):.2f} tokens/sec", flush=True)
        print(f"Final result: speed={1 / np.mean(step_times):.2f} tokens/sec ({1e3 * np.mean(step_times):.3f} ms/token)")


if __name__ == "__main__":
    main()
