#!/usr/bin/env python3
"""Prompt-tuning speed (reference: benchmarks/benchmark_training.py:50-103): ``deep_ptune`` with ``pre_seq_len`` trainable
prompt tokens, task ``cls`` (sequence classification head) or ``causal_lm``, Adam on the client; reports forward and
backward tokens/s separately."""
import argparse
from time import perf_counter

import numpy as np
import torch

from _common import add_common_args, swarm_and_model, sync


def main():
    parser = argparse.ArgumentParser(formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    add_common_args(parser)
    parser.add_argument("--task", type=str, default="cls", choices=["cls", "causal_lm"])
    parser.add_argument("--seq_len", type=int, default=128)
    parser.add_argument("--batch_size", type=int, default=8)
    parser.add_argument("--pre_seq_len", type=int, default=16)
    parser.add_argument("--n_steps", type=int, default=10)
    args = parser.parse_args()
    cls = "model_for_sequence_classification" if args.task == "cls" else "model_for_causal_lm"
    with swarm_and_model(args, model_class=cls, tuning_mode="deep_ptune", pre_seq_len=args.pre_seq_len) as model:
        params = [p for p in model.parameters() if p.requires_grad]
        opt = torch.optim.Adam(params, lr=1e-3)
        fwd_times, bwd_times = [], []
        for step in range(args.warmup_steps + args.n_steps):
            ids = torch.randint(0, model.config.vocab_size, (args.batch_size, args.seq_len), device=args.device)
            labels = torch.randint(0, 2, (args.batch_size,), device=args.device) if args.task == "cls" else ids
            sync(args.device)
            start = perf_counter()
            loss = model(input_ids=ids, labels=labels).loss
            sync(args.device)
            mid = perf_counter()
            loss.backward()
            opt.step()
            opt.zero_grad()
            sync(args.device)
            if step >= args.warmup_steps:
                fwd_times.append(mid - start)
                bwd_times.append(perf_counter() - mid)
        tokens = args.batch_size * args.seq_len
        print(f"Final result: fwd_speed={tokens / np.mean(fwd_times):.2f} tokens/sec, bwd_speed={tokens / np.mean(bwd_times):.2f} tokens/sec")


if __name__ == "__main__":
    main()
