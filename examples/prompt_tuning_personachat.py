#!/usr/bin/env python3
"""Prompt-tune a causal LM as a chatbot and talk to it across several generate() calls in ONE inference session
(reference: examples/prompt-tuning-personachat.ipynb). Synthetic "dialogues" replace PersonaChat (offline box).

    python examples/prompt_tuning_personachat.py --steps 20
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "benchmarks"))
from _common import add_common_args, swarm_and_model  # noqa: E402


def main():
    parser = argparse.ArgumentParser()
    add_common_args(parser)
    parser.add_argument("--steps", type=int, default=20)
    parser.add_argument("--batch_size", type=int, default=4)
    parser.add_argument("--seq_len", type=int, default=24)
    parser.add_argument("--pre_seq_len", type=int, default=8)
    args = parser.parse_args()
    if args.device == "cpu" and args.torch_dtype == "bfloat16":
        args.torch_dtype = "float32"
    gen = torch.Generator().manual_seed(0)
    with swarm_and_model(args, tuning_mode="ptune", pre_seq_len=args.pre_seq_len) as model:
        opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-2)
        pattern = torch.randint(0, model.config.vocab_size, (1, 6), generator=gen)  # the "persona": a phrase the bot should repeat
        for step in range(args.steps):
            ids = pattern.repeat(args.batch_size, args.seq_len // 6).to(args.device)
            loss = model(input_ids=ids, labels=ids).loss
            loss.backward()
            opt.step()
            opt.zero_grad()
            if step % 5 == 0 or step == args.steps - 1:
                print(f"step {step:3d}  lm loss {loss.item():.4f}")
        # interactive-style inference: one session, several user turns, the KV cache persists between calls
        with model.inference_session(max_length=128) as sess:
            reply = model.generate(pattern.to(args.device), max_new_tokens=6, session=sess)
            print("turn 1:", reply[0, -6:].tolist())
            user = torch.randint(0, model.config.vocab_size, (1, 3), generator=gen).to(args.device)
            reply = model.generate(user, max_new_tokens=6, session=sess)  # appended to the same conversation
            print("turn 2:", reply[0, -6:].tolist(), f"(session position {sess.position})")


if __name__ == "__main__":
    main()
