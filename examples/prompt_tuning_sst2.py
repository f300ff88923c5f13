#!/usr/bin/env python3
"""Prompt-tune a sequence-classification head on top of a distributed Llama (reference: examples/prompt-tuning-sst2.ipynb).

The notebook fine-tunes deep prompts + a linear `score` head on SST-2 with AMP. There is no dataset on an offline box, so
this script uses a synthetic two-class problem with the same moving parts: a stage swarm serving frozen blocks, a client
with `tuning_mode="deep_ptune"`, AdamW on the client-side parameters only, autocast + GradScaler-free bf16, and
save/load of the trained prompts (the "checkpoint" of a Petals fine-tune is just the client state_dict).

    python examples/prompt_tuning_sst2.py --steps 30            # CPU, tiny random model
    python examples/prompt_tuning_sst2.py --device cuda:0 --model llama-tiny --torch_dtype bfloat16
"""
import argparse
import os
import sys
import tempfile

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "benchmarks"))
from _common import add_common_args, swarm_and_model  # noqa: E402


def synthetic_batch(vocab, batch, seq_len, device, gen):
    """Class 1 sequences contain mostly tokens from the upper half of the vocabulary."""
    labels = torch.randint(0, 2, (batch,), generator=gen)
    lo = torch.randint(0, vocab // 2, (batch, seq_len), generator=gen)
    hi = torch.randint(vocab // 2, vocab, (batch, seq_len), generator=gen)
    ids = torch.where(labels[:, None].bool(), hi, lo)
    return ids.to(device), labels.to(device)


def main():
    parser = argparse.ArgumentParser()
    add_common_args(parser)
    parser.add_argument("--steps", type=int, default=30)
    parser.add_argument("--batch_size", type=int, default=8)
    parser.add_argument("--seq_len", type=int, default=16)
    parser.add_argument("--pre_seq_len", type=int, default=8)
    parser.add_argument("--lr", type=float, default=2e-3)
    args = parser.parse_args()
    if args.device == "cpu" and args.torch_dtype == "bfloat16":
        args.torch_dtype = "float32"
    gen = torch.Generator().manual_seed(0)
    with swarm_and_model(args, model_class="model_for_sequence_classification", tuning_mode="deep_ptune", pre_seq_len=args.pre_seq_len,
                         num_labels=2) as model:
        trainable = [p for p in model.parameters() if p.requires_grad]
        print(f"trainable client parameters: {sum(p.numel() for p in trainable):,} (blocks are frozen on the stages)")
        opt = torch.optim.AdamW(trainable, lr=args.lr, weight_decay=0.0)
        for step in range(args.steps):
            ids, labels = synthetic_batch(model.config.vocab_size, args.batch_size, args.seq_len, args.device, gen)
            out = model(input_ids=ids, labels=labels)
            out.loss.backward()
            opt.step()
            opt.zero_grad()
            if step % 5 == 0 or step == args.steps - 1:
                acc = (out.logits.argmax(-1) == labels).float().mean().item()
                print(f"step {step:3d}  loss {out.loss.item():.4f}  batch acc {acc:.2f}")
        ckpt = os.path.join(tempfile.gettempdir(), "petals_b200_ptune_sst2.pt")
        torch.save({k: v for k, v in model.state_dict().items() if "prompt" in k or k.startswith("score")}, ckpt)
        missing = model.load_state_dict(torch.load(ckpt), strict=False)
        print(f"saved and re-loaded {ckpt} (prompt + head parameters; {len(missing.missing_keys)} frozen keys untouched)")


if __name__ == "__main__":
    main()
