"""`Server(tensor_parallel_devices=[cuda:0, cuda:1, ...])` end to end (run as a plain script on a box with >= 2 GPUs).

A real checkpoint directory is written, ONE server process is started the way `python -m petals.cli.run_server PATH
--tensor_parallel_devices cuda:0 cuda:1` starts it (this process leads, a spawned worker per extra device follows), and a client
generates through the public API. The logits are compared with the dense blocks loaded from the same checkpoint."""
import json
import os
import sys
import tempfile

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from petals_b200.parallel.swarm import Swarm
    from petals_b200.server.from_pretrained import load_pretrained_block
    from petals_b200.server.server import Server
    from petals_b200.utils.auto_config import AutoDistributedConfig, AutoDistributedModelForCausalLM
    from petals_b200.utils.checkpoints import make_random_checkpoint

    world = min(torch.cuda.device_count(), int(os.environ.get("TP_WORLD", "2")))
    path = make_random_checkpoint(tempfile.mkdtemp(prefix="pb200-tpserver-"), "llama", dtype=torch.bfloat16, hidden_size=512, intermediate_size=1024,
                                  num_attention_heads=8, num_key_value_heads=max(2, world), num_hidden_layers=3)
    swarm = Swarm("tp-server-selftest")
    server = Server(initial_peers=swarm, converted_model_name_or_path=path, block_indices="0:3", torch_dtype="bfloat16", device="cuda:0",
                    tensor_parallel_devices=[f"cuda:{i}" for i in range(world)], throughput=1.0, update_period=0.5, mean_balance_check_period=1000,
                    attn_cache_tokens=1024, inference_max_length=512)
    server.run_in_background(timeout=600)
    ok, report = False, {}
    try:
        model = AutoDistributedModelForCausalLM.from_pretrained(path, initial_peers=swarm, torch_dtype=torch.bfloat16).to("cuda:0")
        config = AutoDistributedConfig.from_pretrained(path)
        torch.manual_seed(0)
        ids = torch.randint(0, config.vocab_size, (1, 180), device="cuda:0")
        with torch.inference_mode():
            h = model.model.embed(ids)
            for i in range(config.num_hidden_layers):
                h = load_pretrained_block(path, i, torch_dtype=torch.bfloat16).to("cuda:0").forward_cached(h, None, None, 0)
            ref = model.lm_head(model.model.final_norm(h)).float()
            with model.inference_session(max_length=256):
                got = torch.cat([model(ids[:, :150]).logits, model(ids[:, 150:151]).logits, model(ids[:, 151:]).logits], 1).float()
            fwd = model(ids).logits.float()  # cache-less parallel forward (rpc_forward) through the TP group
            out = model.generate(ids[:, :8], max_new_tokens=5)
        bwd = {}
        if os.environ.get("TP_SELFTEST_BACKWARD", "0") == "1":
            # rpc_backward through the worker group (parallel/tp_worker.py: tp_collective_backward) vs autograd through the dense blocks
            x = (torch.randn(2, 24, config.hidden_size, device="cuda:0") * 0.5).to(torch.bfloat16).requires_grad_(True)
            model.model.layers(x).float().pow(2).sum().backward()
            g_tp = x.grad.float().clone()
            x.grad = None
            hh = x
            for i in range(config.num_hidden_layers):
                hh = load_pretrained_block(path, i, torch_dtype=torch.bfloat16).to("cuda:0").forward_cached(hh, None, None, 0)
            hh.float().pow(2).sum().backward()
            bwd = {"backward_rel_err": round(((g_tp - x.grad.float()).abs().mean() / (x.grad.float().abs().mean() + 1e-9)).item(), 5)}
        err = (got - ref).abs().mean().item() / (ref.abs().mean().item() + 1e-9)
        err_f = (fwd - ref).abs().mean().item() / (ref.abs().mean().item() + 1e-9)
        agree = (got.argmax(-1) == ref.argmax(-1)).float().mean().item()
        ok = err < 0.05 and err_f < 0.05 and agree > 0.9 and all(0 <= t < config.vocab_size for t in out[0].tolist())
        ok = ok and bwd.get("backward_rel_err", 0.0) < 0.08
        report = {"tp_server_selftest": "ok" if ok else "FAILED", "world": world, "rel_err": round(err, 5), "forward_rel_err": round(err_f, 5),
                  "argmax_agreement": round(agree, 4), "generated": out[0, 8:].tolist(), **bwd}
    finally:
        server.shutdown()
    print(json.dumps(report))
    if not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
