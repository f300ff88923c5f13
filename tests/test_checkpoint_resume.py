"""Checkpoint / resume of the client-side trainable state.

The reference has no training-checkpoint subsystem of its own: the prompt embeddings and the classification head are
ordinary client parameters, saved with ``save_pretrained`` and allowed to be absent at load (reference
src/petals/client/ptune.py:22).  The contract checked here: train a few steps against a swarm -> ``save_pretrained`` ->
``from_pretrained`` of the saved directory gives bit-identical trainable tensors, identical outputs, and the same
optimiser trajectory when training continues; the remote blocks are never written.
"""
import json
import os

import pytest
import torch

from petals_b200.utils.auto_config import AutoDistributedModelForCausalLM, AutoDistributedModelForSequenceClassification
from petals_b200.utils.safetensors_io import SafetensorsFile
from tests.utils import checkpoint, swarm_of


@pytest.fixture(scope="module")
def served():
    path = checkpoint("llama")
    with swarm_of(path, ["0:2", "2:4"]) as (swarm, servers):
        yield path, swarm


def _train_steps(model, ids, labels, n, lr=1e-2):
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.SGD(params, lr=lr)
    losses = []
    for _ in range(n):
        out = model(input_ids=ids, labels=labels)
        opt.zero_grad()
        out.loss.backward()
        opt.step()
        losses.append(out.loss.item())
    return losses


@pytest.mark.parametrize("mode", ["ptune", "deep_ptune"])
def test_prompt_tuning_state_round_trips(served, tmp_path, mode):
    path, swarm = served
    torch.manual_seed(0)
    model = AutoDistributedModelForSequenceClassification.from_pretrained(path, initial_peers=swarm, tuning_mode=mode, pre_seq_len=3,
                                                                        num_labels=3)
    ids = torch.randint(0, model.config.vocab_size, (4, 6), generator=torch.Generator().manual_seed(1))
    labels = torch.tensor([0, 1, 2, 1])
    _train_steps(model, ids, labels, 2)

    ckpt = str(tmp_path / "ckpt")
    model.save_pretrained(ckpt)
    with open(os.path.join(ckpt, "config.json")) as f:
        saved_cfg = json.load(f)
    assert saved_cfg["tuning_mode"] == mode and saved_cfg["pre_seq_len"] == 3 and saved_cfg["dht_prefix"] == model.config.dht_prefix
    with SafetensorsFile(os.path.join(ckpt, "model.safetensors")) as f:
        keys = set(f.keys())
    assert {"model.prompt_embeddings.weight", "score.weight", "model.embed_tokens.weight", "model.norm.weight"} <= keys
    assert ("model.intermediate_prompt_embeddings.weight" in keys) == (mode == "deep_ptune")
    assert not any(".layers." in k for k in keys), "transformer blocks belong to the servers, not to a client checkpoint"

    resumed = AutoDistributedModelForSequenceClassification.from_pretrained(ckpt, initial_peers=swarm)
    assert resumed.config.tuning_mode == mode and resumed.num_labels == 3
    for (n1, p1), (n2, p2) in zip(model.named_parameters(), resumed.named_parameters()):
        assert n1 == n2 and p1.dtype == p2.dtype and torch.equal(p1, p2), n1
        assert p1.requires_grad == p2.requires_grad, n1
    with torch.no_grad():
        assert torch.equal(model(input_ids=ids).logits, resumed(input_ids=ids).logits)
    # continuing from the checkpoint follows the trajectory of the uninterrupted run
    assert _train_steps(model, ids, labels, 2) == pytest.approx(_train_steps(resumed, ids, labels, 2), rel=1e-5)


def test_trainable_state_keeps_fp32_through_a_low_precision_reload(served, tmp_path):
    path, swarm = served
    model = AutoDistributedModelForCausalLM.from_pretrained(path, initial_peers=swarm, tuning_mode="ptune", pre_seq_len=2)
    with torch.no_grad():
        model.model.prompt_embeddings.weight.copy_(torch.full_like(model.model.prompt_embeddings.weight, 1.0 + 2.0 ** -12))
    ckpt = str(tmp_path / "ckpt")
    model.save_pretrained(ckpt)
    resumed = AutoDistributedModelForCausalLM.from_pretrained(ckpt, initial_peers=swarm, torch_dtype=torch.bfloat16)
    w = resumed.model.prompt_embeddings.weight
    assert w.dtype == torch.float32 and torch.equal(w, model.model.prompt_embeddings.weight)  # 1 + 2^-12 is not a bf16 number
    assert resumed.model.embed_tokens.weight.dtype == torch.bfloat16


def test_missing_or_mismatched_trainable_state(served, tmp_path):
    path, swarm = served
    # a plain checkpoint has no prompts: they are freshly initialised (allowed), everything else loads
    model = AutoDistributedModelForCausalLM.from_pretrained(path, initial_peers=swarm, tuning_mode="ptune", pre_seq_len=4)
    assert model.model.prompt_embeddings.weight.shape == (4, model.config.hidden_size)
    ckpt = str(tmp_path / "ckpt")
    model.save_pretrained(ckpt)
    # asking for another prompt length than the one saved is an error, not a silent re-initialisation
    with pytest.raises(ValueError, match="pre_seq_len"):
        AutoDistributedModelForCausalLM.from_pretrained(ckpt, initial_peers=swarm, pre_seq_len=5)
    # dropping prompt tuning at load time simply ignores the saved prompts
    plain = AutoDistributedModelForCausalLM.from_pretrained(ckpt, initial_peers=swarm, tuning_mode=None)
    assert not hasattr(plain.model, "prompt_embeddings")
