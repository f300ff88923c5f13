"""Speculative generation for Llama (reference: src/petals/models/llama/speculative_model.py:13-111).

A small local draft model proposes ``speculative_chunk`` tokens greedily; the distributed model validates them in
ONE multi-token remote step (prefill-style, through the stage engine's flash-attention path); the longest matching
prefix is accepted plus one corrected token, and the server-side KV caches are rolled back to the last accepted
position by moving ``session.position`` (a block-table truncate on the stages — no cache copy). Greedy only, batch
size 1, like the reference."""
from __future__ import annotations

from typing import Optional

import torch

from petals_b200.client.remote_generation import RemotePastKeyValues
from petals_b200.models.llama.model import DistributedLlamaForCausalLM


class DistributedLlamaForSpeculativeGeneration(DistributedLlamaForCausalLM):
    def __init__(self, config, small_model=None, *, dht=None):
        super().__init__(config, dht=dht)
        self.small_model = small_model

    @classmethod
    def from_pretrained(cls, model_name_or_path, *args, small_model=None, **kwargs):
        model = super().from_pretrained(model_name_or_path, *args, **kwargs)
        model.small_model = small_model
        return model

    @torch.inference_mode()
    def generate(self, inputs: Optional[torch.Tensor] = None, *, max_new_tokens: Optional[int] = None, max_length: Optional[int] = None,
                 speculative_chunk: int = 10, session=None, eos_token_id=None, do_sample: bool = False, **kwargs) -> torch.LongTensor:
        if self.small_model is None:
            raise ValueError("a draft `small_model` is required for speculative generation")
        if do_sample:
            raise NotImplementedError("speculative generation supports greedy decoding only")
        if inputs is None or inputs.shape[0] != 1:
            raise ValueError("speculative generation expects a single prompt [1, seq]")
        if (max_length is None) == (max_new_tokens is None):
            raise ValueError("set exactly one of max_length / max_new_tokens")
        total = max_length if max_length is not None else inputs.shape[1] + max_new_tokens
        eos = set([] if eos_token_id is None else ([eos_token_id] if isinstance(eos_token_id, int) else eos_token_id))
        pre = getattr(self.model, "pre_seq_len", 0) or 0
        own_session = session is None and self.active_session is None
        ctx = self.inference_session(max_length=pre + total + speculative_chunk + 1) if own_session else (self.use_session(session) if session is not None else _null())
        with ctx as active:
            sess = active if active is not None else self.active_session
            ids = inputs.to(self.device)
            fed = 0  # tokens of `ids` whose KV is valid on the stages
            past = RemotePastKeyValues()
            while ids.shape[1] < total:
                k = min(speculative_chunk, total - ids.shape[1])
                draft = self._draft(ids, k)  # [1, k]
                cand = torch.cat([ids, draft], dim=1)
                # one remote step validates every draft token: logits[i] predicts cand[fed + i + 1]
                logits = self(input_ids=cand[:, fed:], past_key_values=past).logits
                pred = logits.argmax(-1)  # [1, L - fed]
                base = ids.shape[1] - fed - 1  # index in `pred` of the prediction for the first draft token
                verify = pred[0, base: base + k]
                match = (verify == draft[0]).long()
                n_ok = int(match.cumprod(0).sum())
                accepted = draft[:, :n_ok]
                correction = pred[:, base + n_ok: base + n_ok + 1]  # the model's own token after the accepted prefix
                ids = torch.cat([ids, accepted, correction], dim=1)[:, :total]
                # KV is valid for everything fed except rejected draft tokens; the correction token is not fed yet
                fed = min(cand.shape[1] - (k - n_ok), ids.shape[1] - 1)
                sess.position = pre + fed
                if eos:
                    new_from = ids.shape[1] - min(n_ok + 1, ids.shape[1])
                    hits = [i for i in range(new_from, ids.shape[1]) if int(ids[0, i]) in eos]
                    if hits:  # stop right after the first end-of-sequence token, like token-by-token decoding would
                        ids = ids[:, : hits[0] + 1]
                        sess.position = pre + min(fed, ids.shape[1] - 1)
                        break
            sess.output_ids = ids
            return ids

    def _draft(self, ids: torch.Tensor, k: int) -> torch.Tensor:
        sm = self.small_model
        dev = next(sm.parameters()).device if hasattr(sm, "parameters") else ids.device
        out = sm.generate(ids.to(dev), max_new_tokens=k, do_sample=False)
        return out[:, ids.shape[1]: ids.shape[1] + k].to(ids.device)


class _null:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False
