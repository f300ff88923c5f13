"""Text generation over remote blocks (reference: src/petals/client/remote_generation.py:1-164).

The reference wraps ``transformers.GenerationMixin.generate`` and smuggles a fake ``Cache`` object through it; that
coupling is what pins it to transformers 4.43. Here the decoding loop is self-contained —
greedy, temperature / top-k / top-p sampling with repetition penalty, and beam search — and keeps the reference's
*session* semantics:

* without an active session, one is created for the call and ``max_length`` xor ``max_new_tokens`` is required
  (``session max_length = pre_seq_len + prompt_len + max_new_tokens``);
* ``generate(..., session=sess)`` or calling inside ``with model.inference_session(...)`` continues the same
  server-side KV cache across calls; ``session.output_ids`` accumulates everything generated so far and new
  ``inputs`` are appended to it; the last generated token has not been fed yet and is sent first on the next call;
* beam search re-orders the server-side caches through ``hypo_ids`` carried by :class:`RemotePastKeyValues`.

On a CUDA client the per-token path is: embedding gather -> stage graph replays -> fused final-norm + LM-head GEMV ->
arg-max kernel, with a single 8-byte device->host read per token (the EOS check).
"""
from __future__ import annotations

import contextlib
import dataclasses
from typing import List, Optional, Union

import torch
import torch.nn.functional as F

from petals_b200.client.inference_session import InferenceSession
from petals_b200.utils.logging import get_logger
from petals_b200.utils.misc import DUMMY

logger = get_logger(__name__)


class RemotePastKeyValues:
    """Stand-in for a KV cache object: the real cache lives on the stages. Counts seen tokens and carries the beam
    permutation for the next step (reference :20-41)."""

    def __init__(self) -> None:
        self._seen_tokens = 0
        self.hypo_ids: Optional[torch.LongTensor] = None

    def __getitem__(self, _index: int) -> List[torch.Tensor]:
        return [DUMMY]  # for code that probes past_key_values[0][0].shape

    def get_seq_length(self, layer_idx: Optional[int] = 0) -> int:
        return self._seen_tokens

    def get_max_length(self) -> Optional[int]:
        return None

    def update_seen(self, new_seen: int) -> None:
        self._seen_tokens += new_seen

    def reorder_cache(self, beam_idx: torch.LongTensor) -> None:
        self.hypo_ids = beam_idx


_SKIP = contextlib.nullcontext


def _apply_repetition_penalty(logits: torch.Tensor, ids: torch.Tensor, penalty: float) -> torch.Tensor:
    if penalty == 1.0:
        return logits
    score = torch.gather(logits, 1, ids)
    score = torch.where(score < 0, score * penalty, score / penalty)
    return logits.scatter(1, ids, score)


def _warp(logits: torch.Tensor, temperature: float, top_k: Optional[int], top_p: Optional[float]) -> torch.Tensor:
    if temperature is not None and temperature != 1.0:
        logits = logits / temperature
    if top_k is not None and 0 < top_k < logits.shape[-1]:
        kth = torch.topk(logits, top_k, dim=-1).values[..., -1, None]
        logits = logits.masked_fill(logits < kth, float("-inf"))
    if top_p is not None and 0.0 < top_p < 1.0:
        sorted_logits, sorted_idx = torch.sort(logits, descending=False, dim=-1)
        cum = sorted_logits.softmax(dim=-1).cumsum(dim=-1)
        remove = cum <= (1 - top_p)
        remove[..., -1:] = False  # keep at least the most likely token
        logits = logits.masked_fill(remove.scatter(1, sorted_idx, remove), float("-inf"))
    return logits


class RemoteGenerationMixin:
    """``generate()`` / ``inference_session()`` / ``use_session()`` for causal-LM shells."""

    def inference_session(self, **kwargs) -> InferenceSession:
        return self.layers.inference_session(**kwargs)

    def use_session(self, session: Optional[InferenceSession]):
        return self.layers.use_session(session)

    @property
    def active_session(self) -> Optional[InferenceSession]:
        return self.layers.active_session

    @torch.inference_mode()
    def generate(self, inputs: Optional[torch.Tensor] = None, *args, session: Optional[InferenceSession] = None,
                 input_ids: Optional[torch.Tensor] = None, max_length: Optional[int] = None, max_new_tokens: Optional[int] = None,
                 do_sample: Union[bool, int, None] = False, temperature: float = 1.0, top_k: Optional[int] = None,
                 top_p: Optional[float] = None, repetition_penalty: float = 1.0, num_beams: int = 1, num_return_sequences: int = 1,
                 length_penalty: float = 1.0, eos_token_id: Union[int, List[int], None] = None, pad_token_id: Optional[int] = None,
                 generator: Optional[torch.Generator] = None, attention_mask=None, **kwargs) -> torch.LongTensor:
        if args:
            raise TypeError("generate() takes at most one positional argument (inputs)")
        if inputs is None:
            inputs = input_ids
        elif input_ids is not None:
            raise ValueError("pass either `inputs` or `input_ids`, not both")
        if inputs is not None and (not isinstance(inputs, torch.Tensor) or inputs.dim() != 2 or inputs.dtype != torch.int64):
            raise ValueError("inputs must be an int64 tensor [batch, seq]")
        if inputs is not None and inputs.numel():
            lo, hi = int(inputs.min()), int(inputs.max())  # one host sync per call; an id outside the table would read out of bounds on the GPU
            if lo < 0 or hi >= self.config.vocab_size:
                raise ValueError(f"token ids must be within [0, {self.config.vocab_size}), got values in [{lo}, {hi}]")
        if max_length is not None and inputs is not None and session is None and max_length < inputs.shape[1]:
            raise ValueError(f"max_length={max_length} is shorter than the prompt ({inputs.shape[1]} tokens)")
        if max_new_tokens is not None and max_new_tokens < 0:
            raise ValueError("max_new_tokens must be >= 0")
        if num_beams < 1 or num_return_sequences < 1:
            raise ValueError("num_beams and num_return_sequences must be >= 1")
        do_sample = bool(do_sample)  # ints were accepted by older Petals releases (reference :157-160)
        if do_sample and not temperature > 0:
            raise ValueError("temperature must be positive when sampling (use do_sample=False for greedy decoding)")
        if top_p is not None and not 0 < top_p <= 1:
            raise ValueError("top_p must be in (0, 1]")
        if top_k is not None and top_k < 0:
            raise ValueError("top_k must be >= 0 (0 disables it)")
        if num_beams > 1 and do_sample:
            raise NotImplementedError("beam search with sampling is not supported")
        if num_return_sequences > num_beams and not do_sample and num_return_sequences > 1:
            raise ValueError("num_return_sequences must be <= num_beams for beam search")
        eos_ids = [] if eos_token_id is None else ([eos_token_id] if isinstance(eos_token_id, int) else list(eos_token_id))
        if eos_token_id is None and getattr(self.config, "eos_token_id", None) is not None and kwargs.get("use_config_eos", False):
            e = self.config.eos_token_id
            eos_ids = [e] if isinstance(e, int) else list(e)
        if pad_token_id is None:
            pad_token_id = getattr(self.config, "pad_token_id", None)
            if pad_token_id is None:
                pad_token_id = eos_ids[0] if eos_ids else 0
        pre_seq_len = getattr(self.model, "pre_seq_len", 0) or 0

        context = contextlib.ExitStack()
        with context:
            if session is not None:
                context.enter_context(self.use_session(session))
            elif self.active_session is not None:
                session = self.active_session
            else:
                if (max_length is None) == (max_new_tokens is None):
                    raise ValueError("You should set `max_length` or `max_new_tokens` (but not both) to reserve server-side attention caches")
                prompt_len = 0 if inputs is None else inputs.shape[1]
                session_max_length = pre_seq_len + (max_length if max_length is not None else prompt_len + max_new_tokens)
                session = context.enter_context(self.inference_session(max_length=session_max_length))

            # ---- assemble the full id history and find what still has to be fed -------------------------------
            if session.output_ids is not None:
                ids = session.output_ids if inputs is None else torch.cat([session.output_ids, inputs.to(session.output_ids.device)], dim=1)
            else:
                if inputs is None:
                    bos = getattr(self.config, "bos_token_id", None)
                    if bos is None:
                        raise ValueError("`inputs` is required (the config defines no bos_token_id to start from)")
                    inputs = torch.tensor([[bos]], dtype=torch.int64)
                ids = inputs
            dev = self.device
            ids = ids.to(dev)
            n_processed = max(session.position - pre_seq_len, 0) if session.position > 0 else 0
            if n_processed > ids.shape[1]:
                raise ValueError("the session is ahead of the provided token history")
            total_cap = None
            if max_length is not None:
                total_cap = max_length
            elif max_new_tokens is not None:
                total_cap = ids.shape[1] + max_new_tokens
            else:
                total_cap = session.max_length - pre_seq_len
            total_cap = min(total_cap, session.max_length - pre_seq_len + 1)
            if num_beams > 1:
                out = self._beam_search(session, ids, n_processed, total_cap, num_beams, num_return_sequences, length_penalty, eos_ids, pad_token_id)
            else:
                if num_return_sequences > 1:
                    if n_processed:
                        raise ValueError("num_return_sequences > 1 requires a fresh session")
                    ids = ids.repeat_interleave(num_return_sequences, dim=0)
                out = self._sample(session, ids, n_processed, total_cap, do_sample, temperature, top_k, top_p, repetition_penalty,
                                   eos_ids, pad_token_id, generator)
            session.output_ids = out
            return out

    # ---- decoding loops ------------------------------------------------------------------------------------------
    def _next_logits(self, ids: torch.Tensor, n_processed: int, past: RemotePastKeyValues) -> torch.Tensor:
        out = self(input_ids=ids[:, n_processed:], past_key_values=past)
        return out.logits[:, -1]

    def _sample(self, session, ids, n_processed, total_cap, do_sample, temperature, top_k, top_p, repetition_penalty, eos_ids,
                pad_token_id, generator) -> torch.Tensor:
        past = RemotePastKeyValues()
        B = ids.shape[0]
        unfinished = torch.ones(B, dtype=torch.bool, device=ids.device)
        eos_t = torch.tensor(eos_ids, device=ids.device) if eos_ids else None
        greedy_fast = not do_sample and repetition_penalty == 1.0
        while ids.shape[1] < total_cap:
            logits = self._next_logits(ids, n_processed, past)
            n_processed = ids.shape[1]
            if greedy_fast and logits.is_cuda and logits.dtype in (torch.bfloat16, torch.float32):
                from petals_b200.ops import functional as Fn

                nxt = Fn.argmax(logits)
            else:
                lf = _apply_repetition_penalty(logits.float(), ids, repetition_penalty)
                if do_sample:
                    probs = _warp(lf, temperature, top_k, top_p).softmax(-1)
                    nxt = torch.multinomial(probs, 1, generator=generator).squeeze(1)
                else:
                    nxt = lf.argmax(-1)
            if eos_t is not None:
                nxt = torch.where(unfinished, nxt, torch.full_like(nxt, pad_token_id))
            ids = torch.cat([ids, nxt[:, None]], dim=1)
            if eos_t is not None:
                unfinished = unfinished & ~torch.isin(nxt, eos_t)
                if not bool(unfinished.any()):
                    break
        return ids

    def _beam_search(self, session, ids, n_processed, total_cap, num_beams, num_return, length_penalty, eos_ids, pad_token_id) -> torch.Tensor:
        if n_processed:
            raise ValueError("beam search needs a fresh session (server-side caches are expanded per beam)")
        B, prompt_len = ids.shape
        dev = ids.device
        ids = ids.repeat_interleave(num_beams, dim=0)  # [B*nb, L]
        beam_scores = torch.zeros(B, num_beams, device=dev)
        beam_scores[:, 1:] = -1e9  # all beams start identical: only expand the first one at step 0
        beam_scores = beam_scores.view(-1)
        finished: List[List[tuple]] = [[] for _ in range(B)]
        done = [False] * B
        past = RemotePastKeyValues()
        eos_set = set(eos_ids)
        while ids.shape[1] < total_cap and not all(done):
            logits = self._next_logits(ids, n_processed, past).float()
            n_processed = ids.shape[1]
            logp = F.log_softmax(logits, dim=-1) + beam_scores[:, None]
            V = logp.shape[-1]
            top_scores, top_idx = torch.topk(logp.view(B, num_beams * V), 2 * num_beams, dim=1)
            src_beam, token = top_idx // V, top_idx % V
            new_ids, new_scores, new_src = [], [], []
            cur_len = ids.shape[1] + 1
            for b in range(B):
                chosen = []
                for rank in range(2 * num_beams):
                    tok, sc, sb = int(token[b, rank]), float(top_scores[b, rank]), int(src_beam[b, rank])
                    if tok in eos_set:
                        if rank < num_beams:
                            hyp = torch.cat([ids[b * num_beams + sb], torch.tensor([tok], device=dev)])
                            finished[b].append((sc / (cur_len - prompt_len) ** length_penalty, hyp))
                        continue
                    chosen.append((sb, tok, sc))
                    if len(chosen) == num_beams:
                        break
                while len(chosen) < num_beams:  # degenerate tiny vocabularies
                    chosen.append(chosen[-1] if chosen else (0, pad_token_id, -1e9))
                if len(finished[b]) >= num_beams:
                    best_open = max(c[2] for c in chosen) / (cur_len - prompt_len) ** length_penalty
                    worst_done = sorted(f[0] for f in finished[b])[-num_beams]
                    done[b] = done[b] or worst_done >= best_open
                for sb, tok, sc in chosen:
                    new_src.append(b * num_beams + sb)
                    new_ids.append(tok)
                    new_scores.append(sc)
            src = torch.tensor(new_src, device=dev, dtype=torch.int64)
            ids = torch.cat([ids[src], torch.tensor(new_ids, device=dev, dtype=torch.int64)[:, None]], dim=1)
            beam_scores = torch.tensor(new_scores, device=dev)
            past.reorder_cache(src)
        results = []
        for b in range(B):
            cands = list(finished[b])
            for k in range(num_beams):
                row = b * num_beams + k
                cands.append((float(beam_scores[row]) / max(ids.shape[1] - prompt_len, 1) ** length_penalty, ids[row]))
            cands.sort(key=lambda c: c[0], reverse=True)
            results.extend(h for _, h in cands[:num_return])
        L = max(h.shape[0] for h in results)
        out = torch.full((len(results), L), pad_token_id, dtype=torch.int64, device=dev)
        for i, h in enumerate(results):
            out[i, : h.shape[0]] = h
        return out
