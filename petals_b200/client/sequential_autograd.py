"""Autograd through a chain of remote stages (reference: src/petals/client/sequential_autograd.py:1-277).

Forward: the batch is cut into micro-batches of at most ``MAX_TOKENS_IN_BATCH`` tokens which travel through the
chain **concurrently** (one worker thread each; the reference uses asyncio tasks) — with stages on different GPUs
this is the pipeline-parallel schedule: micro-batch i is on stage s+1 while micro-batch i+1 is on stage s, and each
stage's runtime orders them by priority/arrival. Every span's input is remembered. Backward walks the spans in
reverse; if a stage died in between, the forward of that sub-chain is re-run on a replacement route first. Returns
gradients w.r.t. the inputs and the deep prompts. Stages keep no activation state between the two calls
(stateless => fault tolerant), exactly like the reference."""
from __future__ import annotations

import itertools
import time
from collections import deque
from concurrent.futures import ThreadPoolExecutor
from typing import List, Optional, Sequence, Tuple

import torch

from petals_b200.client.remote_forward_backward import run_remote_backward, run_remote_forward
from petals_b200.client.routing import RemoteSequenceManager, maybe_log_traceback
from petals_b200.data_structures import RemoteSpanInfo
from petals_b200.utils.logging import get_logger
from petals_b200.utils.misc import DUMMY, is_dummy

logger = get_logger(__name__)
MAX_TOKENS_IN_BATCH = 1024
_executor = ThreadPoolExecutor(max_workers=16, thread_name_prefix="petals-autograd")


def sequential_forward(inputs: torch.Tensor, prompts: torch.Tensor, sequence_manager: RemoteSequenceManager,
                       start_index: int = 0, end_index: Optional[int] = None) -> Tuple[torch.Tensor, Sequence[torch.Tensor], Sequence[RemoteSpanInfo]]:
    """Forward through blocks [start_index, end_index); returns (outputs, per-span inputs, spans)."""
    assert isinstance(inputs, torch.Tensor) and inputs.ndim == 3, f"{type(inputs)}: {inputs.ndim}"
    inputs_device, inputs_dtype = inputs.device, inputs.dtype
    end_index = end_index if end_index is not None else len(sequence_manager.block_uids)
    assert start_index >= 0 and end_index <= len(sequence_manager.block_uids)
    assert is_dummy(prompts) or len(prompts) == len(sequence_manager.block_uids)
    sequences: deque = deque()
    intermediate_inputs, done_sequences = [], []
    block_idx = start_index
    while block_idx < end_index:
        for attempt_no in itertools.count():
            logger.debug(f"Forward: block {block_idx}, attempt {attempt_no}")
            span = None
            try:
                if not sequences or attempt_no >= 1:
                    sequences = deque(sequence_manager.make_sequence(block_idx, end_index, mode="max_throughput"))
                    logger.debug(f"Found path from block {block_idx} to {end_index} via {len(sequences)} servers")
                span = sequences.popleft()
                stub = sequence_manager.connect(span.peer_id)
                uids = sequence_manager.block_uids[span.start: span.end]
                span_prompts = prompts[span.start: span.end] if not is_dummy(prompts) else DUMMY
                metadata = sequence_manager.get_request_metadata("rpc_forward", None, *uids)
                (outputs,) = run_remote_forward(stub, uids, inputs, span_prompts, metadata=metadata, timeout=sequence_manager.config.request_timeout)
                assert outputs.shape == inputs.shape, f"Expected output {inputs.shape}, got {outputs.shape}"
                intermediate_inputs.append(inputs)
                done_sequences.append(span)
                inputs = outputs
                block_idx = span.end
                sequence_manager.on_request_success(span.peer_id)
                break
            except Exception as e:  # noqa: BLE001
                sequence_manager.on_request_failure(span.peer_id if span is not None else None)
                if sequence_manager.config.max_retries is not None and attempt_no + 1 >= sequence_manager.config.max_retries:
                    raise
                delay = sequence_manager.get_retry_delay(attempt_no)
                logger.warning(f"Caught exception when running forward via {span} (retry in {delay:.0f} sec): {e!r}")
                maybe_log_traceback(e)
                time.sleep(delay)
    outputs = inputs.to(device=inputs_device, dtype=inputs_dtype)
    intermediate_inputs = [t.to(device=inputs_device, dtype=inputs_dtype) for t in intermediate_inputs]
    return outputs, intermediate_inputs, done_sequences


def sequential_backward(grad_outputs: Sequence[torch.Tensor], intermediate_inputs: List[torch.Tensor], prompts: torch.Tensor,
                        forward_sequences: List[RemoteSpanInfo], sequence_manager: RemoteSequenceManager) -> Tuple[Sequence[torch.Tensor], torch.Tensor]:
    """Backward through the spans recorded by ``sequential_forward``; returns (grad_inputs, grad_prompts)."""
    assert len(intermediate_inputs) == len(forward_sequences)
    grad_outputs = list(grad_outputs)
    grad_device, grad_dtype = grad_outputs[0].device, grad_outputs[0].dtype
    grad_prompts_reversed = []
    while len(forward_sequences) > 0 and len(intermediate_inputs) > 0:
        inputs = intermediate_inputs.pop()
        span = forward_sequences.pop()
        for attempt_no in itertools.count():
            logger.debug(f"Backward: block {span.end - 1}, attempt {attempt_no}")
            try:
                if attempt_no >= 1:
                    # the stage that produced this activation is gone: recompute the sub-chain on a fresh route
                    _, backup_inputs, backup_sequences = sequential_forward(inputs, prompts, sequence_manager, start_index=span.start, end_index=span.end)
                    assert len(backup_inputs) == len(backup_sequences)
                    assert backup_sequences[0].start == span.start and backup_sequences[-1].end == span.end
                    intermediate_inputs.extend(backup_inputs)
                    forward_sequences.extend(backup_sequences)
                    inputs = intermediate_inputs.pop()
                    span = forward_sequences.pop()
                stub = sequence_manager.connect(span.peer_id)
                uids = sequence_manager.block_uids[span.start: span.end]
                span_prompts = prompts[span.start: span.end] if not is_dummy(prompts) else DUMMY
                metadata = sequence_manager.get_request_metadata("rpc_backward", None, *uids)
                grads = run_remote_backward(stub, uids, inputs, grad_outputs[0], span_prompts, metadata=metadata, timeout=sequence_manager.config.request_timeout)
                grad_outputs = [grads[0]]
                grad_prompts_reversed.extend(reversed(grads[1:]))
                sequence_manager.on_request_success(span.peer_id)
                break
            except Exception as e:  # noqa: BLE001
                sequence_manager.on_request_failure(span.peer_id if span is not None else None)
                if sequence_manager.config.max_retries is not None and attempt_no + 1 >= sequence_manager.config.max_retries:
                    raise
                delay = sequence_manager.get_retry_delay(attempt_no)
                logger.warning(f"Caught exception when running backward via {span} (retry in {delay:.0f} sec): {e!r}")
                maybe_log_traceback(e)
                time.sleep(delay)
    grad_prompts = [g.to(device=grad_device, dtype=grad_dtype) for g in grad_prompts_reversed[::-1]]
    grad_prompts = torch.cat(grad_prompts, dim=0) if grad_prompts else DUMMY
    return [g.to(device=grad_device, dtype=grad_dtype) for g in grad_outputs], grad_prompts


def _gather_forward(input_batches, prompt_batches, sequence_manager):
    """Run every micro-batch through the chain concurrently."""
    futures = [_executor.submit(sequential_forward, x, p, sequence_manager) for x, p in zip(input_batches, prompt_batches)]
    return [f.result() for f in futures]


def _gather_backward(grad_output_batches, intermediate_input_batches, prompt_batches, forward_sequences, sequence_manager):
    """Backward of every micro-batch, in the calling thread.

    This runs inside ``torch.autograd``'s backward pass — for CUDA tensors that is the device's autograd worker
    thread. An in-process stage computes its own backward with a (re-entrant) autograd call, which must therefore be
    issued from this very thread: handing it to another thread would queue it behind the worker we are blocking."""
    return [sequential_backward((g,), inp, p, spans, sequence_manager)
            for g, inp, p, spans in zip(grad_output_batches, intermediate_input_batches, prompt_batches, forward_sequences)]


class _RemoteSequentialAutogradFunction(torch.autograd.Function):
    """Differentiable w.r.t. inputs and deep prompts; weights live on the stages and are frozen."""

    @staticmethod
    def forward(ctx, inputs: torch.Tensor, prompts: torch.Tensor, sequence_manager: RemoteSequenceManager):
        batch_size = max(MAX_TOKENS_IN_BATCH // inputs.shape[1], 1)
        input_batches: Sequence[torch.Tensor] = inputs.detach().split(batch_size)
        if prompts is None or is_dummy(prompts):
            prompt_batches = [DUMMY] * len(input_batches)
        elif prompts.shape[1] == 1:
            prompt_batches = [prompts.detach()] * len(input_batches)
        else:
            prompt_batches = prompts.detach().split(batch_size, dim=1)
        sequence_manager.rpc_info  # noqa: B018 - make sure the route/schema is resolved before fanning out
        outputs = _gather_forward(input_batches, prompt_batches, sequence_manager)
        assert len(outputs) == len(input_batches)
        output_batches = [output[0] for output in outputs]
        ctx.prompt_batches = prompt_batches
        ctx.sequence_manager = sequence_manager
        ctx.intermediate_input_batches = [output[1] for output in outputs]
        ctx.sequences_for_batches = [output[2] for output in outputs]
        ctx.prompts_broadcast = prompts is not None and not is_dummy(prompts) and prompts.shape[1] == 1
        return torch.cat(output_batches, dim=0)

    @staticmethod
    def backward(ctx, grad_outputs: torch.Tensor):
        intermediate_input_batches: List[List[torch.Tensor]] = ctx.intermediate_input_batches
        forward_sequences: List[List[RemoteSpanInfo]] = ctx.sequences_for_batches
        ctx.sequence_manager.rpc_info  # noqa: B018
        batch_size = max(MAX_TOKENS_IN_BATCH // grad_outputs.shape[1], 1)
        grad_output_batches: Sequence[torch.Tensor] = grad_outputs.split(batch_size)
        assert len(intermediate_input_batches) == len(grad_output_batches) == len(forward_sequences)
        outputs = _gather_backward(grad_output_batches, intermediate_input_batches, ctx.prompt_batches, forward_sequences, ctx.sequence_manager)
        grad_input_batches = [output[0][0] for output in outputs]
        grad_prompt_batches = [output[1] for output in outputs]
        grad_inputs = torch.cat(grad_input_batches, dim=0)
        dummy_grad_prompts = [is_dummy(g) for g in grad_prompt_batches]
        if all(dummy_grad_prompts):
            grad_prompts = None
        elif ctx.prompts_broadcast:
            grad_prompts = torch.stack(grad_prompt_batches, 0).sum(0)
        else:
            grad_prompts = torch.cat(grad_prompt_batches, dim=1)
        return (grad_inputs, grad_prompts, None)
