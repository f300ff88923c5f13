"""BLOOM block (reference: src/petals/models/bloom/block.py:15-45).

The reference wraps HF's layer with a hand-optimised attention, three CUDA-graphed sub-ops and BLOOM<->Llama
cache layout conversions. Here the block is the generic oracle specialised by the BLOOM ``BlockSpec`` (ALiBi generated from positions inside the attention op, no mask tensors); on
a B200 the very same parameters are executed by the stage engine's fused kernels."""
import torch

from petals_b200.models.block_oracle import GenericBlock


class WrappedBloomBlock(GenericBlock):
    def __init__(self, config, layer_idx: int = 0, dtype: torch.dtype = torch.float32, device="cpu", init_std=None):
        super().__init__(config.block_spec(), dtype=dtype, device=device, init_std=init_std)
        self.layer_idx = layer_idx
