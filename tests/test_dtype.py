"""The dtype a block is served in (reference: tests/test_dtype.py:10-16): an explicit ``torch_dtype`` wins; ``"auto"`` follows the
checkpoint's config except that fp32 checkpoints are served in bf16 (server/block_utils.py:resolve_block_dtype)."""
import pytest
import torch

from petals_b200.server.block_utils import resolve_block_dtype
from petals_b200.server.from_pretrained import load_pretrained_block
from petals_b200.utils.auto_config import AutoDistributedConfig
from tests.utils import checkpoint

EXPLICIT = [torch.float32, torch.float16, torch.bfloat16]


@pytest.fixture(scope="module", params=["llama", "falcon"])
def model(request):
    path = checkpoint(request.param)
    return path, AutoDistributedConfig.from_pretrained(path)


def _dtypes_of(block):
    return {p.dtype for p in block.parameters()}


@pytest.mark.parametrize("requested", EXPLICIT)
def test_explicit_dtype_is_honoured(model, requested):
    path, config = model
    assert resolve_block_dtype(config, requested) is requested
    assert _dtypes_of(load_pretrained_block(path, 0, config=config, torch_dtype=requested)) == {requested}


def test_auto_dtype_follows_the_checkpoint(model):
    path, config = model
    resolved = resolve_block_dtype(config, "auto")
    assert resolved in EXPLICIT
    assert _dtypes_of(load_pretrained_block(path, 0, config=config, torch_dtype="auto")) == {resolved}
    # the synthetic checkpoints are written in fp32: "auto" must not serve fp32 blocks
    if config.torch_dtype in (torch.float32, "float32", None):
        assert resolved == torch.bfloat16
