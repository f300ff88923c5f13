"""Block dtype resolution when loading from a checkpoint (reference: tests/test_dtype.py)."""
import pytest
import torch

from petals_b200.server.block_utils import resolve_block_dtype
from petals_b200.server.from_pretrained import load_pretrained_block
from petals_b200.utils.auto_config import AutoDistributedConfig
from tests.utils import checkpoint


@pytest.mark.parametrize("torch_dtype", [torch.float32, torch.float16, torch.bfloat16, "auto"])
@pytest.mark.parametrize("family", ["llama", "falcon"])
def test_block_dtype(family, torch_dtype):
    path = checkpoint(family)
    config = AutoDistributedConfig.from_pretrained(path)
    block = load_pretrained_block(path, 0, config=config, torch_dtype=torch_dtype)
    expected_dtype = resolve_block_dtype(config, torch_dtype)
    assert expected_dtype in (torch.float32, torch.float16, torch.bfloat16)
    assert all(param.dtype == expected_dtype for param in block.parameters())
    if torch_dtype != "auto":
        assert expected_dtype == torch_dtype
