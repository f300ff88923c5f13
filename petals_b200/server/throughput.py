"""Stage self-benchmark and advertised throughput (reference: src/petals/server/throughput.py:1-255).

Measures, on the actual device with the actual executor: ``inference_rps`` (1 token x n steps against a KV
cache), ``forward_rps`` (1024-token forward x n steps) — tokens/s per block — and, instead of an Internet
speed test, ``network_rps`` from the NVLink peer bandwidth (tokens/s = bytes/s / (hidden x dtype bytes)).
Results are cached in a json file under a file lock keyed by model / device / dtype / quantisation / TP."""
from __future__ import annotations

import fcntl
import json
import os
import time
from pathlib import Path
from typing import Dict, Optional, Sequence, Union

import torch

from petals_b200.server.block_utils import resolve_block_dtype
from petals_b200.utils.convert_block import QuantType
from petals_b200.utils.logging import get_logger
from petals_b200.utils.peaks import NVLINK_PEER_GBS

logger = get_logger(__name__)
DEFAULT_CACHE_DIR = os.getenv("PETALS_CACHE", str(Path(Path.home(), ".cache", "petals_b200")))
RELAY_PENALTY = 0.2  # kept for parity; there are no relays inside one box


def get_server_throughput(model_name: str, config, device: torch.device, dtype: Union[str, torch.dtype], *, num_blocks: int,
                          quant_type: QuantType = QuantType.NONE, tensor_parallel_devices: Sequence[torch.device] = (),
                          reachable_via_relay: bool = False, force_eval: bool = False, cache_dir: Optional[str] = None) -> Dict[str, float]:
    dtype = resolve_block_dtype(config, dtype)
    cache_dir = cache_dir or DEFAULT_CACHE_DIR
    os.makedirs(cache_dir, exist_ok=True)
    lock_path, cache_path = Path(cache_dir, "throughput.lock"), Path(cache_dir, "throughput_v1.json")
    key = f"model_{model_name}_device_{_device_name(device)}_dtype_{str(dtype).replace('torch.', '')}_quant_{quant_type.name.lower()}"
    if tensor_parallel_devices:
        key += f"_tp_{len(tensor_parallel_devices)}"
    with open(lock_path, "wb+") as lock_fd:
        fcntl.flock(lock_fd.fileno(), fcntl.LOCK_EX)
        cache = {}
        if cache_path.exists() and not force_eval:
            try:
                cache = json.loads(cache_path.read_text())
            except Exception:  # noqa: BLE001 - a corrupt cache is simply re-measured
                logger.warning("throughput cache is unreadable; re-measuring")
        if key not in cache:
            cache[key] = measure_throughput_info(config, device, dtype, quant_type=quant_type, tensor_parallel_devices=tensor_parallel_devices)
            cache_path.write_text(json.dumps(cache, indent=1))
    info = dict(cache[key])
    # a stage of n blocks spends on average (n + 1) / 2 blocks of compute per routed request (reference :96-106)
    avg_blocks = (num_blocks + 1) / 2
    network = info["network_rps"] * (RELAY_PENALTY if reachable_via_relay else 1.0)
    info["throughput"] = min(info["forward_rps"] / avg_blocks, network)
    return info


def _device_name(device: torch.device) -> str:
    device = torch.device(device)
    return torch.cuda.get_device_name(device) if device.type == "cuda" else "cpu"


def measure_throughput_info(config, device, dtype, *, quant_type: QuantType, tensor_parallel_devices: Sequence[torch.device] = ()) -> Dict[str, float]:
    """The generic (in-process) tensor-parallel split is measured as it will run; a stage served by the NVLink worker group is
    measured on its leader device alone (a lower bound: the group is faster), since the group does not exist yet at this point."""
    logger.info("Measuring this stage's throughput (a few seconds)")
    tp = tuple(tensor_parallel_devices) if tensor_parallel_devices and not all(torch.device(d).type == "cuda" for d in tensor_parallel_devices) else ()
    return dict(
        inference_rps=measure_compute_rps(config, device, dtype, quant_type=quant_type, tensor_parallel_devices=tp, n_tokens=1, n_steps=50, inference=True),
        forward_rps=measure_compute_rps(config, device, dtype, quant_type=quant_type, tensor_parallel_devices=tp,
                                        n_tokens=1024 if torch.device(device).type == "cuda" else 64, n_steps=5, inference=False),
        network_rps=measure_network_rps(config, dtype),
    )


def measure_network_rps(config, dtype: torch.dtype) -> float:
    """Tokens/s one NVLink direction can move between adjacent stages."""
    bytes_per_token = config.hidden_size * torch.finfo(dtype).bits / 8
    return NVLINK_PEER_GBS * 1e9 / bytes_per_token


def measure_compute_rps(config, device, dtype: torch.dtype, *, quant_type: QuantType = QuantType.NONE,
                        tensor_parallel_devices: Sequence[torch.device] = (), n_tokens: int, n_steps: int, inference: bool) -> float:
    """Tokens/s of ONE block on this device with the serving executor (random weights)."""
    from petals_b200.server.backend import Stage
    from petals_b200.server.block_utils import get_model_block
    from petals_b200.server.memory_cache import MemoryCache
    from petals_b200.server.stage_engine import fast_path_supported
    from petals_b200.utils.convert_block import convert_block

    device = torch.device(device)
    spec = config.block_spec()
    with torch.inference_mode(False):
        block = get_model_block(config, dtype=dtype, device=device)
        for p in block.parameters():
            p.data = torch.randn_like(p.data.float()).mul_(0.02).to(dtype) if p.dim() > 1 else torch.ones_like(p.data)
        block = convert_block(block, 0, config, tensor_parallel_devices or (device,), device, quant_type, freeze=True)
        paged = device.type == "cuda" and dtype == torch.bfloat16 and fast_path_supported(spec)
        max_len = max(n_steps * n_tokens if inference else n_tokens, 64)
        cache = MemoryCache(2 * max_len + 128, None, n_blocks=1, spec=spec, dtype=dtype, device=device, paged=paged, max_length=max_len)
        stage = Stage(config, [block], 0, device=device, memory_cache=cache, torch_dtype=dtype)
        dummy = torch.randn(1, n_tokens, config.hidden_size, device=device, dtype=dtype)
        elapsed = 0.0
        with cache.allocate_cache(1, max_len, None) as session, torch.no_grad():
            for step in range(n_steps + 1):
                start = time.perf_counter()
                if inference:
                    stage.inference_step(session, dummy)
                else:
                    stage.forward(dummy)
                if device.type == "cuda":
                    torch.cuda.synchronize(device)
                if step >= 1:  # first step is warm-up (graph capture etc.)
                    elapsed += time.perf_counter() - start
    rps = n_steps * n_tokens / max(elapsed, 1e-9)
    logger.info(f"{'Inference' if inference else 'Forward pass'} throughput: {rps:.1f} tokens/sec per block "
                f"({n_tokens} tokens/batch, {_device_name(device)}, {str(dtype).replace('torch.', '')}, quant {quant_type.name.lower()})")
    return rps
