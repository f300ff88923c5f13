"""Stage worker: picks its block span, loads it, serves it, re-balances
(reference: src/petals/server/server.py:46-767).

One ``Server`` = one GPU (or a CPU device in tests). Kept from the reference: the constructor's parameter
names and defaults (``inference_max_length`` 8192 for GQA/MQA models else 2048, ``attn_cache_tokens``
16384/4096, ``max_chunk_size_bytes`` 256 MiB, timeouts, ``balance_quality`` 0.75, ``mean_balance_check_period``
...), automatic ``num_blocks`` from free device memory, throughput self-measurement, the JOINING -> ONLINE ->
OFFLINE announcements with expiring records, the health-check / re-balancing loop that tears the container down
and reloads other blocks, and ``--adapters``. Dropped: everything about NATs, relays, libp2p identities
(accepted and ignored for CLI compatibility).
"""
from __future__ import annotations

import gc
import math
import os
import random
import threading
import time
from typing import Dict, List, Optional, Sequence, Union

import torch

import petals_b200
from petals_b200.data_structures import ModelInfo, ServerInfo, ServerState, make_uid
from petals_b200.parallel.swarm import Swarm, get_dht_time, resolve_swarm
from petals_b200.server import block_selection
from petals_b200.server.backend import Stage, TransformerBackend, merge_inference_pools_inplace
from petals_b200.server.block_utils import get_block_size, resolve_block_dtype
from petals_b200.server.from_pretrained import load_pretrained_block
from petals_b200.server.handler import TransformerConnectionHandler
from petals_b200.server.memory_cache import MemoryCache
from petals_b200.server.task_pool import PrioritizedTaskPool, Runtime
from petals_b200.server.reachability import validate_direct_reachability
from petals_b200.server.throughput import get_server_throughput
from petals_b200.utils.auto_config import AutoDistributedConfig
from petals_b200.utils.convert_block import QuantType, check_device_balance, convert_block, resolve_quant_type
from petals_b200.utils.dht import declare_active_modules, get_remote_module_infos
from petals_b200.utils.logging import get_logger
from petals_b200.utils.ping import PingAggregator

logger = get_logger(__name__)
MAX_DHT_TIME_DISCREPANCY_SECONDS = 3.0
_peer_counter = [0]
_peer_lock = threading.Lock()


def _new_peer_id(device: torch.device) -> str:
    with _peer_lock:
        _peer_counter[0] += 1
        return f"{device.type}{device.index if device.index is not None else ''}-{os.getpid()}-{_peer_counter[0]}"


class Server:
    def __init__(self, *, initial_peers=None, dht_prefix: Optional[str] = None, converted_model_name_or_path: str,
                 public_name: Optional[str] = None, throughput: Union[float, str] = "auto", num_blocks: Optional[int] = None,
                 block_indices: Optional[str] = None, num_handlers: int = 8, inference_max_length: Optional[int] = None,
                 min_batch_size: int = 1, max_batch_size: Optional[int] = None, max_chunk_size_bytes: int = 256 * 1024 * 1024,
                 max_alloc_timeout: float = 600, attn_cache_tokens: Optional[int] = None, torch_dtype: str = "auto",
                 revision: Optional[str] = None, cache_dir: Optional[str] = None, max_disk_space: Optional[int] = None,
                 device: Optional[Union[str, torch.device]] = None, compression=None, stats_report_interval: Optional[int] = None,
                 custom_module_path=None, update_period: float = 60, expiration: Optional[float] = None,
                 request_timeout: float = 3 * 60, session_timeout: float = 30 * 60, step_timeout: float = 5 * 60,
                 prefetch_batches: int = 1, sender_threads: int = 1, balance_quality: float = 0.75,
                 mean_balance_check_period: float = 120, mean_block_selection_delay: float = 5, token=None,
                 quant_type: Optional[QuantType] = None, tensor_parallel_devices: Optional[Sequence[torch.device]] = None,
                 skip_reachability_check: bool = False, reachable_via_relay: Optional[bool] = None, use_relay: bool = True,
                 use_auto_relay: bool = True, adapters: Sequence[str] = (), peer_id: Optional[str] = None,
                 use_cuda_graphs: bool = True, force_oracle: bool = False, host_maddrs: Optional[Sequence[str]] = None,
                 announce_maddrs: Optional[Sequence[str]] = None, public_ip: Optional[str] = None, metrics_port: Optional[int] = None,
                 fabric_address: Optional[str] = None, fabric_rank: Optional[int] = None, fabric_world: Optional[int] = None,
                 fabric_max_tokens: int = 8192, **kwargs):
        if kwargs:
            logger.debug(f"ignoring networking options that have no meaning on one box: {sorted(kwargs)}")
        self.converted_model_name_or_path = converted_model_name_or_path
        self.num_handlers, self.compression = num_handlers, compression
        self.skip_reachability_check = skip_reachability_check
        # Prometheus endpoint (utils/metrics.py): survives re-balancing because it reads whichever container is live
        self.metrics_server = None
        if metrics_port is not None:
            from petals_b200.utils.metrics import MetricsServer

            self.metrics_server = MetricsServer(lambda: (lambda c: c.handler.metrics if c is not None else None)(getattr(self, "module_container", None)),
                                                metrics_port).start()
            logger.info(f"Serving metrics on http://0.0.0.0:{self.metrics_server.port}/metrics")
        self.stats_report_interval, self.update_period = stats_report_interval, update_period
        self.prefetch_batches, self.sender_threads = prefetch_batches, sender_threads
        self.revision, self.token, self.adapters = revision, token, tuple(adapters)
        self.use_cuda_graphs, self.force_oracle = use_cuda_graphs, force_oracle

        self.block_config = AutoDistributedConfig.from_pretrained(converted_model_name_or_path)
        if dht_prefix is None:
            dht_prefix = self.block_config.dht_prefix
        if "." in dht_prefix or " " in dht_prefix:
            raise ValueError(f"dht_prefix {dht_prefix!r} must not contain '.' or spaces")
        self.dht_prefix = dht_prefix
        self.expiration = expiration if expiration is not None else max(2 * update_period, MAX_DHT_TIME_DISCREPANCY_SECONDS)
        self.request_timeout, self.session_timeout, self.step_timeout = request_timeout, session_timeout, step_timeout
        self.module_uids = [make_uid(self.dht_prefix, i) for i in range(self.block_config.num_hidden_layers)]

        self.dht: Swarm = resolve_swarm(initial_peers)
        if hasattr(self.dht, "bind_host"):  # a network registry (parallel/registry.py): where to listen, what to announce
            from petals_b200.parallel.transport import parse_address

            if host_maddrs:
                self.dht.bind_host = parse_address(host_maddrs[0])[1]
            if announce_maddrs:
                self.dht.announce_host = parse_address(announce_maddrs[0])[1]
            elif public_ip:
                self.dht.announce_host = public_ip
        if device is None:
            device = "cuda" if torch.cuda.is_available() else "cpu"
        device = torch.device(device)
        if device.type == "cuda" and device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = device
        self.peer_id = peer_id or _new_peer_id(device)
        self.public_name = public_name

        if isinstance(torch_dtype, str):
            key = torch_dtype.replace("torch.", "")
            if key not in petals_b200.DTYPE_MAP:
                raise ValueError(f"torch_dtype must be one of {sorted(petals_b200.DTYPE_MAP)}, got {torch_dtype!r}")
            torch_dtype = petals_b200.DTYPE_MAP[key]
        torch_dtype = resolve_block_dtype(self.block_config, torch_dtype)
        if device.type == "cpu" and torch_dtype == torch.float16:
            raise ValueError("float16 is not supported on CPU; use float32 or bfloat16")
        self.torch_dtype = torch_dtype
        self.tensor_parallel_devices = tuple(torch.device(d) for d in tensor_parallel_devices) if tensor_parallel_devices else (device,)
        if len(self.tensor_parallel_devices) > 1:
            check_device_balance(self.tensor_parallel_devices)
        self.quant_type = resolve_quant_type(quant_type) if quant_type is not None else QuantType.NONE
        self._owns_process_group = False
        self.fabric = self._join_fabric(fabric_address, fabric_rank, fabric_world, fabric_max_tokens)

        spec = self.block_config.block_spec()
        is_multiquery_attn = spec.num_kv_heads < spec.num_heads
        if max_batch_size is None:
            max_batch_size = 8192 if is_multiquery_attn else 2048
        if inference_max_length is None:
            inference_max_length = 8192 if is_multiquery_attn else 2048
        if spec.rotary and inference_max_length > spec.max_position:
            # positions beyond the rotary table would be rotated like its last row: cap the session length the server accepts instead
            logger.warning(f"inference_max_length={inference_max_length} exceeds the {spec.max_position} positions of this model's rotary embedding "
                           f"(max_position_embeddings); serving sessions of up to {spec.max_position} tokens")
            inference_max_length = spec.max_position
        self.min_batch_size, self.max_batch_size, self.inference_max_length = min_batch_size, max_batch_size, inference_max_length
        self.max_chunk_size_bytes, self.max_alloc_timeout = max_chunk_size_bytes, max_alloc_timeout
        if attn_cache_tokens is None:
            attn_cache_tokens = 16384 if is_multiquery_attn else 4096
        if attn_cache_tokens < 1:
            raise ValueError(f"attn_cache_tokens must be positive, got {attn_cache_tokens}")
        self.attn_cache_tokens = attn_cache_tokens
        self.cache_bytes_per_block = attn_cache_tokens * spec.kv_bytes_per_token(torch_dtype)

        assert num_blocks is None or block_indices is None, "Please specify num_blocks or block_indices, not both"
        if block_indices is not None:
            try:
                start_block, end_block = [int(x.strip()) for x in block_indices.split(":")]
            except Exception as e:
                raise ValueError(f"block_indices must be 'start:end', got {block_indices!r}") from e
            if not 0 <= start_block < end_block <= self.block_config.num_hidden_layers:
                raise ValueError(f"block_indices {block_indices} outside the model's {self.block_config.num_hidden_layers} blocks")
            block_indices = list(range(start_block, end_block))
            num_blocks = len(block_indices)
        self.strict_block_indices = block_indices
        self.num_blocks = num_blocks if num_blocks is not None else self._choose_num_blocks()
        if not 1 <= self.num_blocks <= self.block_config.num_hidden_layers:
            raise ValueError(f"num_blocks={self.num_blocks} must be between 1 and the model's {self.block_config.num_hidden_layers} blocks")

        if throughput in ("auto", "eval", "dry_run"):
            info = get_server_throughput(converted_model_name_or_path, self.block_config, device, torch_dtype, num_blocks=self.num_blocks,
                                         quant_type=self.quant_type, tensor_parallel_devices=self.tensor_parallel_devices[1:] and self.tensor_parallel_devices,
                                         force_eval=throughput in ("eval", "dry_run"), cache_dir=cache_dir)
            if throughput == "dry_run":
                logger.info("dry_run: throughput measured, exiting")
                raise SystemExit(0)
        else:
            info = {"throughput": float(throughput)}
        self.server_info = ServerInfo(state=ServerState.JOINING, public_name=public_name, version=petals_b200.__version__,
                                      adapters=tuple(adapters), torch_dtype=str(torch_dtype).replace("torch.", ""),
                                      quant_type=self.quant_type.name.lower(), using_relay=False, **{k: float(v) for k, v in info.items()})
        self.model_info = ModelInfo(num_blocks=self.block_config.num_hidden_layers, repository=str(converted_model_name_or_path))
        self.balance_quality = balance_quality
        self.mean_balance_check_period, self.mean_block_selection_delay = mean_balance_check_period, mean_block_selection_delay
        self.module_container: Optional[ModuleContainer] = None
        self.stop = threading.Event()
        self._thread: Optional[threading.Thread] = None

    # ---- capacity planning (reference :275-326) ------------------------------------------------------------------
    def _choose_num_blocks(self) -> int:
        if self.device.type != "cuda":
            return min(self.block_config.num_hidden_layers, 4)
        free, total = torch.cuda.mem_get_info(self.device)
        block_size = get_block_size(self.block_config, "memory", dtype=self.torch_dtype, quant_type=self.quant_type)
        # activation / autograd reserve scales with the hidden size like the reference's 2 GiB @ 14336
        reserve = int(2 * 2**30 * self.block_config.hidden_size / 14336) + 2**30
        per_block = block_size + self.cache_bytes_per_block
        n = min(int((free - reserve) // per_block), self.block_config.num_hidden_layers)
        if n < 1:
            raise RuntimeError(f"not enough free GPU memory for a single block ({block_size / 2**30:.1f} GiB + cache)")
        logger.info(f"Server will fill this GPU with {n} transformer blocks")
        return n

    # ---- main loop (reference :328-384) ----------------------------------------------------------------------------
    def run(self) -> None:
        while not self.stop.is_set():
            block_indices = self._choose_blocks()
            self.module_container = ModuleContainer.create(
                dht=self.dht, dht_prefix=self.dht_prefix, converted_model_name_or_path=self.converted_model_name_or_path,
                block_config=self.block_config, attn_cache_tokens=self.attn_cache_tokens, server_info=self.server_info,
                model_info=self.model_info, block_indices=block_indices, min_batch_size=self.min_batch_size,
                max_batch_size=self.max_batch_size, max_chunk_size_bytes=self.max_chunk_size_bytes,
                max_alloc_timeout=self.max_alloc_timeout, inference_max_length=self.inference_max_length,
                torch_dtype=self.torch_dtype, device=self.device, quant_type=self.quant_type,
                tensor_parallel_devices=self.tensor_parallel_devices, adapters=self.adapters, update_period=self.update_period,
                expiration=self.expiration, request_timeout=self.request_timeout, session_timeout=self.session_timeout,
                step_timeout=self.step_timeout, stats_report_interval=self.stats_report_interval, peer_id=self.peer_id,
                use_cuda_graphs=self.use_cuda_graphs, force_oracle=self.force_oracle)
            self.module_container.handler.compression = self.compression
            try:
                self.module_container.ready.wait()
                if not self.skip_reachability_check:
                    validate_direct_reachability(self.dht, self.peer_id)
                from petals_b200.utils.version import validate_version

                validate_version(self.dht, self.module_uids)  # newer peers in the swarm?
                while not self.stop.is_set():
                    timeout = random.random() * 2 * self.mean_balance_check_period
                    if self.stop.wait(timeout):
                        return
                    self.module_container.handler.sweep_sessions()
                    if not self.module_container.is_healthy():
                        logger.warning("One of the subsystems died, restarting the stage")
                        break
                    if self._should_choose_other_blocks():
                        logger.info("Swarm is imbalanced, this stage will load other blocks")
                        break
            finally:
                self.module_container.shutdown()
                self.module_container = None
            self._clean_memory()

    def run_in_background(self, await_ready: bool = True, timeout: Optional[float] = None) -> None:
        self._thread = threading.Thread(target=self.run, name=f"server-{self.peer_id}", daemon=True)
        self._thread.start()
        if await_ready:
            deadline = None if timeout is None else time.monotonic() + timeout
            while self.module_container is None or not self.module_container.ready.is_set():
                if not self._thread.is_alive():
                    raise RuntimeError("server thread died during start-up")
                if deadline is not None and time.monotonic() > deadline:
                    raise TimeoutError("server did not become ready in time")
                time.sleep(0.01)

    def _clean_memory(self) -> None:
        gc.collect()
        if self.device.type == "cuda":
            torch.cuda.empty_cache()

    def _choose_blocks(self) -> List[int]:
        if self.strict_block_indices is not None:
            return self.strict_block_indices
        # jitter so that stages starting together see each other's JOINING records (reference :407-409)
        delay = math.sqrt(-2 * math.log(max(random.random(), 1e-9))) * self.mean_block_selection_delay if self.mean_block_selection_delay > 0 else 0
        if delay:
            self.stop.wait(min(delay, 2 * self.mean_block_selection_delay))
        module_infos = get_remote_module_infos(self.dht, self.module_uids, latest=True)
        return block_selection.choose_best_blocks(self.num_blocks, module_infos)

    def _should_choose_other_blocks(self) -> bool:
        if self.strict_block_indices is not None:
            return False
        module_infos = get_remote_module_infos(self.dht, self.module_uids, latest=True)
        return block_selection.should_choose_other_blocks(self.peer_id, module_infos, self.balance_quality)

    def _join_fabric(self, address: Optional[str], rank: Optional[int], world: Optional[int], max_tokens: int):
        """``--fabric_address HOST:PORT --fabric_rank R --fabric_world N``: the N stage processes of one NVLink box form a landing-ring
        fabric (parallel/fabric.py) — between them hidden states, training micro-batches and gradients hop GPU to GPU instead of
        travelling with the RPCs, for any client (a client needs no membership; it learns from ``rpc_info`` which stages share a
        fabric). Collective: returns when all N processes have joined. Without the flags, under ``torchrun`` (RANK / WORLD_SIZE /
        MASTER_ADDR / MASTER_PORT in the environment) the same happens with the launcher's values when ``PETALS_B200_FABRIC=env``."""
        from petals_b200.parallel.fabric import get_fabric, join_fabric

        if address is None and os.environ.get("PETALS_B200_FABRIC", "") == "env" and "RANK" in os.environ and "WORLD_SIZE" in os.environ:
            address = f"{os.environ.get('MASTER_ADDR', '127.0.0.1')}:{os.environ.get('MASTER_PORT', '29500')}"
            rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
        if address is None:
            return get_fabric()  # possibly created by the embedding process (benchmarks, self-tests)
        if rank is None or world is None or not 0 <= rank < world or world < 2:
            raise ValueError("--fabric_address needs --fabric_rank R and --fabric_world N with 0 <= R < N and N >= 2")
        if len(self.tensor_parallel_devices) > 1:
            raise ValueError("a tensor-parallel server owns its own process group: it cannot also join a box fabric (use one GPU per fabric member)")
        fabric, self._owns_process_group = join_fabric(address, rank, world, self.block_config.hidden_size, device=self.device, max_tokens=max_tokens,
                                                       host_dtype=self.torch_dtype if self.device.type == "cpu" else torch.float32)
        return fabric

    def shutdown(self, timeout: Optional[float] = 5) -> None:
        self.stop.set()
        if self.metrics_server is not None:
            self.metrics_server.shutdown()
            self.metrics_server = None
        if self._thread is not None and self._thread.is_alive() and threading.current_thread() is not self._thread:
            self._thread.join(timeout)
        if self.module_container is not None:
            self.module_container.shutdown()
            self.module_container = None
        if self._owns_process_group:
            from petals_b200.parallel.fabric import leave_fabric

            leave_fabric(True)
            self._owns_process_group = False


class ModuleContainer:
    """Everything one loaded span needs: blocks, stage, pools + runtime, handler, announcer."""

    @classmethod
    def create(cls, *, dht: Swarm, dht_prefix: str, converted_model_name_or_path: str, block_config, attn_cache_tokens: int,
               server_info: ServerInfo, model_info: ModelInfo, block_indices: List[int], min_batch_size: int, max_batch_size: int,
               max_chunk_size_bytes: int, max_alloc_timeout: float, inference_max_length: int, torch_dtype: torch.dtype,
               device: torch.device, quant_type: QuantType, tensor_parallel_devices: Sequence[torch.device], adapters: Sequence[str],
               update_period: float, expiration: float, request_timeout: float, session_timeout: float, step_timeout: float,
               stats_report_interval: Optional[float], peer_id: str, use_cuda_graphs: bool = True, force_oracle: bool = False,
               prebuilt_blocks: Optional[Sequence] = None) -> "ModuleContainer":
        module_uids = [make_uid(dht_prefix, i) for i in block_indices]
        server_info.start_block, server_info.end_block = block_indices[0], block_indices[-1] + 1
        announcer = ModuleAnnouncerThread(module_uids, dht, server_info, model_info, peer_id=peer_id, block_config=block_config,
                                          update_period=update_period, expiration=expiration, daemon=True)
        announcer.announce(ServerState.JOINING)
        announcer.start()
        logger.info(f"Announced that blocks {block_indices[0]}:{block_indices[-1] + 1} are joining")
        tp_group = None
        try:
            tp_world = len(tensor_parallel_devices)
            use_tp_engine = False
            if tp_world > 1 and device.type == "cuda" and torch_dtype == torch.bfloat16 and all(d.type == "cuda" for d in tensor_parallel_devices):
                from petals_b200.parallel.tensor_parallel import tp_supported

                use_tp_engine = tp_supported(block_config.block_spec(), tp_world)
            if use_tp_engine:
                # one stage = a tensor-parallel group of worker processes (reference: --tensor_parallel_devices, run_server.py:154-157):
                # this process leads, every other device gets a worker; the weights live in the group's shards
                from petals_b200.parallel.tp_worker import TPGroup

                if quant_type != QuantType.NONE or adapters or prebuilt_blocks is not None:
                    raise ValueError("tensor-parallel stages serve bf16 checkpoints without adapters; use pipeline stages for quantised / LoRA serving")
                tp_group = TPGroup(block_config, converted_model_name_or_path, block_indices, tensor_parallel_devices, torch_dtype=torch_dtype,
                                   attn_cache_tokens=attn_cache_tokens, inference_max_length=inference_max_length, use_cuda_graphs=use_cuda_graphs)
                blocks = [torch.nn.Identity() for _ in block_indices]
                stage = Stage(block_config, blocks, block_indices[0], device=device, memory_cache=tp_group.cache, torch_dtype=torch_dtype,
                              engine=tp_group.leader)
            else:
                blocks = []
                for i, block_index in enumerate(block_indices):
                    if prebuilt_blocks is not None:
                        block = prebuilt_blocks[i]
                    else:
                        block = load_pretrained_block(converted_model_name_or_path, block_index, config=block_config, torch_dtype=torch_dtype)
                    block = convert_block(block, block_index, block_config, tensor_parallel_devices, device, quant_type, freeze=True, adapters=adapters)
                    blocks.append(block)
                spec = block_config.block_spec()
                from petals_b200.server.stage_engine import fast_path_supported

                if tp_world > 1:
                    logger.info(f"Blocks are split over {[str(d) for d in tensor_parallel_devices]} by the generic tensor-parallel path "
                                f"(parallel/tp_generic.py); the NVLink engine shards bf16 Llama-style blocks on CUDA devices")
                paged = (device.type == "cuda" and torch_dtype == torch.bfloat16 and fast_path_supported(spec) and not force_oracle and tp_world == 1
                         and quant_type in (QuantType.NONE, QuantType.FP8) and not (quant_type == QuantType.FP8 and adapters))
                memory_cache = MemoryCache(attn_cache_tokens, max_alloc_timeout, n_blocks=len(blocks), spec=spec, dtype=torch_dtype, device=device,
                                           paged=paged, max_length=inference_max_length)
                stage = Stage(block_config, blocks, block_indices[0], device=device, memory_cache=memory_cache, torch_dtype=torch_dtype,
                              max_chunk_size_bytes=max_chunk_size_bytes, use_cuda_graphs=use_cuda_graphs, force_oracle=force_oracle,
                              fp8=(quant_type == QuantType.FP8 and paged))
            container_parts = cls._build_pools(stage, module_uids, blocks, max_batch_size, stats_report_interval, peer_id, device)
        except BaseException:
            announcer.announce(ServerState.OFFLINE)
            announcer.stop.set()
            if tp_group is not None:
                tp_group.shutdown()
            raise
        backends, runtime, inference_pool, forward_pool, backward_pool = container_parts
        return cls(dht, dht_prefix, backends, stage=stage, runtime=runtime, inference_pool=inference_pool, forward_pool=forward_pool,
                   backward_pool=backward_pool, announcer=announcer, peer_id=peer_id, adapters=adapters,
                   inference_max_length=inference_max_length, request_timeout=request_timeout, session_timeout=session_timeout,
                   step_timeout=step_timeout, quant_type=quant_type, tp_group=tp_group)

    @staticmethod
    def _build_pools(stage: Stage, module_uids: List[str], blocks, max_batch_size: int, stats_report_interval, peer_id: str, device):
        runtime = Runtime(name=f"runtime-{peer_id}", stats_report_interval=stats_report_interval, device=device)
        backends: Dict[str, TransformerBackend] = {}
        for slot, (uid, block) in enumerate(zip(module_uids, blocks)):
            backends[uid] = TransformerBackend(uid, block, stage=stage, slot=slot, max_batch_size=max_batch_size, runtime=runtime)
        inference_pool = merge_inference_pools_inplace(backends, stage, runtime, max_batch_size)

        def span_forward(hidden, prompts, lo, hi, active_adapter, hops=None):
            stage.use_adapter(active_adapter)
            return stage.forward(hidden, prompts, lo, hi, **(hops or {}))

        def span_backward(hidden, grad, prompts, lo, hi, active_adapter, hops=None):
            stage.use_adapter(active_adapter)
            return stage.backward(hidden, grad, prompts, lo, hi, **(hops or {}))

        forward_pool = PrioritizedTaskPool(span_forward, max_batch_size, "span_forward", runtime)
        backward_pool = PrioritizedTaskPool(span_backward, max_batch_size, "span_backward", runtime, in_caller_thread=True)
        return backends, runtime, inference_pool, forward_pool, backward_pool

    @classmethod
    def from_stage(cls, *, dht: Swarm, dht_prefix: str, block_config, stage: Stage, server_info: ServerInfo, model_info: ModelInfo,
                   peer_id: str, max_batch_size: int = 1 << 20, inference_max_length: int = 8192, update_period: float = 30,
                   expiration: float = 3600, request_timeout: float = 180, session_timeout: float = 1800, step_timeout: float = 300,
                   adapters: Sequence[str] = (), quant_type: QuantType = QuantType.NONE) -> "ModuleContainer":
        """Serve an already constructed stage (e.g. the leader of a tensor-parallel worker group)."""
        block_indices = list(range(stage.start_block, stage.end_block))
        module_uids = [make_uid(dht_prefix, i) for i in block_indices]
        server_info.start_block, server_info.end_block = block_indices[0], block_indices[-1] + 1
        announcer = ModuleAnnouncerThread(module_uids, dht, server_info, model_info, peer_id=peer_id, block_config=block_config,
                                          update_period=update_period, expiration=expiration, daemon=True)
        announcer.announce(ServerState.JOINING)
        announcer.start()
        parts = cls._build_pools(stage, module_uids, stage.blocks, max_batch_size, None, peer_id, stage.device)
        backends, runtime, inference_pool, forward_pool, backward_pool = parts
        return cls(dht, dht_prefix, backends, stage=stage, runtime=runtime, inference_pool=inference_pool, forward_pool=forward_pool,
                   backward_pool=backward_pool, announcer=announcer, peer_id=peer_id, adapters=adapters,
                   inference_max_length=inference_max_length, request_timeout=request_timeout, session_timeout=session_timeout,
                   step_timeout=step_timeout, quant_type=quant_type)

    def __init__(self, dht: Swarm, dht_prefix: str, module_backends: Dict[str, TransformerBackend], *, stage: Stage, runtime: Runtime,
                 inference_pool, forward_pool, backward_pool, announcer: "ModuleAnnouncerThread", peer_id: str, adapters,
                 inference_max_length: int, request_timeout: float, session_timeout: float, step_timeout: float, quant_type, tp_group=None):
        self.dht, self.dht_prefix, self.module_backends, self.stage, self.runtime = dht, dht_prefix, module_backends, stage, runtime
        self.tp_group = tp_group
        self.announcer, self.peer_id = announcer, peer_id
        self.handler = TransformerConnectionHandler(
            dht, module_backends, stage=stage, peer_id=peer_id, inference_pool=inference_pool, forward_pool=forward_pool,
            backward_pool=backward_pool, adapters=adapters, inference_max_length=inference_max_length, request_timeout=request_timeout,
            session_timeout=session_timeout, step_timeout=step_timeout, quant_type=quant_type)
        self.announcer.memory_cache = stage.memory_cache
        self.announcer.n_blocks = len(stage)
        self.ready = threading.Event()
        self.runtime.start()
        self.runtime.ready.wait()
        self.dht.register_endpoint(peer_id, self.handler)
        self.announcer.announce(ServerState.ONLINE)
        logger.info(f"Started serving blocks {stage.start_block}:{stage.end_block} on {stage.device} "
                    f"({'sm_100a engine' if stage.engine is not None else 'oracle executor'})")
        self.ready.set()

    def is_healthy(self) -> bool:
        return self.runtime.is_alive() and self.announcer.is_alive()

    def shutdown(self) -> None:
        self.announcer.announce(ServerState.OFFLINE)
        self.announcer.stop.set()
        self.dht.unregister_endpoint(self.peer_id)
        self.handler.shutdown()
        self.runtime.shutdown()
        for backend in self.module_backends.values():
            backend.shutdown()
        if self.tp_group is not None:
            self.tp_group.shutdown()
            self.tp_group = None
        logger.info(f"Stage {self.peer_id} shut down")


class ModuleAnnouncerThread(threading.Thread):
    """Periodically publishes this stage's ``ServerInfo`` under every served block uid (reference :674-767)."""

    def __init__(self, module_uids: List[str], dht: Swarm, server_info: ServerInfo, model_info: ModelInfo, *, peer_id: str,
                 block_config, update_period: float, expiration: float, max_pinged: int = 5, **kwargs):
        super().__init__(**kwargs)
        self.module_uids, self.dht, self.server_info, self.model_info = module_uids, dht, server_info, model_info
        self.peer_id, self.block_config = peer_id, block_config
        self.update_period, self.expiration = update_period, expiration
        self.memory_cache: Optional[MemoryCache] = None
        self.n_blocks = len(module_uids)
        self.trigger, self.stop = threading.Event(), threading.Event()
        self.max_pinged = max_pinged
        self.ping_aggregator = PingAggregator(dht)
        prefix = module_uids[0].rsplit(".", 1)[0]
        last = int(module_uids[-1].rsplit(".", 1)[1])
        self.next_uids = [make_uid(prefix, last + 1)] if last + 1 < block_config.num_hidden_layers else []

    def run(self) -> None:
        while True:
            start = time.perf_counter()
            self.server_info.cache_tokens_left = (self.memory_cache.tokens_left * self.n_blocks) if self.memory_cache is not None else None
            try:
                if self.server_info.state != ServerState.OFFLINE:
                    self._ping_next_servers()
                    self.server_info.next_pings = {p: r for p, r in self.ping_aggregator.to_dict().items()}
                else:
                    self.server_info.next_pings = None
                self._publish()
                if self.server_info.state != ServerState.OFFLINE and hasattr(self.dht, "refresh_endpoint"):
                    self.dht.refresh_endpoint(self.peer_id)  # a restarted network registry learns our address again
            except ConnectionError as e:  # the registry of a multi-box swarm is away: keep serving, announce again later
                logger.warning(f"could not announce to the swarm registry: {e}")
            if self.server_info.state == ServerState.OFFLINE:
                break
            delay = self.update_period - (time.perf_counter() - start)
            if delay < 0:
                logger.warning("announcing took longer than update_period; consider increasing it")
            self.trigger.wait(max(delay, 0))
            self.trigger.clear()
            if self.stop.is_set() and self.server_info.state != ServerState.OFFLINE:
                self.server_info.state = ServerState.OFFLINE

    def _publish(self) -> None:
        declare_active_modules(self.dht, self.module_uids, self.server_info, expiration_time=get_dht_time() + self.expiration, peer_id=self.peer_id)
        if self.server_info.state == ServerState.ONLINE:
            self.dht.store("_petals.models", self.module_uids[0].rsplit(".", 1)[0], self.model_info.to_dict(), get_dht_time() + self.expiration)

    def announce(self, state: ServerState) -> None:
        self.server_info.state = state
        self._publish()  # synchronous so that callers (and tests) observe the new state immediately
        self.trigger.set()

    def _ping_next_servers(self) -> None:
        if not self.next_uids:
            return
        infos = get_remote_module_infos(self.dht, self.next_uids, latest=True)
        peers = [p for info in infos for p, s in info.servers.items() if s.state == ServerState.ONLINE and p != self.peer_id]
        if peers:
            self.ping_aggregator.ping(random.sample(peers, min(self.max_pinged, len(peers))))
