"""``petals`` compatibility namespace: the reference's import paths on top of :mod:`petals_b200`.

Scripts written against bigscience-workshop/petals (``from petals import AutoDistributedModelForCausalLM``,
``python -m petals.cli.run_server``, ``petals.server.from_pretrained.load_pretrained_block`` ...) run unmodified.
Only the *names* are shared with the reference; every module re-exports the from-scratch implementation."""
import importlib
import sys

import petals_b200
from petals_b200 import __version__  # noqa: F401
from petals_b200.client import *  # noqa: F401,F403
from petals_b200.client import ClientConfig, InferenceSession, RemoteSequenceManager, RemoteSequential  # noqa: F401
from petals_b200.models import *  # noqa: F401,F403
from petals_b200.utils import *  # noqa: F401,F403
from petals_b200.utils.auto_config import (AutoDistributedConfig, AutoDistributedModel, AutoDistributedModelForCausalLM,  # noqa: F401
                                           AutoDistributedModelForSequenceClassification, AutoDistributedSpeculativeModel)

_ALIASES = [
    "constants", "data_structures", "dht_utils",
    "client", "client.config", "client.inference_session", "client.remote_sequential", "client.sequential_autograd",
    "client.remote_forward_backward", "client.remote_generation", "client.lm_head", "client.ptune", "client.from_pretrained", "client.routing",
    "client.routing.sequence_manager", "client.routing.sequence_info", "client.routing.spending_policy",
    "server", "server.server", "server.backend", "server.handler", "server.block_functions", "server.task_pool",
    "server.task_prioritizer", "server.memory_cache", "server.block_selection", "server.throughput", "server.block_utils",
    "server.from_pretrained", "server.reachability",
    "utils", "utils.auto_config", "utils.convert_block", "utils.cuda_graphs", "utils.peft", "utils.packaging", "utils.misc",
    "utils.disk_cache", "utils.compression", "utils.dht", "utils.ping", "utils.logging", "utils.version", "utils.hf_auth", "utils.random", "utils.asyncio",
    "models", "models.llama", "models.bloom", "models.falcon", "models.mixtral",
]
for _name in _ALIASES:
    try:
        _module = importlib.import_module(f"petals_b200.{_name}")
        sys.modules[f"petals.{_name}"] = _module
        if "." not in _name:
            globals()[_name] = _module  # `petals.models`, `petals.client`, ... as attributes, like real sub-packages
    except ModuleNotFoundError as _e:  # a module that is not implemented yet must not break `import petals`
        if not str(_e).startswith("No module named 'petals_b200"):
            raise
