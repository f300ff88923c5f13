"""Utilities (reference: src/petals/utils/). Like the reference's package, it offers the ``AutoDistributed*`` factories and the two
membership helpers at package level — resolved lazily (PEP 562), because low-level modules (``utils.logging``) are imported by
the very packages those helpers depend on."""
import importlib

_EXPORTS = {
    "AutoDistributedConfig": "auto_config", "AutoDistributedModel": "auto_config", "AutoDistributedModelForCausalLM": "auto_config",
    "AutoDistributedModelForSequenceClassification": "auto_config", "AutoDistributedSpeculativeModel": "auto_config",
    "declare_active_modules": "dht", "get_remote_module_infos": "dht",
}
__all__ = sorted(_EXPORTS)


def __getattr__(name: str):
    module = _EXPORTS.get(name)
    if module is None:
        raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
    value = getattr(importlib.import_module(f"{__name__}.{module}"), name)
    globals()[name] = value
    return value
