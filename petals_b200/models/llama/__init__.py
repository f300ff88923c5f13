"""Llama family: config + client shells + block wrapper, registered with the ``AutoDistributed*`` factories on import."""
from petals_b200.utils.auto_config import register_family

_classes = register_family(__name__, "Llama", speculative=True)
globals().update(_classes)
__all__ = sorted(_classes)
