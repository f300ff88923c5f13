"""Model families (reference: src/petals/models/). Importing this package registers every family with
:mod:`petals_b200.utils.auto_config` and re-exports their classes (``from petals_b200.models import DistributedLlamaForCausalLM``)."""
from petals_b200.models import bloom, falcon, llama, mixtral
from petals_b200.models.bloom import *  # noqa: F401,F403
from petals_b200.models.falcon import *  # noqa: F401,F403
from petals_b200.models.llama import *  # noqa: F401,F403
from petals_b200.models.mixtral import *  # noqa: F401,F403

__all__ = sorted(set(bloom.__all__ + falcon.__all__ + llama.__all__ + mixtral.__all__))
