from petals_b200.models.mixtral.block import WrappedMixtralBlock
from petals_b200.models.mixtral.config import DistributedMixtralConfig
from petals_b200.models.mixtral.model import (DistributedMixtralForCausalLM, DistributedMixtralForSequenceClassification,
                                              DistributedMixtralModel)
from petals_b200.utils.auto_config import register_model_classes

register_model_classes(config=DistributedMixtralConfig, model=DistributedMixtralModel, model_for_causal_lm=DistributedMixtralForCausalLM,
                       model_for_sequence_classification=DistributedMixtralForSequenceClassification, block=WrappedMixtralBlock)
