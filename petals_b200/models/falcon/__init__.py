from petals_b200.models.falcon.block import WrappedFalconBlock
from petals_b200.models.falcon.config import DistributedFalconConfig
from petals_b200.models.falcon.model import (DistributedFalconForCausalLM, DistributedFalconForSequenceClassification,
                                             DistributedFalconModel)
from petals_b200.utils.auto_config import register_model_classes

register_model_classes(config=DistributedFalconConfig, model=DistributedFalconModel, model_for_causal_lm=DistributedFalconForCausalLM,
                       model_for_sequence_classification=DistributedFalconForSequenceClassification, block=WrappedFalconBlock)
