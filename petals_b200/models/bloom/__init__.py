from petals_b200.models.bloom.block import WrappedBloomBlock
from petals_b200.models.bloom.config import DistributedBloomConfig
from petals_b200.models.bloom.model import (DistributedBloomForCausalLM, DistributedBloomForSequenceClassification,
                                            DistributedBloomModel)
from petals_b200.utils.auto_config import register_model_classes

register_model_classes(config=DistributedBloomConfig, model=DistributedBloomModel, model_for_causal_lm=DistributedBloomForCausalLM,
                       model_for_sequence_classification=DistributedBloomForSequenceClassification, block=WrappedBloomBlock)
