from petals_b200.models.llama.block import WrappedLlamaBlock
from petals_b200.models.llama.config import DistributedLlamaConfig
from petals_b200.models.llama.model import (DistributedLlamaForCausalLM, DistributedLlamaForSequenceClassification,
                                            DistributedLlamaModel)
from petals_b200.models.llama.speculative_model import DistributedLlamaForSpeculativeGeneration
from petals_b200.utils.auto_config import register_model_classes

register_model_classes(config=DistributedLlamaConfig, model=DistributedLlamaModel, model_for_causal_lm=DistributedLlamaForCausalLM,
                       model_for_speculative=DistributedLlamaForSpeculativeGeneration,
                       model_for_sequence_classification=DistributedLlamaForSequenceClassification, block=WrappedLlamaBlock)
