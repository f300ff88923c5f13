"""Falcon client shells (reference: src/petals/models/falcon/model.py:27-150). Unlike the reference (Q8) the same
``RemotePastKeyValues`` object is threaded through the whole generation so its seen-token count stays correct."""
from petals_b200.models.client_base import (DistributedModelBase, DistributedModelForCausalLM,
                                            DistributedModelForSequenceClassification)
from petals_b200.models.falcon.config import DistributedFalconConfig


class DistributedFalconModel(DistributedModelBase):
    config_class = DistributedFalconConfig

    @property
    def word_embeddings(self):
        return self.embed_tokens

    @property
    def h(self):
        return self.layers

    @property
    def ln_f(self):
        return self.final_norm


class DistributedFalconForCausalLM(DistributedModelForCausalLM):
    config_class = DistributedFalconConfig
    base_model_class = DistributedFalconModel

    @property
    def transformer(self):
        return self.model


class DistributedFalconForSequenceClassification(DistributedModelForSequenceClassification):
    config_class = DistributedFalconConfig
    base_model_class = DistributedFalconModel

    @property
    def transformer(self):
        return self.model
