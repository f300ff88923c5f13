"""Mixtral client shells (reference: src/petals/models/mixtral/model.py:21-178)."""
from petals_b200.models.client_base import (DistributedModelBase, DistributedModelForCausalLM,
                                            DistributedModelForSequenceClassification)
from petals_b200.models.mixtral.config import DistributedMixtralConfig


class DistributedMixtralModel(DistributedModelBase):
    config_class = DistributedMixtralConfig

    @property
    def word_embeddings(self):
        return self.embed_tokens

    @property
    def h(self):
        return self.layers

    @property
    def ln_f(self):
        return self.final_norm

    @property
    def norm(self):  # Hugging Face's name of the final RMSNorm
        return self.final_norm


class DistributedMixtralForCausalLM(DistributedModelForCausalLM):
    config_class = DistributedMixtralConfig
    base_model_class = DistributedMixtralModel

    @property
    def transformer(self):
        return self.model


class DistributedMixtralForSequenceClassification(DistributedModelForSequenceClassification):
    config_class = DistributedMixtralConfig
    base_model_class = DistributedMixtralModel

    @property
    def transformer(self):
        return self.model
