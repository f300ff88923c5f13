"""BLOOM client shells (reference: src/petals/models/bloom/model.py:21-197): embedding LayerNorm, LayerNorm ``ln_f``,
LM head tied to the embeddings."""
from petals_b200.models.bloom.config import DistributedBloomConfig
from petals_b200.models.client_base import (DistributedModelBase, DistributedModelForCausalLM,
                                            DistributedModelForSequenceClassification)


class DistributedBloomModel(DistributedModelBase):
    config_class = DistributedBloomConfig
    has_embedding_layernorm = True

    @property
    def word_embeddings(self):
        return self.embed_tokens

    @property
    def word_embeddings_layernorm(self):
        return self.embed_layernorm

    @property
    def h(self):
        return self.layers

    @property
    def ln_f(self):
        return self.final_norm


class DistributedBloomForCausalLM(DistributedModelForCausalLM):
    config_class = DistributedBloomConfig
    base_model_class = DistributedBloomModel

    @property
    def transformer(self):
        return self.model


class DistributedBloomForSequenceClassification(DistributedModelForSequenceClassification):
    config_class = DistributedBloomConfig
    base_model_class = DistributedBloomModel

    @property
    def transformer(self):
        return self.model
