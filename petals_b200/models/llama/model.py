"""Llama client shells (reference: src/petals/models/llama/model.py:20-174)."""
from petals_b200.models.client_base import (DistributedModelBase, DistributedModelForCausalLM,
                                            DistributedModelForSequenceClassification)
from petals_b200.models.llama.config import DistributedLlamaConfig


class DistributedLlamaModel(DistributedModelBase):
    config_class = DistributedLlamaConfig

    # compatibility aliases used by BLOOM-era scripts (reference :115-129)
    @property
    def word_embeddings(self):
        return self.embed_tokens

    @property
    def word_embeddings_layernorm(self):
        import torch.nn as nn

        return nn.Identity()

    @property
    def h(self):
        return self.layers

    @property
    def ln_f(self):
        return self.final_norm

    @property
    def norm(self):  # Hugging Face's name of the final RMSNorm
        return self.final_norm


class DistributedLlamaForCausalLM(DistributedModelForCausalLM):
    config_class = DistributedLlamaConfig
    base_model_class = DistributedLlamaModel

    @property
    def transformer(self):  # for compatibility with RemoteGenerationMixin users written against BLOOM
        return self.model


class DistributedLlamaForSequenceClassification(DistributedModelForSequenceClassification):
    config_class = DistributedLlamaConfig
    base_model_class = DistributedLlamaModel

    @property
    def transformer(self):
        return self.model
