"""Import-path parity with the reference's ``petals.client.from_pretrained`` (src/petals/client/from_pretrained.py:17-84).

The reference patches Hugging Face's shard resolution so that a client never downloads the shards that only hold
transformer blocks.  Here the client shells are not HF modules: :class:`FromPretrainedMixin` reads the safetensors
index itself and opens only the files that hold embeddings / final norm / head (``load_client_tensors``), and it also
owns the save side of the checkpoint/resume contract (``save_pretrained`` of the client-only state).  The implementation
lives next to the model shells in :mod:`petals_b200.models.client_base`; this module is the reference's name for it.
"""
from petals_b200.models.client_base import (TRAINABLE_STATE_NAMES, FromPretrainedMixin, client_state_names,  # noqa: F401
                                            load_client_tensors)

__all__ = ["FromPretrainedMixin", "load_client_tensors", "client_state_names", "TRAINABLE_STATE_NAMES"]
