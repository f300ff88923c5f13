"""Model families (reference: src/petals/models/). Importing this package registers every family with
:mod:`petals_b200.utils.auto_config`."""
from petals_b200.models import bloom, falcon, llama, mixtral  # noqa: F401
