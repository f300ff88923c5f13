"""Client side: RemoteSequential, inference sessions, routing (reference: src/petals/client/)."""
from petals_b200.client.config import ClientConfig  # noqa: F401
from petals_b200.client.inference_session import InferenceSession  # noqa: F401
from petals_b200.client.remote_sequential import RemoteSequential  # noqa: F401
from petals_b200.client.routing import NoSpendingPolicy, RemoteSequenceManager, SpendingPolicyBase  # noqa: F401
