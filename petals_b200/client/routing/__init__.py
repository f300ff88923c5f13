"""Routing: which stage serves which blocks, and which chain a request should take."""
from petals_b200.client.routing.sequence_manager import MissingBlocksError, RemoteSequenceManager, maybe_log_traceback  # noqa: F401
from petals_b200.client.routing.spending_policy import NoSpendingPolicy, SpendingPolicyBase  # noqa: F401
