"""Auth-token plumbing kept for call-site compatibility (reference: src/petals/utils/hf_auth.py). Offline: ignored."""
from typing import Optional, Union


def always_needs_auth(model_name: Union[str, None]) -> bool:
    return False


def resolve_token(token: Optional[Union[str, bool]] = None) -> Optional[str]:
    return token if isinstance(token, str) else None
