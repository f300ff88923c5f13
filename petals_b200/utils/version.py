"""Version helpers (reference: src/petals/utils/version.py).

The reference asks PyPI whether a newer release exists and warns. An offline box has no index, but it has something more relevant: the
versions the OTHER peers of the swarm announce in their ``ServerInfo`` records. ``validate_version`` therefore reports the running
version and, given a swarm, warns when peers serving the same model run a newer one (a mixed-version swarm is the situation the
reference's check is there to prevent). The ``-petals`` repo-name compatibility shim of the reference is kept as a resolver."""
import os
import re
from typing import Iterable, Optional, Sequence, Tuple, Union

import petals_b200
from petals_b200.utils.logging import get_logger

logger = get_logger(__name__)


def parse_version(text: Optional[str]) -> Tuple[int, ...]:
    """"2.3.0.dev2" -> (2, 3, 0) (numeric release segment only; anything unparsable sorts lowest)."""
    m = re.match(r"\s*v?(\d+(?:\.\d+)*)", str(text or ""))
    return tuple(int(p) for p in m.group(1).split(".")) if m else ()


def newest_version(versions: Iterable[Optional[str]]) -> Optional[str]:
    best = None
    for v in versions:
        if v and (best is None or parse_version(v) > parse_version(best)):
            best = v
    return best


def swarm_versions(dht, uids: Sequence[str]) -> Sequence[str]:
    """Versions announced by the peers currently serving ``uids`` (empty when the registry is unreachable)."""
    from petals_b200.utils.dht import get_remote_module_infos

    try:
        infos = get_remote_module_infos(dht, list(uids), latest=True)
    except Exception as e:  # noqa: BLE001 - a version hint must never stop a server or a client from starting
        logger.debug(f"could not read peer versions from the swarm: {e!r}")
        return []
    return [srv.version for info in infos if info is not None for srv in info.servers.values() if getattr(srv, "version", None)]


def validate_version(dht=None, uids: Optional[Sequence[str]] = None) -> Optional[str]:
    """Log the running version; with a swarm, warn if peers serving the same blocks run a newer one. Returns that newer version (or None)."""
    mine = petals_b200.__version__
    logger.info(f"Running petals_b200 {mine}")
    if dht is None or not uids:
        return None
    newest = newest_version(swarm_versions(dht, uids))
    if newest is not None and parse_version(newest) > parse_version(mine):
        logger.warning(f"Peers of this swarm run petals_b200 {newest}, this process runs {mine}: update to avoid protocol mismatches")
        return newest
    return None


def get_compatible_model_repo(model_name_or_path: Union[str, os.PathLike, None]) -> Union[str, os.PathLike, None]:
    if model_name_or_path is None:
        return None
    # early Petals mirrors were named "<org>/<model>-petals"; plain names are the only ones used here
    return re.sub(r"-petals$", "", str(model_name_or_path)) if not os.path.isdir(str(model_name_or_path)) else model_name_or_path
