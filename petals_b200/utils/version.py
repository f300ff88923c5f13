"""Version helpers (reference: src/petals/utils/version.py). There is no index to query for updates on an
offline box, so ``validate_version`` only reports the running version; the ``-petals`` repo-name compatibility
shim of the reference is kept as a no-op resolver."""
import os
import re
from typing import Union

import petals_b200
from petals_b200.utils.logging import get_logger

logger = get_logger(__name__)


def validate_version() -> None:
    logger.info(f"Running petals_b200 {petals_b200.__version__}")


def get_compatible_model_repo(model_name_or_path: Union[str, os.PathLike, None]) -> Union[str, os.PathLike, None]:
    if model_name_or_path is None:
        return None
    # early Petals mirrors were named "<org>/<model>-petals"; plain names are the only ones used here
    return re.sub(r"-petals$", "", str(model_name_or_path)) if not os.path.isdir(str(model_name_or_path)) else model_name_or_path
