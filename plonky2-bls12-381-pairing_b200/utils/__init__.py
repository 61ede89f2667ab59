"""Mirror of /root/reference/src/utils/mod.rs (constants + helpers)."""
from . import constants  # noqa: F401
from ..fields import helpers  # noqa: F401  (utils/helpers.rs is a byte-identical copy of fields/helpers.rs)
