"""B200-native batched BLS12-381 pairing engine (host-side Python binding).

The directory name contains hyphens, so import it with
    importlib.import_module("plonky2-bls12-381-pairing_b200")
or through the repo-root shim `import b381`.
"""
from . import _lib  # noqa: F401
