"""Import shim: `import b381` -> the package in ./plonky2-bls12-381-pairing_b200/ (hyphenated name).
Every submodule is aliased too, so `b381.fields.types.Fq12` and the package's own class are the
same object."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_REAL = "plonky2-bls12-381-pairing_b200"
_pkg = importlib.import_module(_REAL)
for _name, _mod in list(sys.modules.items()):
    if _name == _REAL or _name.startswith(_REAL + "."):
        sys.modules["b381" + _name[len(_REAL):]] = _mod
