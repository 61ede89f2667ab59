"""The C-ABI library loads and exports every symbol include/b381.h declares; without a GPU the
compute entry points fail loudly (no CPU fallback).  No compute calls here."""
import ctypes
import os
import re

import pytest

import util


def header_symbols():
    text = open(os.path.join(util.ROOT, "include", "b381.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b381_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import b381
    lib = b381._lib.load()
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), "libb381.so does not export %s" % s
    assert sorted(b381._lib.SIGNATURES) == syms, "ctypes signature table and header disagree"


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import b381
    lib = b381._lib.load()
    assert lib.b381_init(0) == -1                         # B381_E_CUDA
    assert b"no CPU fallback" in lib.b381_last_error()
    out = (ctypes.c_uint32 * 12)()
    assert lib.b381_fp_mul(out, out, out, 1) == -5        # B381_E_NOT_INIT
    with pytest.raises(RuntimeError):
        b381._lib.init(0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(util.ROOT, "plonky2-bls12-381-pairing_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".inc", ".rs")):
                text = open(os.path.join(dp, f)).read()
                for needle in ("b381_oracle", "b381_ref", "libb381_hostsim", "oracle/_build", "oracle/_ref"):
                    assert needle not in text, (f, needle)


def test_roofline_probe_executes_multiplies():
    """The roofline denominator must come from real multiplier instructions: the loop body of
    k_imad_peak has to contain its 128 fused IMAD.WIDE.U32 (the first version of the probe was
    strength-reduced by ptxas to IADD3 chains and overstated the multiplier peak twofold)."""
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    so = os.path.join(util.ROOT, "plonky2-bls12-381-pairing_b200", "libb381.so")
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    body, on = [], False
    for line in sass.split("\n"):
        if "Function :" in line:
            on = "k_imad_peak" in line
        elif on:
            body.append(line)
    wide = [l for l in body if "IMAD.WIDE.U32" in l]
    fused = [l for l in wide if not re.search(r", RZ ;", l)]
    adds = [l for l in body if re.search(r"\bIADD3", l)]
    assert len(fused) >= 128, "probe multiplies were optimised away (%d fused IMAD.WIDE)" % len(fused)
    assert len(adds) < 40, "probe was turned into adds (%d IADD3)" % len(adds)
