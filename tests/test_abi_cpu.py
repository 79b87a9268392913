"""CPU-side checks of the boundary: the shared library loads without a GPU, exports
every symbol include/decaf377_b200.h declares, and refuses to compute without CUDA."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "decaf377_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(d377_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from decaf377_b200 import _lib
    lib = _lib.load()
    syms = _header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), "missing export %s" % s
    # and the Python binding table covers the header
    assert set(syms) == set(_lib.EXPORTS)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under decaf377_b200/ may import, link or
    load it (comments may mention it)."""
    pkg = os.path.join(ROOT, "decaf377_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".inc")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
                assert "d377_oracle" not in src and "c_oracle" not in src and "decaf377_ref" not in src, f


def test_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import decaf377_b200 as d
    from decaf377_b200._lib import D377Error
    with pytest.raises(D377Error):
        d.init(0)
    with pytest.raises(D377Error):
        d.batch_compress(np.zeros((1, 128), np.uint8))


def test_host_types_mirror_reference_semantics():
    from decaf377_b200 import EncodingError, Fq, Fr
    # fields/fq.rs:149-153, fr.rs:129-133
    assert Fq.from_bytes_checked(bytes(32)) == Fq(0)
    with pytest.raises(EncodingError):
        Fq.from_bytes_checked(b"\xff" * 32)
    with pytest.raises(EncodingError):
        Fr.from_bytes_checked(b"\xff" * 32)
    with pytest.raises(EncodingError) as ei:
        Fq.from_bytes_checked(bytes(31))
    assert ei.value.kind == "InvalidSliceLength"
    # fq/arkworks.rs:603-673
    assert Fq.from_le_bytes_mod_order((Fq.MODULUS + 1).to_bytes(32, "little")) == Fq(1)
    assert (Fq(-1)).square() == Fq(1)
    assert Fq(3) + Fq(4) == Fq(7) and Fq(3) * Fq(4) == Fq(12)
    assert Fq(5).is_negative() and not Fq(4).is_negative() and Fq(5).abs() == -Fq(5)
    x = Fq(123456789)
    assert Fq.from_montgomery_bytes(x.to_montgomery_bytes()) == x
    assert Fr(7).inverse() * Fr(7) == Fr(1)
