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


def _header_prototypes():
    """name -> number of parameters, from include/decaf377_b200.h."""
    text = open(os.path.join(ROOT, "include", "decaf377_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(d377_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        args = m.group(2).strip()
        protos[m.group(1)] = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
    return protos


def test_rust_shim_declares_the_header_faithfully():
    """rust/src/gpu.rs cannot be compiled in this image (no cargo / rustc); at least its
    `extern "C"` block must name real entry points with the right number of arguments."""
    src = open(os.path.join(ROOT, "rust", "src", "gpu.rs")).read()
    block = src[src.index('extern "C" {'):]
    block = block[:block.index("\n}\n")]
    protos = _header_prototypes()
    decls = re.findall(r"fn\s+(d377_[a-z0-9_]+)\s*\((.*?)\)\s*(?:->\s*[^;]+)?;", block, flags=re.S)
    assert len(decls) >= 20
    for name, args in decls:
        assert name in protos, "gpu.rs declares %s, which the header does not have" % name
        n = len([a for a in args.split(",") if a.strip()])
        assert n == protos[name], "%s: %d arguments in gpu.rs, %d in the header" % (name, n, protos[name])
    # and every call in the file goes to a declared function
    used = set(re.findall(r"\b(d377_[a-z0-9_]+)\s*\(", src)) | set(re.findall(r"\b(d377_batch_[a-z]+)\b", src))
    assert used <= set(protos), sorted(used - set(protos))


def test_c_host_program_compiles_and_links_against_the_library(tmp_path):
    """tests/abi_smoke.c against the header and the built library with -Wall -Werror (it is RUN
    by the GPU suite; here only the boundary's C view is checked)."""
    import subprocess
    libdir = os.path.join(ROOT, "decaf377_b200")
    if not os.path.exists(os.path.join(libdir, "libdecaf377_b200.so")):
        pytest.skip("library not built")
    exe = tmp_path / "abi_smoke"
    r = subprocess.run(["gcc", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "tests", "abi_smoke.c"), "-o", str(exe),
                        "-L", libdir, "-ldecaf377_b200", "-Wl,-rpath," + libdir],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    import torch
    if not torch.cuda.is_available():
        # without a GPU the program must fail loudly, not pretend
        r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
        assert r.returncode != 0 and "check(s) failed" in r.stdout
