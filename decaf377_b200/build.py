"""In-tree build of libdecaf377_b200.so (hand-written CUDA for sm_100a).

``python -m decaf377_b200.build`` or ``build_lib()``; the shared object lands
next to this file so that it travels with the source snapshot.  There is no
JIT cache and no CPU fallback: if the library is missing, importing the engine
fails loudly.

``build_lib(debug=True)`` (``--debug``) builds libdecaf377_b200_dbg.so with
-DD377_DEBUG_ON_CURVE -DD377_DEBUG_ORDER: every point a kernel produces is checked
against the reference's OnCurve predicate (ark_curve/on_curve.rs:17-38).  The engine
loads it instead of the release library when D377_DEBUG_LIB=1.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB = HERE / "libdecaf377_b200.so"
LIB_DEBUG = HERE / "libdecaf377_b200_dbg.so"
SOURCES = ["kernels.cu", "codec.cu", "scalar.cu", "msm.cu", "multi.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
]
DEBUG_FLAGS = ["-DD377_DEBUG_ON_CURVE", "-DD377_DEBUG_ORDER"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _headers():
    return list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + list(CSRC.glob("*.inc")) \
        + [HERE.parent / "include" / "decaf377_b200.h"]


def _newer(deps, target: Path) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps)


def build_lib(force: bool = False, verbose: bool = False, debug: bool = False) -> Path:
    lib = LIB_DEBUG if debug else LIB
    nvcc = _nvcc()
    objdir = HERE / "build" / ("dbg" if debug else "rel")
    objdir.mkdir(parents=True, exist_ok=True)
    hdrs = _headers()
    extra = DEBUG_FLAGS if debug else []

    def compile_one(src: str) -> Path:
        obj = objdir / (src + ".o")
        if not force and not _newer([CSRC / src] + hdrs, obj):
            return obj
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", str(CSRC / src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    if force or _newer(objs, lib):
        cmd = [nvcc, "-shared", "-o", str(lib), *map(str, objs), "-gencode",
               "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return lib


def build_all(force: bool = False) -> None:
    """Release and debug libraries, compiled side by side."""
    with ThreadPoolExecutor(max_workers=2) as ex:
        futs = [ex.submit(build_lib, force, False, dbg) for dbg in (False, True)]
        for f in futs:
            f.result()


if __name__ == "__main__":
    if "--all" in sys.argv:
        build_all(force="--force" in sys.argv)
        print(LIB, LIB_DEBUG)
    else:
        p = build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv, debug="--debug" in sys.argv)
        print(p)
