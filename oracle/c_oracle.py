"""ctypes front end of oracle/d377_oracle.c (TEST INFRASTRUCTURE ONLY -- see the
header of that file).  Used by tests/ and by bench.py's cpu_baseline /
`--impl reference` legs; never by the product."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "_build" / "libd377_oracle.so"
_lib = None


def build(force: bool = False) -> Path:
    src = HERE / "d377_oracle.c"
    if force or not LIB.exists() or LIB.stat().st_mtime < src.stat().st_mtime:
        # -march=native objects must not travel between hosts: rebuild when missing/stale only
        subprocess.run(["make", "-C", str(HERE), "-s", "-B"], check=True)
    return LIB


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        try:
            _lib = C.CDLL(str(LIB))
        except OSError:
            build(force=True)
            _lib = C.CDLL(str(LIB))
        _lib.d377o_init()
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _u8(a, w):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a.reshape(-1, w)


def threads_default() -> int:
    return os.cpu_count() or 1


def decompress(enc, threads=1):
    enc = _u8(enc, 32); n = enc.shape[0]
    out = np.empty((n, 128), np.uint8); ok = np.empty((n,), np.uint8)
    lib().d377o_decompress(_p(enc), C.c_size_t(n), _p(out), _p(ok), threads)
    return out, ok


def compress(el, threads=1):
    el = _u8(el, 128); n = el.shape[0]
    out = np.empty((n, 32), np.uint8)
    lib().d377o_compress(_p(el), C.c_size_t(n), _p(out), threads)
    return out


def encode_to_curve(r, out_enc=False, threads=1):
    r = _u8(r, 32); n = r.shape[0]
    out = np.empty((n, 32 if out_enc else 128), np.uint8)
    lib().d377o_encode_to_curve(_p(r), C.c_size_t(n), _p(out), int(out_enc), threads)
    return out


def hash_to_curve(r1, r2, out_enc=False, threads=1):
    r1 = _u8(r1, 32); r2 = _u8(r2, 32); n = r1.shape[0]
    out = np.empty((n, 32 if out_enc else 128), np.uint8)
    lib().d377o_hash_to_curve(_p(r1), _p(r2), C.c_size_t(n), _p(out), int(out_enc), threads)
    return out


def scalar_mul(pts, sc, out_enc=False, threads=1):
    pts = _u8(pts, 128); sc = _u8(sc, 32); n = pts.shape[0]
    out = np.empty((n, 32 if out_enc else 128), np.uint8)
    lib().d377o_scalar_mul(_p(pts), _p(sc), C.c_size_t(n), _p(out), int(out_enc), threads)
    return out


def pipeline(enc, sc, threads=1):
    """config 1: vartime_decompress -> * Fr -> vartime_compress"""
    enc = _u8(enc, 32); sc = _u8(sc, 32); n = enc.shape[0]
    out = np.empty((n, 32), np.uint8); ok = np.empty((n,), np.uint8)
    lib().d377o_pipeline(_p(enc), _p(sc), C.c_size_t(n), _p(out), _p(ok), threads)
    return out, ok


def fixed_base(sc, out_enc=True, threads=1):
    sc = _u8(sc, 32); n = sc.shape[0]
    out = np.empty((n, 32 if out_enc else 128), np.uint8)
    lib().d377o_fixed_base(_p(sc), C.c_size_t(n), _p(out), int(out_enc), threads)
    return out


def add(a, b, threads=1):
    a = _u8(a, 128); b = _u8(b, 128)
    out = np.empty_like(a)
    lib().d377o_add(_p(a), _p(b), C.c_size_t(a.shape[0]), _p(out), threads)
    return out


def fq_mul(a, b, threads=1):
    a = _u8(a, 32); b = _u8(b, 32)
    out = np.empty_like(a)
    lib().d377o_fq_mul(_p(a), _p(b), C.c_size_t(a.shape[0]), _p(out), threads)
    return out


def sqrt_ratio_zeta(num, den):
    num = _u8(num, 32); den = _u8(den, 32); n = num.shape[0]
    out = np.empty((n, 32), np.uint8); ws = np.empty((n,), np.uint8)
    lib().d377o_sqrt_ratio_zeta(_p(num), _p(den), C.c_size_t(n), _p(out), _p(ws))
    return out, ws


def msm_fold(sc, pts):
    sc = _u8(sc, 32); pts = _u8(pts, 128); n = min(sc.shape[0], pts.shape[0])
    el = np.empty((128,), np.uint8); enc = np.empty((32,), np.uint8)
    lib().d377o_msm_fold(_p(sc), _p(pts), C.c_size_t(n), _p(el), _p(enc))
    return el, enc


def msm_pippenger(sc, pts, threads=1):
    sc = _u8(sc, 32); pts = _u8(pts, 128); n = min(sc.shape[0], pts.shape[0])
    el = np.empty((128,), np.uint8); enc = np.empty((32,), np.uint8)
    lib().d377o_msm_pippenger(_p(sc), _p(pts), C.c_size_t(n), _p(el), _p(enc), threads)
    return el, enc
