"""GPU parity, second batch: the entry points added for the callers either side of the hot
path -- wide Fq inputs (fields/fq.rs:90-102), Sub / Neg / double
(ark_curve/ops/projective.rs:50-104), OnCurve (ark_curve/on_curve.rs:17-38), the
asynchronous and the multi-GPU MSM (element/projective.rs:99-117 over SURVEY 8e), and the
robustness contracts of the C ABI (threads, untrusted limbs, prepared-bases registry).
All through the C ABI, all against the oracle, bit-exact."""
import ctypes as C
import os
import random
import subprocess
import sys
import threading

import numpy as np
import pytest

from oracle import c_oracle as co
from oracle import decaf377_ref as o
from tests.util import canon, mont, np_bytes, oracle_points, oracle_scalars, unmont, unwire, wire

pytestmark = pytest.mark.gpu

Q, R = o.Q, o.R
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- wide inputs (Missing #3 of round 1) ----------------------------------------------------
def test_from_le_bytes_mod_order_80_bytes_property(engine):
    """fq/arkworks.rs:586-593: any 80 bytes, chunked fold == naive reduction."""
    rnd = random.Random(11)
    rows = [bytes(80), b"\xff" * 80] + [rnd.randbytes(80) for _ in range(3000)]
    got = unmont(engine.fq_batch_from_le_bytes_mod_order(np_bytes(rows, 80)))
    assert got == [int.from_bytes(b, "little") % Q for b in rows]
    assert got == [o.fq_from_le_bytes_mod_order_chunked(b) for b in rows]


@pytest.mark.parametrize("width", [1, 31, 32, 33, 48, 64, 65, 96, 255, 256])
def test_from_le_bytes_mod_order_any_width(engine, width):
    rnd = random.Random(width)
    rows = [b"\xff" * width] + [rnd.randbytes(width) for _ in range(257)]
    got = unmont(engine.fq_batch_from_le_bytes_mod_order(np_bytes(rows, width)))
    assert got == [o.fq_from_le_bytes_mod_order_chunked(b) for b in rows]


def test_hash_and_encode_to_curve_64_byte_inputs(engine):
    """64-byte hash outputs are the usual hash_to_curve input (elligator.rs:67-76 after
    Fq::from_le_bytes_mod_order, fq.rs:90-102)."""
    n = 300
    b1 = [o.xof_bytes("wide1/%d" % i, 2) for i in range(n)]      # 64 bytes each
    b2 = [o.xof_bytes("wide2/%d" % i, 2) for i in range(n)]
    r1 = [o.fq_from_le_bytes_mod_order_chunked(b) for b in b1]
    r2 = [o.fq_from_le_bytes_mod_order_chunked(b) for b in b2]
    A, B = np_bytes(b1, 64), np_bytes(b2, 64)
    want_h = [o.compress(o.hash_to_curve(x, y)) for x, y in zip(r1, r2)]
    want_e = [o.compress(o.encode_to_curve(x)) for x in r1]
    got_h = engine.batch_hash_to_curve(A, B, engine.OUT_ENCODING)
    got_e = engine.batch_encode_to_curve(A, engine.OUT_ENCODING)
    assert [got_h[i].tobytes() for i in range(n)] == want_h
    assert [got_e[i].tobytes() for i in range(n)] == want_e
    # Element outputs of the same calls compress to the same bytes
    assert np.array_equal(engine.batch_compress(engine.batch_hash_to_curve(A, B)), got_h)
    assert np.array_equal(engine.batch_compress(engine.batch_encode_to_curve(A)), got_e)
    # a 64-byte input whose upper half is zero is the 32-byte input
    lo = np_bytes([b[:32] for b in b1], 32)
    padded = np.concatenate([lo, np.zeros_like(lo)], axis=1)
    assert np.array_equal(engine.batch_encode_to_curve(padded, engine.OUT_ENCODING),
                          engine.batch_encode_to_curve(lo, engine.OUT_ENCODING))


# ---- Sub / Neg / double ----------------------------------------------------------------------
def test_sub_neg_double_match_oracle(engine):
    n = 200
    P = oracle_points("r2/P", n) + [o.IDENTITY, o.GENERATOR]
    Qs = oracle_points("r2/Q", n) + [o.GENERATOR, o.GENERATOR]
    A, B = wire(P), wire(Qs)
    sub = engine.batch_compress(engine.batch_sub(A, B))
    neg = engine.batch_compress(engine.batch_neg(A))
    dbl = engine.batch_compress(engine.batch_double(A))
    for i, (p, q) in enumerate(zip(P, Qs)):
        assert sub[i].tobytes() == o.compress(o.point_add(p, o.point_neg(q))), i
        assert neg[i].tobytes() == o.compress(o.point_neg(p)), i
        assert dbl[i].tobytes() == o.compress(o.point_double(p)), i
    # Neg is exactly (X, Y, Z, T) -> (-X, Y, Z, -T) (ops/projective.rs:90-96)
    got = unwire(engine.batch_neg(A))
    for p, g in zip(P, got):
        assert g == ((-p[0]) % Q, p[1] % Q, p[2] % Q, (-p[3]) % Q)
    # the host mirror uses them
    a, b = engine.Element(A[0].tobytes()), engine.Element(B[0].tobytes())
    assert (a - b) + b == a and -(-a) == a and a.double() == a + a
    assert (a - a).is_identity()


# ---- AffinePoint wire format ---------------------------------------------------
def test_affine_and_xyz_point_layouts_of_compress_and_decompress(engine):
    """CanonicalSerialize / CanonicalDeserialize for AffinePoint (ark_curve/serialize.rs:8-46):
    serialize = into Element, then vartime_compress; deserialize = vartime_decompress, then
    into AffinePoint.  Same bytes as the Element path, no 128-byte detour."""
    from decaf377_b200 import device as dev
    from tests import golden_vectors as gv
    import torch
    n = 300
    P = oracle_points("r2/aff", n) + [o.IDENTITY, o.GENERATOR]
    assert any(p[2] != 1 for p in P)
    W = wire(P)
    want = [o.compress(p) for p in P]
    # AffinePoint inputs: x = X/Z, y = Y/Z
    aff = []
    for x, y, z, _ in P:
        zi = pow(z, Q - 2, Q)
        aff.append(o.fq_to_mont_bytes(x * zi % Q) + o.fq_to_mont_bytes(y * zi % Q))
    A = np_bytes(aff, 64)
    got_a = engine.batch_compress_fmt(A, engine.PT_AFFINE)
    got_x = engine.batch_compress_fmt(np.ascontiguousarray(W[:, :96]), engine.PT_XYZ)
    got_e = engine.batch_compress_fmt(W, engine.PT_ELEMENT)
    for i in range(len(P)):
        assert got_a[i].tobytes() == want[i], i
        assert got_x[i].tobytes() == want[i], i
        assert got_e[i].tobytes() == want[i], i
    assert np.array_equal(engine.batch_affine_serialize(A), got_a)
    # ... and the way back, invalid encodings included
    raw = o.xof_blocks("r2/affraw", 1500)
    raw = [b[:31] + bytes([b[31] & 0x1F]) for b in raw] + want + [bytes(x) for x in gv.EDGE_ENCODINGS]
    E = np_bytes(raw, 32)
    xy, ok = engine.batch_affine_deserialize(E)
    el, ok_el = engine.batch_decompress(E)
    assert np.array_equal(ok, ok_el) and 0 < int(ok.sum()) < len(raw)
    assert np.array_equal(xy, el[:, :64])
    ident = o.fq_to_mont_bytes(0) + o.fq_to_mont_bytes(1)
    for i, b in enumerate(raw):
        p = o.decompress(b)
        assert bool(ok[i]) == (p is not None), i
        if p is None:
            assert xy[i].tobytes() == ident, i
        else:
            assert p[2] == 1
            assert xy[i].tobytes() == o.fq_to_mont_bytes(p[0]) + o.fq_to_mont_bytes(p[1]), i
    # the 16 known answers of tests/encoding.rs:61-78 survive deserialize -> serialize
    kat = np_bytes([bytes.fromhex(h) if isinstance(h, str) else bytes(h) for h in gv.GENERATOR_MULTIPLES], 32)
    kxy, kok = engine.batch_affine_deserialize(kat)
    assert kok.all() and np.array_equal(engine.batch_affine_serialize(kxy), kat)
    # the host mirror of AffinePoint goes through the same calls
    ap = engine.AffinePoint.deserialize_compressed(want[0])
    assert ap.serialize_compressed() == want[0] and ap.wire == xy[1500].tobytes()
    with pytest.raises(engine.EncodingError):
        engine.AffinePoint.deserialize_compressed(bytes([1]) + bytes(31))
    # device-pointer twins
    dE = torch.from_numpy(E).cuda()
    dxy, dok = dev.decompress_fmt(dE, engine.PT_AFFINE)
    assert np.array_equal(dxy.cpu().numpy(), xy) and np.array_equal(dok.cpu().numpy(), ok)
    assert np.array_equal(dev.compress_fmt(torch.from_numpy(A).cuda(), engine.PT_AFFINE).cpu().numpy(), got_a)
    # layouts that make no sense for the call are refused
    with pytest.raises(ValueError):
        engine.batch_compress_fmt(E, engine.PT_ENCODING)
    lib = __import__("decaf377_b200")._lib.load()
    buf = np.zeros((1, 128), np.uint8)
    assert lib.d377_batch_compress_fmt(E.ctypes.data_as(C.POINTER(C.c_uint8)), engine.PT_ENCODING, 1,
                                       buf.ctypes.data_as(C.POINTER(C.c_uint8))) != 0
    assert lib.d377_batch_decompress_fmt(E.ctypes.data_as(C.POINTER(C.c_uint8)), 1, engine.PT_XYZ,
                                         buf.ctypes.data_as(C.POINTER(C.c_uint8)), None) != 0
    # empty batches
    assert engine.batch_affine_deserialize(np.zeros((0, 32), np.uint8))[0].shape == (0, 64)
    assert engine.batch_affine_serialize(np.zeros((0, 64), np.uint8)).shape == (0, 32)


# ---- OnCurve ---------------------------------------------------------------------------------
def test_on_curve_predicate(engine):
    P = oracle_points("r2/oc", 64) + [o.IDENTITY, o.GENERATOR]
    W = wire(P)
    assert engine.batch_on_curve(W).all()
    assert engine.batch_on_curve(W[:16], check_order=True).all()
    # break each clause of on_curve.rs:17-30 in turn
    x, y, z, t = P[0]
    bad = [(x, y, z, (t + 1) % Q),               # off the Segre embedding
           ((x + 1) % Q, y, z, t),               # off the curve
           (0, 0, 0, 0)]                         # Z = 0
    assert not engine.batch_on_curve(wire(bad)).any()
    # a curve point outside the image of decaf: G + T4 with T4 = (sqrt(-1), 0) of order 4;
    # [2r](G + T4) = [2]T4 = (0, -1) != identity, so only the order clause rejects it
    i = pow(o.ZETA, (Q - 1) // 4, Q)
    assert i * i % Q == Q - 1
    t4 = (i, 0, 1, 0)
    assert o.on_curve(t4)
    q4 = o.point_add(o.GENERATOR, t4)
    w4 = wire([q4])
    assert engine.batch_on_curve(w4)[0] == 1
    assert engine.batch_on_curve(w4, check_order=True)[0] == 0
    assert engine.Element(wire([o.GENERATOR])[0].tobytes()).is_on_curve()


# ---- untrusted limbs (ADVICE: fq_load) -------------------------------------------------------
def test_noncanonical_montgomery_limbs_are_reduced_not_trusted(engine):
    """Any 256-bit string is read as an integer mod q: x + kq behaves as x, and multiples of
    q (the input that could spin the binary-GCD inverse) give 0."""
    rnd = random.Random(5)
    xs = [rnd.randrange(Q) for _ in range(64)]
    ks = [rnd.randrange(1, 13) for _ in range(64)]
    A = mont(xs)
    raw = [(int.from_bytes(A[i].tobytes(), "little") + ks[i] * Q) for i in range(64)]
    raw = [v for v in raw if v < (1 << 256)]
    xs2 = [v % Q * o.MONT_R_INV_Q % Q for v in raw]
    N = canon(raw)
    ys = mont([rnd.randrange(Q) for _ in raw])
    yv = unmont(ys)
    assert unmont(engine.fq_batch_op(0, N, ys)) == [a * b % Q for a, b in zip(xs2, yv)]
    assert unmont(engine.fq_batch_op(8, N)) == [pow(a, -1, Q) for a in xs2]
    mult = canon([Q, 2 * Q, 3 * Q, 13 * Q])
    assert unmont(engine.fq_batch_op(8, mult)) == [0, 0, 0, 0]
    assert unmont(engine.fq_batch_op(9, mult)) == [0, 0, 0, 0]
    # an Element whose coordinates carry extra multiples of q compresses to the same bytes
    P = oracle_points("r2/nc", 8)
    W = wire(P)
    W2 = W.copy()
    for r in range(8):
        for c in range(4):
            v = int.from_bytes(W[r, 32 * c:32 * c + 32].tobytes(), "little") + (1 + (r + c) % 11) * Q
            if v < (1 << 256):
                W2[r, 32 * c:32 * c + 32] = np.frombuffer(v.to_bytes(32, "little"), np.uint8)
    assert np.array_equal(engine.batch_compress(W2), engine.batch_compress(W))
    sc = canon(oracle_scalars("r2/ncs", 8))
    assert engine.vartime_multiscalar_mul(sc, W2)[1].tobytes() == engine.vartime_multiscalar_mul(sc, W)[1].tobytes()


# ---- ABI robustness (ADVICE) -----------------------------------------------------------------
def test_msm_window_override_bounds(engine):
    from decaf377_b200._lib import D377Error
    for c in (3, 23, 24, 99):
        with pytest.raises(D377Error):
            engine.msm_set_window(c)
    n = 700
    P, s = oracle_points("r2/w", n), oracle_scalars("r2/w", n)
    want = engine.vartime_multiscalar_mul(canon(s), wire(P))[1].tobytes()
    for c in (4, 22):
        engine.msm_set_window(c)
        assert engine.vartime_multiscalar_mul(canon(s), wire(P))[1].tobytes() == want
    engine.msm_set_window(0)


def test_prepared_bases_registry(engine):
    from decaf377_b200 import _lib
    from decaf377_b200._lib import D377Error
    n = 300
    P, s = oracle_points("r2/b", n), oracle_scalars("r2/b", n)
    W, S = wire(P), canon(s)
    want = engine.vartime_multiscalar_mul(S, W)[1].tobytes()
    bases = engine.MsmBases(W)
    assert engine.vartime_multiscalar_mul(S, bases)[1].tobytes() == want
    assert engine.vartime_multiscalar_mul(S[:100], bases)[1].tobytes() == \
        engine.vartime_multiscalar_mul(S[:100], W[:100])[1].tobytes()
    lib = _lib.load()
    oe = np.empty(128, np.uint8)
    # more scalars than bases, and a pointer the library never handed out, are refused in C
    S2 = np.concatenate([S, S])
    rc = lib.d377_msm(S2.ctypes.data_as(C.c_void_p), C.c_void_p(bases.ptr), engine.PT_BASES, 2 * n,
                      oe.ctypes.data_as(C.c_void_p), None)
    assert rc == _lib.ERR_INVALID_ARG
    rc = lib.d377_msm(S.ctypes.data_as(C.c_void_p), C.c_void_p(bases.ptr + 128), engine.PT_BASES, 8,
                      oe.ctypes.data_as(C.c_void_p), None)
    assert rc == _lib.ERR_INVALID_ARG
    ptr = bases.ptr
    bases.close()
    assert lib.d377_msm_bases_destroy(C.c_void_p(ptr)) == _lib.ERR_INVALID_ARG   # already gone
    rc = lib.d377_msm(S.ctypes.data_as(C.c_void_p), C.c_void_p(ptr), engine.PT_BASES, 8,
                      oe.ctypes.data_as(C.c_void_p), None)
    assert rc == _lib.ERR_INVALID_ARG


def test_calls_from_other_host_threads(engine):
    """The engine may be driven from any host thread (CUDA's current device is per thread;
    every entry point selects the engine's device itself)."""
    n = 512
    P, s = oracle_points("r2/t", n), oracle_scalars("r2/t", n)
    W, S = wire(P), canon(s)
    want_c = engine.batch_compress(W)
    want_m = engine.vartime_multiscalar_mul(S, W)[1].tobytes()
    want_f = engine.fixed_base_mul(S[:64], engine.OUT_ENCODING)
    errs = []

    def work(k):
        try:
            for _ in range(3):
                assert np.array_equal(engine.batch_compress(W), want_c)
                assert engine.vartime_multiscalar_mul(S, W)[1].tobytes() == want_m
                assert np.array_equal(engine.fixed_base_mul(S[:64], engine.OUT_ENCODING), want_f)
        except Exception as ex:      # noqa: BLE001
            errs.append((k, repr(ex)))

    th = [threading.Thread(target=work, args=(k,)) for k in range(4)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs


# ---- asynchronous MSM ------------------------------------------------------------------------
def test_msm_async_matches_blocking_and_reports_errors_at_sync(engine):
    import torch

    from decaf377_b200 import device as dev
    from decaf377_b200._lib import D377Error, ERR_SCALAR_RANGE
    g = torch.Generator(device="cuda").manual_seed(21)
    for n in (1, 33, 5000, 1 << 17, (1 << 20) + 77):
        r = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device="cuda", generator=g)
        sc = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device="cuda", generator=g)
        sc[:, 31] &= 0x03
        el = dev.encode_to_curve(r, engine.OUT_ELEMENT)
        engine.sync()
        want_e, want_c = dev.msm(sc, el)
        outs = []
        for k in range(5):       # back to back: tails overlap the next head, sorts prefetch
            outs.append(dev.msm_async(sc, el, inputs_ready=(k % 2 == 1)))
        engine.sync()
        for oe, oc in outs:
            assert torch.equal(oc, want_c)
            assert torch.equal(dev.compress(oe.reshape(1, 128))[0], want_c)
        # different MSMs in flight at once do not share state
        h = n // 2
        a = dev.msm_async(sc[:h], el[:h], inputs_ready=True) if h else None
        b = dev.msm_async(sc[h:], el[h:], inputs_ready=True)
        engine.sync()
        parts = [x[0] for x in (a, b) if x is not None]
        _, enc = dev.element_sum(torch.stack(parts))
        engine.sync()
        assert torch.equal(enc, want_c)
    # MSMs of very different sizes in flight together: the workspace sets are regrown while
    # the other set's tail is still running
    sizes = [5000, 1 << 17, 33, (1 << 19) + 5, 7, 1 << 18, 5000]
    ins, wants, gots = [], [], []
    for m in sizes:
        r_m = torch.randint(0, 256, (m, 32), dtype=torch.uint8, device="cuda", generator=g)
        s_m = torch.randint(0, 256, (m, 32), dtype=torch.uint8, device="cuda", generator=g)
        s_m[:, 31] &= 0x03
        ins.append((s_m, dev.encode_to_curve(r_m, engine.OUT_ELEMENT)))
    engine.sync()
    for s_m, p_m in ins:
        wants.append(dev.msm(s_m, p_m)[1].clone())
    for rep in range(2):
        for s_m, p_m in ins:
            gots.append(dev.msm_async(s_m, p_m, inputs_ready=True))
    engine.sync()
    for k, (oe_k, oc_k) in enumerate(gots):
        assert torch.equal(oc_k, wants[k % len(sizes)]), (k, sizes[k % len(sizes)])
    # a scalar >= r is reported by the next sync, once
    bad = sc.clone()
    bad[7] = 0xFF
    dev.msm_async(bad, el)
    with pytest.raises(D377Error) as ei:
        engine.sync()
    assert ei.value.code == ERR_SCALAR_RANGE
    engine.sync()
    # with the overlap switched off the same calls run on the engine stream alone
    engine.msm_set_tail_overlap(False)
    oe, oc = dev.msm_async(sc, el)
    engine.sync()
    assert torch.equal(oc, want_c)
    engine.msm_set_tail_overlap(True)


# ---- scalars in the reference's in-memory (Montgomery) form ----------------------------------------
def test_scalars_in_montgomery_form(engine):
    """D377_SCALARS_MONTGOMERY: Fr limbs as fr/u64/wrapper.rs keeps them (x * 2^256 mod r) are
    converted on the GPU (Fr::into_bigint, fr/arkworks.rs:36-57) -- every entry point that
    takes scalars gives the bytes it gives for the canonical form."""
    import torch

    from decaf377_b200 import device as dev
    n = 777
    P = oracle_points("r2/m", n)
    s = oracle_scalars("r2/m", n - 3) + [0, 1, R - 1]
    W, S = wire(P), canon(s)
    SM = canon([x * (1 << 256) % R for x in s])
    want = engine.vartime_multiscalar_mul(S, W)[1].tobytes()
    assert want == o.compress(o.vartime_multiscalar_mul(s, P))
    assert engine.vartime_multiscalar_mul(SM, W, scalars_montgomery=True)[1].tobytes() == want
    assert engine.vartime_multiscalar_mul(SM, engine.batch_compress(W), engine.PT_ENCODING,
                                          scalars_montgomery=True)[1].tobytes() == want
    engine.msm_submit(SM, W, slot=0, scalars_montgomery=True)
    assert engine.msm_wait(0)[1].tobytes() == want
    bases = engine.MsmBases(W)
    assert engine.vartime_multiscalar_mul(SM, bases, scalars_montgomery=True)[1].tobytes() == want
    bases.close()
    engine.init_multi([0])
    assert engine.msm_multi(SM, W, ngpu=1, scalars_montgomery=True)[1].tobytes() == want
    d_sm, d_w = torch.from_numpy(SM).cuda(), torch.from_numpy(W).cuda()
    assert dev.msm(d_sm, d_w, scalars_montgomery=True)[1].cpu().numpy().tobytes() == want
    oe, oc = dev.msm_async(d_sm, d_w, scalars_montgomery=True, inputs_ready=True)
    engine.sync()
    assert oc.cpu().numpy().tobytes() == want
    # element-wise entry points
    assert np.array_equal(engine.batch_scalar_mul(W, SM, out_format=engine.OUT_ENCODING, scalars_montgomery=True),
                          engine.batch_scalar_mul(W, S, out_format=engine.OUT_ENCODING))
    for fmt in (engine.OUT_ENCODING, engine.OUT_ELEMENT):
        a = engine.fixed_base_mul(SM, fmt, scalars_montgomery=True)
        b = engine.fixed_base_mul(S, fmt)
        assert np.array_equal(a if fmt == engine.OUT_ENCODING else engine.batch_compress(a),
                              b if fmt == engine.OUT_ENCODING else engine.batch_compress(b))
    # every 256-bit string is a Montgomery representative of something: no range error
    rnd = random.Random(9)
    raw = [(1 << 256) - 1, R, R + 1] + [rnd.getrandbits(256) for _ in range(60)]
    vals = [v * pow(1 << 256, -1, R) % R for v in raw]
    got = engine.fixed_base_mul(canon(raw), engine.OUT_ENCODING, scalars_montgomery=True)
    for i, k in enumerate(vals):
        assert got[i].tobytes() == o.compress(o.scalar_mul(o.GENERATOR, k)), i
    assert engine.Fr(5).to_montgomery_bytes() == (5 * (1 << 256) % R).to_bytes(32, "little")


# ---- many small MSMs in one call ------------------------------------------------------------------
def test_batch_msm_matches_per_segment_oracle(engine):
    """d377_batch_msm: Element::vartime_multiscalar_mul (element/projective.rs:99-117) over many
    independent segments -- empty, single-pair, ragged and a few hundred pairs long."""
    rnd = random.Random(17)
    sizes = [0, 1, 3, 0, 2, 31, 32, 33, 64, 100, 257, 1, 0]
    n = sum(sizes)
    P, s = oracle_points("r2/bm", n), oracle_scalars("r2/bm", n)
    s[5] = 0
    s[7] = R - 1
    W, S = wire(P), canon(s)
    off = np.cumsum([0] + sizes).astype(np.uint32)
    want = []
    for j, m in enumerate(sizes):
        lo = int(off[j])
        want.append(o.compress(o.vartime_multiscalar_mul(s[lo:lo + m], P[lo:lo + m])))
    enc = engine.batch_msm(S, W, off, out_format=engine.OUT_ENCODING)
    assert [enc[j].tobytes() for j in range(len(sizes))] == want
    el, ok = engine.batch_msm(S, W, off, return_ok=True)
    assert ok.all() and np.array_equal(engine.batch_compress(el), enc)
    # the same segments through the single-MSM entry point
    for j in (2, 8, 10):
        lo, hi = int(off[j]), int(off[j + 1])
        assert engine.vartime_multiscalar_mul(S[lo:hi], W[lo:hi])[1].tobytes() == want[j]
    # Montgomery scalars, affine and encoding inputs
    SM = canon([x * (1 << 256) % R for x in s])
    assert np.array_equal(engine.batch_msm(SM, W, off, out_format=engine.OUT_ENCODING, scalars_montgomery=True), enc)
    assert np.array_equal(engine.batch_msm(S, engine.batch_normalize(W), off, engine.PT_AFFINE, engine.OUT_ENCODING), enc)
    E = engine.batch_compress(W)
    enc2, ok2 = engine.batch_msm(S, E, off, engine.PT_ENCODING, engine.OUT_ENCODING, return_ok=True)
    assert ok2.all() and np.array_equal(enc2, enc)
    # an invalid encoding poisons its own segment only
    bad = E.copy()
    lo9 = int(off[9])
    bad[lo9 + 4] = 0
    bad[lo9 + 4, 0] = 1                       # s = 1: InvalidEncoding (tests/encoding.rs:28-52)
    enc3, ok3 = engine.batch_msm(S, bad, off, engine.PT_ENCODING, engine.OUT_ENCODING, return_ok=True)
    assert [int(v) for v in ok3] == [0 if j == 9 else 1 for j in range(len(sizes))]
    assert all(enc3[j].tobytes() == want[j] for j in range(len(sizes)) if j != 9)
    # argument checks
    from decaf377_b200._lib import D377Error
    with pytest.raises(D377Error):
        engine.batch_msm(S, W, np.array([0, 5, 3], np.uint32))
    with pytest.raises(ValueError):
        engine.batch_msm(S, W, np.array([0, n + 1], np.uint32))
    # throughput shape: 4096 MSMs of 64 pairs in one call equal 4096 separate results on a sample
    m, k = 4096, 64
    raw = np.frombuffer(o.xof_bytes("r2/bm_big", m * k), np.uint8).reshape(m * k, 32).copy()
    sc = np.frombuffer(o.xof_bytes("r2/bm_big_s", m * k), np.uint8).reshape(m * k, 32).copy()
    sc[:, 31] &= 3
    pts = engine.batch_encode_to_curve(raw)
    offs = (np.arange(m + 1) * k).astype(np.uint32)
    out = engine.batch_msm(sc, pts, offs, out_format=engine.OUT_ENCODING)
    for j in (0, 1, 777, m - 1):
        assert out[j].tobytes() == engine.vartime_multiscalar_mul(sc[j * k:(j + 1) * k], pts[j * k:(j + 1) * k])[1].tobytes()
    assert out[5].tobytes() == co.msm_pippenger(sc[5 * k:6 * k], pts[5 * k:6 * k], threads=4)[1].tobytes()


# ---- multi-GPU inside one process ----------------------------------------------------------------
def _dot_mod_r(a, s):
    tot = 0
    for x, y in zip(a, s):
        tot += x * y
    return tot % R


def test_msm_multi_one_process(engine):
    """d377_msm_multi / d377_msm_multi_dev with the known-answer construction of SURVEY 8d:
    P_i = a_i G  =>  sum s_i P_i = (sum a_i s_i mod r) G.  Runs on however many GPUs the box
    has (ngpu = 1 exercises the same workers, peer copy and final sum)."""
    import torch

    from decaf377_b200 import device as dev
    ndev = min(torch.cuda.device_count(), 8)
    engine.init_multi(list(range(ndev)))
    assert engine.device_list()[:ndev] == list(range(ndev))
    n = 40000 + 7
    a, s = oracle_scalars("r2/ma", n), oracle_scalars("r2/ms", n)
    A, S = canon(a), canon(s)
    P = engine.fixed_base_mul(A, engine.OUT_ELEMENT)
    want = o.compress(o.scalar_mul(o.GENERATOR, _dot_mod_r(a, s)))
    assert engine.vartime_multiscalar_mul(S, P)[1].tobytes() == want
    for ngpu in sorted({1, ndev, max(1, ndev // 2)}):
        el, enc = engine.msm_multi(S, P, engine.PT_ELEMENT, ngpu=ngpu)
        assert enc.tobytes() == want, ngpu
        assert engine.batch_compress(el.reshape(1, 128))[0].tobytes() == want
        # other input formats shard the same way
        assert engine.msm_multi(S, engine.batch_normalize(P), engine.PT_AFFINE, ngpu=ngpu)[1].tobytes() == want
        assert engine.msm_multi(S, P[:, :96].copy(), engine.PT_XYZ, ngpu=ngpu)[1].tobytes() == want
    # ragged and empty slices
    for m in (0, 1, ndev - 1, ndev + 1):
        got = engine.msm_multi(S[:m], P[:m], ngpu=ndev)[1].tobytes()
        assert got == engine.vartime_multiscalar_mul(S[:m], P[:m])[1].tobytes(), m
    # device-resident slices, one per GPU
    sl = [(k * n // ndev, (k + 1) * n // ndev) for k in range(ndev)]
    sc_d = [torch.from_numpy(S[lo:hi]).to("cuda:%d" % k) for k, (lo, hi) in enumerate(sl)]
    pt_d = [torch.from_numpy(P[lo:hi]).to("cuda:%d" % k) for k, (lo, hi) in enumerate(sl)]
    assert dev.msm_multi(sc_d, pt_d)[1].tobytes() == want
    # asynchronous form: six calls enqueued back to back (more than the four gather areas),
    # results only read after multi_sync
    outs = [dev.msm_multi_async(sc_d, pt_d) for _ in range(6)]
    half = [(t[: t.shape[0] // 2], u[: u.shape[0] // 2]) for t, u in zip(sc_d, pt_d)]
    out_half = dev.msm_multi_async([h[0] for h in half], [h[1] for h in half])
    dev.multi_sync()
    for oe, oc in outs:
        assert oc.cpu().numpy().tobytes() == want
    want_half = o.compress(o.scalar_mul(o.GENERATOR, sum(
        _dot_mod_r(a[lo:lo + (hi - lo) // 2], s[lo:lo + (hi - lo) // 2]) for lo, hi in sl) % R))
    assert out_half[1].cpu().numpy().tobytes() == want_half
    # blocking calls in between asynchronous ones that are still in flight: the blocking call's
    # gather area is not one of the ring's, so neither disturbs the other
    mixed = []
    for _ in range(3):
        mixed.append(dev.msm_multi_async(sc_d, pt_d))
        assert dev.msm_multi([h[0] for h in half], [h[1] for h in half])[1].tobytes() == want_half
    dev.multi_sync()
    for oe, oc in mixed:
        assert oc.cpu().numpy().tobytes() == want
    # errors travel back from the worker threads
    from decaf377_b200._lib import D377Error, ERR_SCALAR_RANGE
    bad = S.copy()
    bad[n - 1] = 0xFF
    with pytest.raises(D377Error) as ei:
        engine.msm_multi(bad, P, ngpu=ndev)
    assert ei.value.code == ERR_SCALAR_RANGE
    bad_d = [t.clone() for t in sc_d]
    bad_d[-1][0] = 0xFF
    dev.msm_multi_async(bad_d, pt_d)
    with pytest.raises(D377Error) as ei:
        dev.multi_sync()
    assert ei.value.code == ERR_SCALAR_RANGE
    dev.multi_sync()
    # per-thread device selection: the same call on the last device
    if ndev > 1:
        engine.set_device(ndev - 1)
        assert engine.get_device() == ndev - 1
        assert engine.vartime_multiscalar_mul(S[:999], P[:999])[1].tobytes() == \
            o.compress(o.scalar_mul(o.GENERATOR, _dot_mod_r(a[:999], s[:999])))
        engine.set_device(-1)
    assert engine.get_device() == 0


# ---- fixed base: table choice (round-1 weak #9) --------------------------------------------------
_FB_SMALL = r"""
import sys, numpy as np, torch
sys.path.insert(0, %r)
import decaf377_b200 as d
from oracle import decaf377_ref as o
d.init(0)
torch.cuda.synchronize()
free0, _ = torch.cuda.mem_get_info()
sc = np.frombuffer(o.xof_bytes("fb_small", 4), np.uint8).reshape(4, 32).copy()
sc[:, 31] &= 3
got = d.fixed_base_mul(sc, d.OUT_ENCODING)
free1, _ = torch.cuda.mem_get_info()
for i in range(4):
    assert got[i].tobytes() == o.compress(o.scalar_mul(o.GENERATOR, int.from_bytes(sc[i].tobytes(), "little")))
print("GREW", (free0 - free1) >> 20)
big = np.frombuffer(o.xof_bytes("fb_big", 1 << 18), np.uint8).reshape(1 << 18, 32).copy()
big[:, 31] &= 3
e1 = d.fixed_base_mul(big, d.OUT_ENCODING)            # builds the quartic table
free2, _ = torch.cuda.mem_get_info()
print("GREW2", (free1 - free2) >> 20)
assert np.array_equal(d.fixed_base_mul(sc, d.OUT_ENCODING), got)   # now through the quartic table
assert np.array_equal(d.batch_compress(d.fixed_base_mul(big[:4096], d.OUT_ELEMENT)), e1[:4096])
"""


def test_small_fixed_base_call_does_not_build_the_quartic_table():
    r = subprocess.run([sys.executable, "-c", _FB_SMALL % ROOT], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    grew = int(r.stdout.split("GREW ")[1].split()[0])
    grew2 = int(r.stdout.split("GREW2 ")[1].split()[0])
    assert grew < 64 + 48, "a 4-scalar call grew device memory by %d MiB" % grew   # 48 MiB Edwards table
    assert grew2 > 1500, "the large call should have built the 1.6 GB quartic table (%d MiB)" % grew2


# ---- the TMA-staged operand stream (experiment kept behind D377_MSM_ACC_TMA) ---------------------
_TMA = r"""
import sys, numpy as np
sys.path.insert(0, %r)
import decaf377_b200 as d
from oracle import c_oracle as co
from oracle import decaf377_ref as o
d.init(0)
for n in (1, 33, 4097, 1 << 17):
    raw = np.frombuffer(o.xof_bytes("tma_fq/%%d" %% n, n), np.uint8).reshape(n, 32).copy()
    sc = np.frombuffer(o.xof_bytes("tma_sc/%%d" %% n, n), np.uint8).reshape(n, 32).copy()
    sc[:, 31] &= 3
    el = d.batch_encode_to_curve(raw, d.OUT_ELEMENT)
    d.msm_set_normalize(1)          # affine bucket operands: the path the TMA kernel serves
    got = d.vartime_multiscalar_mul(sc, el)[1].tobytes()
    want = co.msm_pippenger(sc, el, threads=8)[1].tobytes()
    assert got == want, n
print("TMA_OK")
"""


def test_tma_staged_accumulation_is_bit_exact():
    """k_msm_accumulate_tma (cp.async.bulk per record + one mbarrier per warp and stage) gives the
    results of the register-staged kernel; it is slower and therefore off by default."""
    env = dict(os.environ, D377_MSM_ACC_TMA="1")
    r = subprocess.run([sys.executable, "-c", _TMA % ROOT], capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0 and "TMA_OK" in r.stdout, r.stdout + r.stderr


# ---- the boundary from a non-Python host ---------------------------------------------------------
def test_c_host_program_drives_the_abi(tmp_path):
    """tests/abi_smoke.c is compiled with gcc against include/decaf377_b200.h and linked with
    the shared library: the same calls, in the same order, as the Rust shim of rust/src/gpu.rs."""
    exe = tmp_path / "abi_smoke"
    libdir = os.path.join(ROOT, "decaf377_b200")
    cmd = ["gcc", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "abi_smoke.c"), "-o", str(exe),
           "-L", libdir, "-ldecaf377_b200", "-Wl,-rpath," + libdir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "abi_smoke: all checks passed" in r.stdout


# ---- debug build (on_curve.rs kept alive the way the reference's CI does) ------------------------
def test_parity_suite_under_the_on_curve_debug_build():
    lib = os.path.join(ROOT, "decaf377_b200", "libdecaf377_b200_dbg.so")
    if not os.path.exists(lib):
        pytest.skip("debug library not built (python -m decaf377_b200.build --debug)")
    env = dict(os.environ, D377_DEBUG_LIB="1", D377_DEBUG_REPORT="1")
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", "-p", "no:cacheprovider",
                        os.path.join(ROOT, "tests", "test_gpu_parity.py"),
                        os.path.join(ROOT, "tests", "test_gpu_round2.py"),
                        "-k", "not debug_build and not c_host and not quartic_table and not tma_staged"],
                       capture_output=True, text=True, timeout=3000, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("D377_DEBUG ")]
    assert line, r.stdout[-2000:]
    fields = dict(kv.split("=") for kv in line[-1].split()[1:])
    assert int(fields["build"]) == 2
    assert int(fields["checked"]) > 10000 and int(fields["failures"]) == 0, line[-1]
