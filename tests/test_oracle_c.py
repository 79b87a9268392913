"""The C restatement (oracle/d377_oracle.c) against the reference's golden vectors
and against the Python big-int oracle.  CPU only."""
import random

import numpy as np

from oracle import c_oracle as co
from oracle import decaf377_ref as o
from tests import golden_vectors as gv
from tests.util import canon, mont, np_bytes, oracle_points, oracle_scalars, unmont, unwire, wire

Q, R = o.Q, o.R


def test_generator_multiples_and_edges():
    enc = np_bytes([bytes.fromhex(h) for h in gv.GENERATOR_MULTIPLES], 32)
    el, ok = co.decompress(enc)
    assert ok.all()
    assert np.array_equal(co.compress(el), enc)
    acc = wire([o.IDENTITY])
    gen = wire([o.GENERATOR])
    for i in range(16):
        assert co.compress(acc)[0].tobytes().hex() == gv.GENERATOR_MULTIPLES[i]
        acc = co.add(acc, gen)
    edge = np_bytes(gv.EDGE_ENCODINGS, 32)
    _, ok = co.decompress(edge)
    assert [bool(x) for x in ok] == [v for _, v in gv.EDGE_CASES]


def test_elligator_vectors():
    el = unwire(co.encode_to_curve(np_bytes([bytes(v) for v in gv.ELLIGATOR_INPUTS], 32)))
    for p, xy in zip(el, gv.ELLIGATOR_XY):
        assert o.to_affine(p) == xy


def test_field_and_sqrt_match_python():
    rnd = random.Random(21)
    a = [0, 1, Q - 1] + [rnd.randrange(Q) for _ in range(500)]
    b = [Q - 1, 0, Q - 1] + [rnd.randrange(Q) for _ in range(500)]
    assert unmont(co.fq_mul(mont(a), mont(b))) == [x * y % Q for x, y in zip(a, b)]
    out, ws = co.sqrt_ratio_zeta(mont(a), mont(b))
    for x, y, r, w in zip(a, b, unmont(out), ws):
        assert (bool(w), r) == o.sqrt_ratio_zeta(x, y)


def test_codec_and_group_match_python():
    raw = [b[:31] + bytes([b[31] & 0x1F]) for b in o.xof_blocks("raw", 300)]
    el, ok = co.decompress(np_bytes(raw, 32), threads=4)
    for i, b in enumerate(raw):
        ref = o.decompress(b)
        assert bool(ok[i]) == (ref is not None)
        if ref is not None:
            assert o.point_from_wire(el[i].tobytes()) == ref
    r1, r2 = o.xof_blocks("fq", 64), o.xof_blocks("fq2", 64)
    henc = co.hash_to_curve(np_bytes(r1, 32), np_bytes(r2, 32), out_enc=True, threads=2)
    for i in range(64):
        assert henc[i].tobytes() == o.compress(o.hash_to_curve(o.fq_from_le_bytes_mod_order(r1[i]),
                                                               o.fq_from_le_bytes_mod_order(r2[i])))
    pts = oracle_points("pt", 32)
    sc = oracle_scalars("sc", 32)
    sc[:3] = [0, 1, R - 1]
    got = co.scalar_mul(wire(pts), canon(sc), out_enc=True, threads=4)
    encs = np_bytes([o.compress(p) for p in pts], 32)
    got2, ok = co.pipeline(encs, canon(sc), threads=4)
    fb = co.fixed_base(canon(sc), threads=4)
    for i in range(32):
        assert got[i].tobytes() == o.compress(o.scalar_mul(pts[i], sc[i]))
        assert fb[i].tobytes() == o.compress(o.scalar_mul(o.GENERATOR, sc[i]))
    assert ok.all() and np.array_equal(got, got2)


def test_msm_fold_and_pippenger():
    for n in (0, 1, 3, 40, 300):
        pts = oracle_points("msm_pt", n)
        sc = oracle_scalars("msm_sc", n)
        want = o.compress(o.vartime_multiscalar_mul(sc, pts))
        S = canon(sc) if n else np.zeros((0, 32), np.uint8)
        P = wire(pts) if n else np.zeros((0, 128), np.uint8)
        assert co.msm_fold(S, P)[1].tobytes() == want
        assert co.msm_pippenger(S, P, threads=4)[1].tobytes() == want
