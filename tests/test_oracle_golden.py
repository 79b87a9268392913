"""Pin the CPU oracle (oracle/decaf377_ref.py) against every golden vector the
reference's own tests hold for the hot path.  CPU only."""
import random

from oracle import decaf377_ref as o
from tests import golden_vectors as gv

Q, R = o.Q, o.R


def test_generator_multiples():
    # tests/encoding.rs:54-95
    acc = o.IDENTITY
    for h in gv.GENERATOR_MULTIPLES:
        b = bytes.fromhex(h)
        p = o.decompress(b)
        assert p is not None and o.on_curve(p)
        assert o.compress(p).hex() == h
        assert o.point_eq(acc, p)
        assert o.compress(acc).hex() == h
        acc = o.point_add(acc, o.GENERATOR)


def test_identity_and_generator():
    # tests/encoding.rs:19-52
    assert o.compress(o.IDENTITY) == bytes(32)
    assert o.point_eq(o.decompress(bytes(32)), o.IDENTITY)
    first = next(b for b in range(1, 256) if o.decompress(bytes([b]) + bytes(31)) is not None)
    assert first == 8
    g = o.decompress(bytes([8]) + bytes(31))
    assert o.point_eq(g, o.GENERATOR)
    assert o.to_affine(g) == (o.B_X, o.B_Y)


def test_elligator_vectors():
    # src/ark_curve/elligator.rs:86-208
    for inp, xy in zip(gv.ELLIGATOR_INPUTS, gv.ELLIGATOR_XY):
        r0 = o.fq_from_bytes_checked(bytes(inp))
        assert r0 is not None
        p = o.elligator_map(r0)
        assert o.on_curve(p)
        assert o.to_affine(p) == xy


def test_edge_encodings():
    for b, valid in gv.EDGE_CASES:
        p = o.decompress(b)
        assert (p is not None) == valid, b.hex()
        if p is not None:
            assert o.compress(p) == b          # tests/encoding.rs:97-106


def test_field_byte_conventions():
    # fields/fq.rs:149-153, fr.rs:129-133, fq/arkworks.rs:603-673
    assert o.fq_from_bytes_checked(bytes(32)) == 0
    assert o.fq_from_bytes_checked(b"\xff" * 32) is None
    assert o.fr_from_bytes_checked(bytes(32)) == 0
    assert o.fr_from_bytes_checked(b"\xff" * 32) is None
    assert o.fq_from_le_bytes_mod_order((Q + 1).to_bytes(32, "little")) == 1
    assert o.fq_from_le_bytes_mod_order(b"\x01" + bytes(79)) == 1
    assert (Q - 1) * (Q - 1) % Q == 1
    # wide inputs: 32-byte chunk Horner of fq.rs:90-102 equals plain reduction
    rnd = random.Random(7)
    for _ in range(50):
        b = bytes(rnd.getrandbits(8) for _ in range(80))
        acc = 0
        for k in reversed(range(0, 80, 32)):
            acc = (acc * (1 << 256) + int.from_bytes(b[k:k + 32], "little")) % Q
        assert acc == o.fq_from_le_bytes_mod_order(b)


def test_sqrt_ratio_zeta_contract():
    # ark_curve/invsqrt.rs:182-211 + proptest-regressions/invsqrt.txt:7
    assert o.sqrt_ratio_zeta(1, 1) in ((True, 1), (True, Q - 1))
    assert o.sqrt_ratio_zeta(0, 5) == (True, 0)
    assert o.sqrt_ratio_zeta(5, 0) == (False, 0)
    rnd = random.Random(11)
    for _ in range(300):
        u, v = rnd.randrange(1, Q), rnd.randrange(1, Q)
        ok, root = o.sqrt_ratio_zeta(u, v)
        lhs = root * root % Q * v % Q
        assert lhs == (u if ok else o.ZETA * u % Q)
        assert ok == (pow(u * pow(v, -1, Q) % Q, (Q - 1) // 2, Q) == 1)


def test_group_laws_and_scalar_mul():
    # tests/operations.rs:19-61, min_curve/element.rs:343-391
    G = o.GENERATOR
    assert o.point_eq(o.point_add(G, G), o.point_double(G))
    assert o.point_eq(o.scalar_mul(G, 1), G)
    assert o.is_identity(o.scalar_mul(G, 0))
    assert o.point_eq(o.scalar_mul(G, R - 1), o.point_neg(G))
    assert o.is_identity(o.point_add(G, o.point_neg(G)))
    assert o.is_identity(o.scalar_mul(G, R))
    rnd = random.Random(13)
    for i in range(5):
        a, b, c = (rnd.randrange(R) for _ in range(3))
        P, Qp, Rp = (o.encode_to_curve(rnd.randrange(Q)) for _ in range(3))
        assert o.point_eq(o.point_add(o.scalar_mul(P, a), o.scalar_mul(P, b)),
                          o.scalar_mul(P, (a + b) % R))
        assert o.point_eq(o.scalar_mul(o.scalar_mul(P, a), b), o.scalar_mul(P, a * b % R))
        want = o.point_add(o.point_add(o.scalar_mul(P, a), o.scalar_mul(Qp, b)), o.scalar_mul(Rp, c))
        assert o.point_eq(o.vartime_multiscalar_mul([a, b, c], [P, Qp, Rp]), want)
    assert o.is_identity(o.vartime_multiscalar_mul([], []))


def test_decompress_compress_roundtrip_random_bytes():
    n_ok = 0
    for b in o.xof_blocks("raw", 400):
        b = b[:31] + bytes([b[31] & 0x1F])
        p = o.decompress(b)
        if p is not None:
            n_ok += 1
            assert o.on_curve(p)
            assert o.compress(p) == b
    assert 20 < n_ok < 200


def test_wire_format_roundtrip():
    p = o.encode_to_curve(12345)
    assert o.point_from_wire(o.point_to_wire(p)) == p
    # Fq::ONE montgomery limbs, fq/u32/wrapper.rs:108-110
    assert o.fq_to_mont_bytes(1).hex() == \
        "f3ffffffff7f1c7df2ffff6f0ff557720ee0f2c517515d8169d9abbb2b32da4b0d"[:0] + \
        (0x0d4bda322bbb9a9d16d81575512c0fee7257f50f6ffffff27d1c7ffffffffff3).to_bytes(32, "little").hex()
