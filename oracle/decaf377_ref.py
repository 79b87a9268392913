"""CPU oracle (big-int restatement) of the decaf377 batch hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``decaf377_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` leg use it, and only as the checker.

Each function cites the reference file:line (relative to the reference tree,
crate v0.10.1) whose algorithm it restates.  The reference itself cannot be
built here (no cargo/rustc; arkworks 0.4 is not vendored), so this restatement
is pinned against the reference's own golden vectors in
``tests/test_oracle_golden.py``:

* 16 encodings of i*G          tests/encoding.rs:61-78
* identity / generator / s=1..7 tests/encoding.rs:19-52
* 8 Elligator (x, y) vectors   src/ark_curve/elligator.rs:88-188
* Fq/Fr byte conventions       src/fields/fq.rs:149-153, fq/arkworks.rs:603-673

Arithmetic is Python ``int`` modulo q / r (canonical, non-Montgomery).
"""
from __future__ import annotations

import hashlib
from typing import Iterable, List, Optional, Sequence, Tuple

# ---------------------------------------------------------------------------
# Fields.  src/fields/fq.rs:29-34 (MODULUS_LIMBS), src/fields/fr.rs:29-34
# ---------------------------------------------------------------------------
Q = 0x12AB655E9A2CA55660B44D1E5C37B00159AA76FED00000010A11800000000001
R = 0x04AAD957A68B2955982D1347970DEC005293A3AFC43C8AFEB95AEE9AC33FD9FF

assert Q == sum(l << (64 * i) for i, l in enumerate(
    [725501752471715841, 6461107452199829505, 6968279316240510977, 1345280370688173398]))
assert R == 2111115437357092606062206234695386632838870926408408195193685246394721360383

MONT_R = 1 << 256            # Montgomery radix used by both reference backends
MONT_R_INV_Q = pow(MONT_R, -1, Q)


def from_mont_limbs64(limbs: Sequence[int], mod: int = Q) -> int:
    """Canonical value of a reference ``from_montgomery_limbs([u64;4])`` constant."""
    v = sum(l << (64 * i) for i, l in enumerate(limbs))
    return v * pow(MONT_R, -1, mod) % mod


# src/ark_curve/constants.rs:20-25 (ZETA), min_curve/constants.rs:3-39
ZETA = from_mont_limbs64([5947794125541564500, 11292571455564096885,
                          11814268415718120036, 155746270000486182])
COEFF_A = Q - 1
COEFF_D = 3021
COEFF_K = 6042               # -2d/a
assert from_mont_limbs64([10157024534604021774, 16668528035959406606,
                          5322190058819395602, 387181115924875961]) == COEFF_A
assert from_mont_limbs64([15008245758212136496, 17341409599856531410,
                          648869460136961410, 719771289660577536]) == COEFF_D
assert from_mont_limbs64([10844245690243005535, 9774967673803681700,
                          12776203677742963460, 94262208632981673]) == COEFF_K

# sqrt constants: src/ark_curve/constants.rs:28-58
SQRT_N = 47
SQRT_M = (Q - 1) >> SQRT_N
assert SQRT_M == 60001509534603559531609739528203892656505753216962260608619555
M_MINUS_ONE_DIV_TWO = (SQRT_M - 1) // 2
assert M_MINUS_ONE_DIV_TWO == 30000754767301779765804869764101946328252876608481130304309777
SQRT_G = pow(ZETA, SQRT_M, Q)
assert SQRT_G == 4732611889701835744065511820927274956354524915951001256593514693060564426294
ZETA_TO_ONE_MINUS_M_DIV_TWO = \
    6762755396584113496485389421189479608933826763106393667349575256979972066439
assert ZETA_TO_ONE_MINUS_M_DIV_TWO == pow(pow(ZETA, (SQRT_M - 1) // 2, Q), -1, Q)
SQRT_W = 8

# basepoint: src/ark_curve/constants.rs:61-79
B_X = from_mont_limbs64([5825153684096051627, 16988948339439369204,
                         186539475124256708, 1230075515893193738])
B_Y = from_mont_limbs64([9786171649960077610, 13527783345193426398,
                         10983305067350511165, 1251302644532346138])
B_T = from_mont_limbs64([7466800842436274004, 14314110021432015475,
                         14108125795146788134, 1305086759679105397])
assert B_T == B_X * B_Y % Q


# ---------------------------------------------------------------------------
# Byte I/O.  src/fields/fq.rs:90-119, src/fields/fr.rs:90-119
# ---------------------------------------------------------------------------
def fq_from_le_bytes_mod_order(b: bytes) -> int:
    """fq.rs:90-102 (32-byte chunks folded with 2^256 mod q) == int(b) mod q."""
    return int.from_bytes(b, "little") % Q


FIELD_SIZE_POWER_OF_TWO = from_mont_limbs64([2726216793283724667, 14712177743343147295,
                                             12091039717619697043, 81024008013859129])
assert FIELD_SIZE_POWER_OF_TWO == (1 << 256) % Q     # fq.rs:83-88


def fq_from_le_bytes_mod_order_chunked(b: bytes) -> int:
    """fq.rs:90-102 as written: 32-byte little-endian chunks (the last one zero-padded),
    each read as a raw 256-bit value mod q, folded from the most significant chunk down
    with acc = acc * FIELD_SIZE_POWER_OF_TWO + chunk.  The reference's property test
    (fq/arkworks.rs:586-593, 80-byte inputs) pins it to the naive reduction; so does
    tests/test_oracle_golden.py."""
    chunks = [b[i:i + 32] for i in range(0, len(b), 32)]
    acc = 0
    for c in reversed(chunks):
        x = int.from_bytes(c + b"\0" * (32 - len(c)), "little") % Q
        acc = (acc * FIELD_SIZE_POWER_OF_TWO + x) % Q
    return acc


def fr_from_le_bytes_mod_order(b: bytes) -> int:
    return int.from_bytes(b, "little") % R


def fq_from_bytes_checked(b: bytes) -> Optional[int]:
    """fq.rs:108-115: None unless the 32 bytes are the canonical encoding."""
    assert len(b) == 32
    v = int.from_bytes(b, "little")
    return v if v < Q else None


def fr_from_bytes_checked(b: bytes) -> Optional[int]:
    assert len(b) == 32
    v = int.from_bytes(b, "little")
    return v if v < R else None


def fq_to_bytes(x: int) -> bytes:
    return (x % Q).to_bytes(32, "little")


def fr_to_bytes(x: int) -> bytes:
    return (x % R).to_bytes(32, "little")


def is_negative(x: int) -> bool:
    """src/sign.rs:19-23: LSB of the canonical value."""
    return (x & 1) == 1


def fq_abs(x: int) -> int:
    """src/sign.rs:10-16."""
    return (Q - x) % Q if is_negative(x) else x


# ---------------------------------------------------------------------------
# sqrt_ratio_zeta.  src/ark_curve/invsqrt.rs:14-166 (Sarkar 2020 tables),
# spec sqrt_alg.sage:35-109.  Followed step by step so that the returned root
# is the very element the reference returns (every step is exact field
# arithmetic, hence representation independent).
# ---------------------------------------------------------------------------
class _SqrtTables:
    def __init__(self) -> None:
        # invsqrt.rs:27-38: key g^(-nu * 2^(n-w)) -> nu
        self.s_lookup = {}
        for nu in range(256):
            g_pow = pow(SQRT_G, nu << (SQRT_N - SQRT_W), Q)
            self.s_lookup[pow(g_pow, -1, Q)] = nu
        # invsqrt.rs:40-49: gtab[k][nu] = g^(nu * 2^k), k in 0,8,..,40
        self.g = {k: [pow(SQRT_G, nu << k, Q) for nu in range(256)]
                  for k in (0, 8, 16, 24, 32, 40)}
        self.nonsquare_lookup = [1, ZETA_TO_ONE_MINUS_M_DIV_TWO]   # invsqrt.rs:51


_TABLES: Optional[_SqrtTables] = None


def sqrt_tables() -> _SqrtTables:
    global _TABLES
    if _TABLES is None:
        _TABLES = _SqrtTables()
    return _TABLES


def sqrt_ratio_zeta(num: int, den: int) -> Tuple[bool, int]:
    """invsqrt.rs:75-166."""
    T = sqrt_tables()
    if num == 0:
        return True, 0
    if den == 0:
        return False, 0
    s = pow(den, (1 << SQRT_N) - 1, Q)                       # :88-89
    t = s * s * den % Q                                      # :90
    w = pow(num * t % Q, M_MINUS_ONE_DIV_TWO, Q) * s % Q     # :91
    v = w * den % Q                                          # :93
    uv = w * num % Q                                         # :94
    x5 = uv * v % Q                                          # :97
    x4 = pow(x5, 1 << 8, Q)
    x3 = pow(x4, 1 << 8, Q)
    x2 = pow(x3, 1 << 8, Q)
    x1 = pow(x2, 1 << 8, Q)
    x0 = pow(x1, 1 << 7, Q)                                  # :101-110
    g = T.g
    q0p = T.s_lookup[x0]                                     # :113
    t = q0p
    a1 = x1 * g[32][t & 0xFF] % Q
    t += T.s_lookup[a1] << 7
    a2 = x2 * g[24][t & 0xFF] * g[32][(t >> 8) & 0xFF] % Q
    t += T.s_lookup[a2] << 15
    a3 = x3 * g[16][t & 0xFF] * g[24][(t >> 8) & 0xFF] * g[32][(t >> 16) & 0xFF] % Q
    t += T.s_lookup[a3] << 23
    a4 = (x4 * g[8][t & 0xFF] * g[16][(t >> 8) & 0xFF] * g[24][(t >> 16) & 0xFF]
          * g[32][(t >> 24) & 0xFF]) % Q
    t += T.s_lookup[a4] << 31
    a5 = (x5 * g[0][t & 0xFF] * g[8][(t >> 8) & 0xFF] * g[16][(t >> 16) & 0xFF]
          * g[24][(t >> 24) & 0xFF] * g[32][(t >> 32) & 0xFF]) % Q
    t += T.s_lookup[a5] << 39
    t = (t + 1) >> 1                                         # :155
    res = (uv * T.nonsquare_lookup[q0p & 1]
           * g[0][t & 0xFF] * g[8][(t >> 8) & 0xFF] * g[16][(t >> 16) & 0xFF]
           * g[24][(t >> 24) & 0xFF] * g[32][(t >> 32) & 0xFF]
           * g[40][(t >> 40) & 0xFF]) % Q                    # :156-163
    return (q0p & 1) == 0, res


def isqrt(x: int) -> Tuple[bool, int]:
    """``sqrt_ratio_zeta(ONE, x)`` – the only form the hot path calls
    (encoding.rs:57,102; elligator.rs:26)."""
    return sqrt_ratio_zeta(1, x)


# ---------------------------------------------------------------------------
# Group.  Element = (X, Y, Z, T) extended twisted Edwards, a=-1, d=3021.
# ---------------------------------------------------------------------------
Point = Tuple[int, int, int, int]
IDENTITY: Point = (0, 1, 1, 0)               # min_curve/element.rs:54-59
GENERATOR: Point = (B_X, B_Y, 1, B_T)        # min_curve/element.rs:62-82


def point_add(p: Point, o: Point) -> Point:
    """min_curve/element.rs:291-322 (8M + 1D, complete for a=-1, d nonsquare)."""
    x1, y1, z1, t1 = p
    x2, y2, z2, t2 = o
    a = (y1 - x1) * (y2 - x2) % Q
    b = (y1 + x1) * (y2 + x2) % Q
    c = COEFF_K * t1 % Q * t2 % Q
    d = (z1 + z1) * z2 % Q
    e, f, g, h = (b - a) % Q, (d - c) % Q, (d + c) % Q, (b + a) % Q
    return (e * f % Q, g * h % Q, f * g % Q, e * h % Q)


def point_double(p: Point) -> Point:
    """min_curve/element.rs:119-136."""
    x, y, z, _ = p
    a = x * x % Q
    b = y * y % Q
    c = 2 * z * z % Q
    d = (-a) % Q
    e = ((x + y) * (x + y) - a - b) % Q
    g = (d + b) % Q
    f = (g - c) % Q
    h = (d - b) % Q
    return (e * f % Q, g * h % Q, f * g % Q, e * h % Q)


def point_neg(p: Point) -> Point:
    """min_curve/element.rs:324-332."""
    x, y, z, t = p
    return ((-x) % Q, y, z, (-t) % Q)


def point_eq(p: Point, o: Point) -> bool:
    """min_curve/element.rs:334-340 / ark element/projective.rs:65-70."""
    return (p[0] * o[1] - o[0] * p[1]) % Q == 0


def is_identity(p: Point) -> bool:
    """min_curve/element.rs:113-117."""
    return p[0] % Q == 0


def on_curve(p: Point) -> bool:
    """ark_curve/on_curve.rs:17-38 without the [2r]P torsion check
    (that one is ``scalar_mul(p, 2*R)`` being the identity)."""
    x, y, z, t = p
    if z % Q == 0:
        return False
    lhs = (y * y + COEFF_A * x * x) % Q
    rhs = (z * z + COEFF_D * t * t) % Q
    return lhs == rhs and (t * z - x * y) % Q == 0


def scalar_mul(p: Point, k: int) -> Point:
    """min_curve/element.rs:138-153 (LSB-first double-and-add over the
    canonical bits of k)."""
    acc = IDENTITY
    ins = p
    while k:
        if k & 1:
            acc = point_add(acc, ins)
        ins = point_double(ins)
        k >>= 1
    return acc


def compress_to_field(p: Point) -> int:
    """ark_curve/encoding.rs:91-114 / min_curve/element.rs:163-183."""
    x, y, z, t = p
    a_minus_d = (COEFF_A - COEFF_D) % Q
    u1 = (x + t) * (x - t) % Q
    _, v = isqrt(u1 * a_minus_d % Q * x % Q * x % Q)
    u2 = fq_abs(v * u1 % Q)
    u3 = (u2 * z - t) % Q
    return fq_abs(a_minus_d * v % Q * u3 % Q * x % Q)


def compress(p: Point) -> bytes:
    """ark_curve/encoding.rs:116-128."""
    return fq_to_bytes(compress_to_field(p))


def decompress(enc: bytes) -> Optional[Point]:
    """ark_curve/encoding.rs:32-83 / min_curve/element.rs:248-288.
    ``None`` stands for ``Err(EncodingError::InvalidEncoding)``."""
    assert len(enc) == 32
    if enc[31] >> 5:
        return None
    s = fq_from_bytes_checked(enc)
    if s is None or is_negative(s):
        return None
    ss = s * s % Q
    u1 = (1 - ss) % Q
    u2 = (u1 * u1 - 4 * COEFF_D * ss) % Q
    was_square, v = isqrt(u2 * u1 % Q * u1 % Q)
    if not was_square:
        return None
    two_s_u1 = 2 * s * u1 % Q
    if is_negative(two_s_u1 * v % Q):
        v = (-v) % Q
    x = two_s_u1 * v % Q * v % Q * u2 % Q
    y = (1 + ss) * v % Q * u1 % Q
    return (x, y, 1, x * y % Q)


def elligator_map(r0: int) -> Point:
    """ark_curve/elligator.rs:15-62 / min_curve/element.rs:190-231."""
    A, D = COEFF_A, COEFF_D
    r = ZETA * r0 % Q * r0 % Q
    den = (D * r - (D - A)) % Q * (((D - A) * r - D) % Q) % Q
    num = (r + 1) * (A - 2 * D) % Q
    iss, isri = isqrt(num * den % Q)
    if iss:
        sgn, twiddle = 1, 1
    else:
        sgn, twiddle = Q - 1, r0
    isri = isri * twiddle % Q
    s = isri * num % Q
    t = (-sgn * isri % Q * s % Q * (r - 1) % Q * pow(A - 2 * D, 2, Q) - 1) % Q
    if is_negative(s) == iss:
        s = (-s) % Q
    E = 2 * s % Q
    F = (1 + A * s * s) % Q
    G = (1 - A * s * s) % Q
    H = t
    return (E * H % Q, F * G % Q, F * H % Q, E * G % Q)


def encode_to_curve(r0: int) -> Point:
    """ark_curve/elligator.rs:74-76."""
    return elligator_map(r0)


def hash_to_curve(r1: int, r2: int) -> Point:
    """ark_curve/elligator.rs:67-71."""
    return point_add(elligator_map(r1), elligator_map(r2))


def vartime_multiscalar_mul(scalars: Iterable[int], points: Iterable[Point]) -> Point:
    """ark_curve/element/projective.rs:99-117: zip + serial fold of s*P."""
    acc = IDENTITY
    for s, p in zip(scalars, points):
        acc = point_add(acc, scalar_mul(p, s % R))
    return acc


def to_affine(p: Point) -> Tuple[int, int]:
    zi = pow(p[2], -1, Q)
    return p[0] * zi % Q, p[1] * zi % Q


# ---------------------------------------------------------------------------
# Wire formats used at the C-ABI seam (SURVEY.md section 8 b): an Element is
# X||Y||Z||T, each 32-byte little-endian *Montgomery* (R=2^256) residue, the
# byte image of both reference backends' limbs (fq/u32/wrapper.rs:93-104).
# ---------------------------------------------------------------------------
def fq_to_mont_bytes(x: int) -> bytes:
    return (x * MONT_R % Q).to_bytes(32, "little")


def fq_from_mont_bytes(b: bytes) -> int:
    return int.from_bytes(b, "little") * MONT_R_INV_Q % Q


def point_to_wire(p: Point) -> bytes:
    return b"".join(fq_to_mont_bytes(c) for c in p)


def point_from_wire(b: bytes) -> Point:
    assert len(b) == 128
    return tuple(fq_from_mont_bytes(b[32 * i:32 * i + 32]) for i in range(4))  # type: ignore


# ---------------------------------------------------------------------------
# Deterministic inputs (SURVEY.md section 8 d / BASELINE.md section 4):
# B(tag, i) = bytes [32 i, 32 i + 32) of SHAKE-256("decaf377-b200/v1/" || tag)
# ---------------------------------------------------------------------------
def xof_blocks(tag: str, n: int, start: int = 0) -> List[bytes]:
    raw = hashlib.shake_256(("decaf377-b200/v1/" + tag).encode()).digest(32 * (start + n))
    return [raw[32 * i:32 * i + 32] for i in range(start, start + n)]


def xof_bytes(tag: str, n: int) -> bytes:
    return hashlib.shake_256(("decaf377-b200/v1/" + tag).encode()).digest(32 * n)
