"""Full BASELINE.json sizes on the GPU, checked through size-independent properties
(round trips, known-answer constructions, linearity, a checksum of checksums) plus an
oracle comparison on a strided sample.  Inputs are generated on the device."""
import hashlib

import numpy as np
import pytest
import torch

from oracle import c_oracle as co
from oracle import decaf377_ref as o

pytestmark = pytest.mark.gpu

R = o.R


def _rand(n, seed, mask_top=False):
    g = torch.Generator(device="cuda").manual_seed(seed)
    t = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device="cuda", generator=g)
    if mask_top:
        t[:, 31] &= 0x03
    return t


def _ints(a):
    a = np.ascontiguousarray(a)
    return [int.from_bytes(a[i].tobytes(), "little") for i in range(a.shape[0])]


def test_config2_encode_2p22_roundtrip_and_sample(engine):
    from decaf377_b200 import device as dev
    n = 1 << 22
    r = _rand(n, 1)
    enc = dev.encode_to_curve(r, engine.OUT_ENCODING)
    el = dev.encode_to_curve(r, engine.OUT_ELEMENT)
    # compress(encode_to_curve) computed two ways, and decompress o compress = id, all 2^22
    assert torch.equal(dev.compress(el), enc)
    back, ok = dev.decompress(enc)
    engine.sync()
    assert bool(ok.all())
    assert torch.equal(dev.compress(back), enc)
    # the AffinePoint and X||Y||Z layouts of the same 2^22 points give the same encodings
    # (ark_curve/serialize.rs:8-46), and the normalised Elements are those affine points
    xy, ok2 = dev.decompress_fmt(enc, engine.PT_AFFINE)
    assert bool(ok2.all()) and torch.equal(xy, back[:, :64])
    assert torch.equal(dev.compress_fmt(xy, engine.PT_AFFINE), enc)
    assert torch.equal(dev.compress_fmt(el[:, :96].contiguous(), engine.PT_XYZ), enc)
    assert torch.equal(dev.compress_fmt(dev.normalize(el), engine.PT_AFFINE), enc)
    engine.sync()
    # strided sample against the C oracle
    idx = torch.arange(0, n, n // 4096, device="cuda")
    want = co.encode_to_curve(r[idx].cpu().numpy(), out_enc=True, threads=8)
    assert np.array_equal(enc[idx].cpu().numpy(), want)
    # checksum of checksums is reproducible run to run
    h1 = hashlib.sha256(enc.cpu().numpy().tobytes()).hexdigest()
    h2 = hashlib.sha256(dev.encode_to_curve(r, engine.OUT_ENCODING).cpu().numpy().tobytes()).hexdigest()
    assert h1 == h2


def test_config3_fixed_base_2p24_linearity_and_sample(engine):
    from decaf377_b200 import device as dev
    n = 1 << 24
    a = _rand(n, 2, mask_top=True)
    encs = dev.fixed_base_mul(a, engine.OUT_ENCODING)
    els = dev.fixed_base_mul(a, engine.OUT_ELEMENT)
    engine.sync()
    # sum_i a_i G == (sum_i a_i) G : one group element ties all 2^24 outputs together
    _, enc_sum = dev.element_sum(els)
    engine.sync()
    a_np = a.cpu().numpy()
    total = int(np.zeros(1)[0])
    limbs = a_np.view("<u8").reshape(n, 4).astype(object)
    total = sum(int(limbs[:, k].sum()) << (64 * k) for k in range(4)) % R
    assert enc_sum.cpu().numpy().tobytes() == o.compress(o.scalar_mul(o.GENERATOR, total))
    # sample against the oracle (no tables there: plain double-and-add)
    idx = torch.arange(0, n, n // 64, device="cuda")
    want = co.fixed_base(a_np[idx.cpu().numpy()], out_enc=True, threads=8)
    assert np.array_equal(encs[idx].cpu().numpy(), want)
    # the fused compress agrees with compress of the element output everywhere
    assert torch.equal(dev.compress(els[: 1 << 20]), encs[: 1 << 20])


def _dot_mod_r(a, s):
    """sum_i a_i s_i mod r for two (n, 32) uint8 device tensors of little-endian integers:
    16-bit limbs in int64 (a limb product is < 2^32, a sum of 2^26 of them < 2^58)."""
    a16 = a.view(torch.int16).to(torch.int64) & 0xFFFF
    s16 = s.view(torch.int16).to(torch.int64) & 0xFFFF
    k = 0
    for i in range(16):
        for j in range(16):
            k += int((a16[:, i] * s16[:, j]).sum().item()) << (16 * (i + j))
    return k % R


@pytest.mark.parametrize("logn", [20, 24, 26])
def test_config4_5_msm_known_answer(engine, logn):
    """P_i = a_i G  =>  sum s_i P_i = (sum s_i a_i mod r) G   at 2^20, 2^24 and
    2^26 (the largest single-GPU slice of config 5)."""
    from decaf377_b200 import device as dev
    n = 1 << logn
    a = _rand(n, 3, mask_top=True)
    s = _rand(n, 4, mask_top=True)
    P = dev.fixed_base_mul(a, engine.OUT_ELEMENT)
    _, enc = dev.msm(s, P, engine.PT_ELEMENT)
    engine.sync()
    k = _dot_mod_r(a, s)
    want = o.compress(o.scalar_mul(o.GENERATOR, k))
    assert enc.cpu().numpy().tobytes() == want
    # same points as encodings and as affine pairs must give the same answer
    if logn == 20:
        Penc = dev.compress(P)
        _, enc2 = dev.msm(s, Penc, engine.PT_ENCODING)
        engine.sync()
        assert enc2.cpu().numpy().tobytes() == want
    # sharding property (what the multi-GPU path relies on): MSM(first half) + MSM(second half)
    h = n // 2
    e1, _ = dev.msm(s[:h], P[:h], engine.PT_ELEMENT, want_encoding=False)
    e2, _ = dev.msm(s[h:], P[h:], engine.PT_ELEMENT, want_encoding=False)
    _, enc3 = dev.element_sum(torch.stack([e1, e2]))
    engine.sync()
    assert enc3.cpu().numpy().tobytes() == want


def test_config1_pipeline_2p16_vs_oracle(engine):
    from decaf377_b200 import device as dev
    n = 1 << 16
    r = _rand(n, 5)
    s = _rand(n, 6, mask_top=True)
    enc = dev.encode_to_curve(r, engine.OUT_ENCODING)
    out = dev.scalar_mul(enc, s, engine.PT_ENCODING, engine.OUT_ENCODING)
    engine.sync()
    # homomorphism over the whole batch: sum_i s_i P_i via the MSM kernel == sum of the outputs
    els, ok = dev.decompress(out)
    engine.sync()
    assert bool(ok.all())
    _, lhs = dev.element_sum(els)
    Pel, _ = dev.decompress(enc)
    _, rhs = dev.msm(s, Pel, engine.PT_ELEMENT)
    engine.sync()
    assert torch.equal(lhs, rhs)
    # oracle on a sample of 256
    idx = np.arange(0, n, n // 256)
    want, okc = co.pipeline(enc.cpu().numpy()[idx], s.cpu().numpy()[idx], threads=8)
    assert okc.all() and np.array_equal(out.cpu().numpy()[idx], want)
