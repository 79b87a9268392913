#!/usr/bin/env python3
"""Profile driver: one MSM (and optionally the codec kernels) on device-resident inputs."""
import sys

import torch

sys.path.insert(0, ".")
import decaf377_b200 as d
from decaf377_b200 import device as dev

logn = int(sys.argv[1]) if len(sys.argv) > 1 else 22
what = sys.argv[2] if len(sys.argv) > 2 else "msm"
d.init(0)
n = 1 << logn
g = torch.Generator(device="cuda").manual_seed(1)
r = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device="cuda", generator=g)
sc = r.clone()
sc[:, 31] &= 0x03
el = dev.encode_to_curve(r, d.OUT_ELEMENT)
d.sync()
if what == "msm":
    for _ in range(2):
        dev.msm(sc, el)
    d.sync()
elif what == "codec2":
    enc = dev.compress(el)
    dev.encode_to_curve(r, d.OUT_ENCODING)
    dev.hash_to_curve(r, sc, d.OUT_ENCODING)
    dev.scalar_mul(enc[: 1 << 16], sc[: 1 << 16], d.PT_ENCODING, d.OUT_ENCODING)
    d.sync()
elif what == "msm1":
    dev.msm(sc, el)
    d.sync()
else:
    enc = dev.compress(el)
    dev.decompress(enc)
    dev.fixed_base_mul(sc, d.OUT_ENCODING)     # warm: builds the tables
    d.sync()
    dev.fixed_base_mul(sc, d.OUT_ENCODING)
    dev.encode_to_curve(r, d.OUT_ENCODING)
    dev.hash_to_curve(r, sc, d.OUT_ENCODING)
    d.sync()
