#!/usr/bin/env python3
"""Latency of SMALL MSMs (developer tool): the Pippenger call against one scalar
multiplication per pair + Sum<Element> (what the reference's fold does), host API and
device API.  usage: python tools/small_msm_probe.py"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import decaf377_b200 as d
from decaf377_b200 import device as dev

d.init(0)
g = torch.Generator(device="cuda").manual_seed(3)
for n in (1, 3, 16, 64, 256, 1024, 4096, 16384, 65536, 262144):
    r = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device="cuda", generator=g)
    sc = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device="cuda", generator=g)
    sc[:, 31] &= 0x03
    el = dev.encode_to_curve(r, d.OUT_ELEMENT)
    d.sync()
    h_sc, h_el = sc.cpu().numpy(), el.cpu().numpy()

    def t(fn, iters=20):
        for _ in range(3):
            fn()
        d.sync()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(iters):
            out = fn()
        d.sync()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / iters * 1e3, out

    ms_p, a = t(lambda: dev.msm(sc, el)[1].cpu())
    ms_f, b = t(lambda: dev.element_sum(dev.scalar_mul(el, sc))[1].cpu())
    ms_h, c = t(lambda: d.vartime_multiscalar_mul(h_sc, h_el)[1])
    info = d.msm_stage_info()
    assert bytes(a.numpy().tobytes()) == bytes(b.numpy().tobytes()) == c.tobytes(), n
    print("n=%7d  pippenger(dev) %.3f ms   mul+sum(dev) %.3f ms   pippenger(host) %.3f ms   c=%s W=%s"
          % (n, ms_p, ms_f, ms_h, info.get("c"), info.get("W")), flush=True)
