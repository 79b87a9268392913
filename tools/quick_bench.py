#!/usr/bin/env python3
"""Quick device-resident timings of every hot-path kernel (developer tool)."""
import sys
import time

import torch

sys.path.insert(0, ".")
import decaf377_b200 as d
from decaf377_b200 import device as dev


def timeit(fn, iters=3):
    st = dev.engine_stream()
    fn()
    d.sync()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        t0.record()
        for _ in range(iters):
            fn()
        t1.record()
    t1.synchronize()
    return t0.elapsed_time(t1) / iters


def main():
    d.init(0)
    print("imad peak G/s:", d.imad_peak())
    g = torch.Generator(device="cuda").manual_seed(1)
    for logn in (16, 20, 22):
        n = 1 << logn
        r = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device="cuda", generator=g)
        sc = r.clone()
        sc[:, 31] &= 0x03
        ms = timeit(lambda: dev.encode_to_curve(r, d.OUT_ENCODING))
        print(f"encode_to_curve+compress n=2^{logn}: {ms:.3f} ms  {n/ms/1e3:.1f} Melem/s")
        el = dev.encode_to_curve(r, d.OUT_ELEMENT)
        ms = timeit(lambda: dev.encode_to_curve(r, d.OUT_ELEMENT))
        print(f"encode_to_curve n=2^{logn}: {ms:.3f} ms  {n/ms/1e3:.1f} Melem/s")
        enc = dev.compress(el)
        ms = timeit(lambda: dev.compress(el))
        print(f"compress n=2^{logn}: {ms:.3f} ms  {n/ms/1e3:.1f} Melem/s")
        ms = timeit(lambda: dev.decompress(enc))
        print(f"decompress n=2^{logn}: {ms:.3f} ms  {n/ms/1e3:.1f} Melem/s")
        ms = timeit(lambda: dev.fixed_base_mul(sc, d.OUT_ENCODING))
        print(f"fixed_base+compress n=2^{logn}: {ms:.3f} ms  {n/ms/1e3:.1f} Melem/s")
        ms = timeit(lambda: dev.fixed_base_mul(sc, d.OUT_ELEMENT))
        print(f"fixed_base n=2^{logn}: {ms:.3f} ms  {n/ms/1e3:.1f} Melem/s")
        if logn <= 20:
            ms = timeit(lambda: dev.scalar_mul(enc, sc, d.PT_ENCODING, d.OUT_ENCODING), iters=1)
            print(f"decompress->mul->compress n=2^{logn}: {ms:.3f} ms  {n/ms/1e3:.3f} Melem/s")
        ms = timeit(lambda: dev.msm(sc, el), iters=2)
        print(f"msm n=2^{logn}: {ms:.3f} ms  {n/ms/1e3:.1f} Mpoints/s")
        if logn == 20:
            for c in (12, 13, 14, 15, 16, 17, 18):
                d.msm_set_window(c)
                ms = timeit(lambda: dev.msm(sc, el), iters=2)
                print(f"  msm c={c}: {ms:.3f} ms  {n/ms/1e3:.1f} Mpoints/s")
            d.msm_set_window(0)
    print("launches:", d.launch_count())


if __name__ == "__main__":
    main()
