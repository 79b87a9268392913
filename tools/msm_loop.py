#!/usr/bin/env python3
"""Back-to-back MSM throughput (developer tool): blocking d377_msm_dev against
d377_msm_dev_async with the tail overlap on / off, every result checked against the
blocking call.  usage: python tools/msm_loop.py [logn ...]"""
import sys

import torch

sys.path.insert(0, ".")
import decaf377_b200 as d
from decaf377_b200 import device as dev


def timed(fn, iters):
    st = dev.engine_stream()
    for _ in range(3):
        fn()
    d.sync()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        t0.record()
    for _ in range(iters):
        fn()
    d.join()
    with torch.cuda.stream(st):
        t1.record()
    d.sync()
    t1.synchronize()
    return t0.elapsed_time(t1) / iters


def main():
    d.init(0)
    args = [a for a in sys.argv[1:] if not a.startswith("-")]
    fast = "--fast" in sys.argv          # only the overlapped + prefetching loop
    cwin = [int(a[3:]) for a in sys.argv[1:] if a.startswith("-c=")]
    if cwin:
        d.msm_set_window(cwin[0])
    logns = [int(a) for a in args] or [20, 22, 24]
    g = torch.Generator(device="cuda").manual_seed(7)
    for logn in logns:
        n = 1 << logn
        r = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device="cuda", generator=g)
        sc = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device="cuda", generator=g)
        sc[:, 31] &= 0x03
        el = dev.encode_to_curve(r, d.OUT_ELEMENT)
        d.sync()
        iters = 30 if logn <= 22 else 10
        want = dev.msm(sc, el)[1].cpu()
        ms_sync = timed(lambda: dev.msm(sc, el), iters) if not fast else float("nan")
        outs = []
        for overlap, ready in (((1, True),) if fast else ((0, False), (1, False), (1, True))):
            d.msm_set_tail_overlap(overlap)
            oe = torch.empty((iters + 3, 128), dtype=torch.uint8, device="cuda")
            oc = torch.empty((iters + 3, 32), dtype=torch.uint8, device="cuda")
            k = [0]

            def step():
                dev.msm_async(sc, el, out_element=oe[k[0] % (iters + 3)], out_encoding=oc[k[0] % (iters + 3)],
                              inputs_ready=ready)
                k[0] += 1

            ms = timed(step, iters)
            ok = bool((oc.cpu() == want).all())
            outs.append(("%d%s" % (overlap, "+prefetch" if ready else ""), ms, ok))
        d.msm_set_tail_overlap(1)
        line = "msm 2^%d: blocking %.3f ms (%.1f Mpts/s)" % (logn, ms_sync, n / ms_sync / 1e3)
        for overlap, ms, ok in outs:
            line += " | async overlap=%s %.3f ms (%.1f Mpts/s) %s" % (overlap, ms, n / ms / 1e3, "ok" if ok else "MISMATCH")
        print(line, flush=True)
        print("   stages:", {k2: round(v, 3) for k2, v in d.msm_stage_info()["ms"].items()}, flush=True)


if __name__ == "__main__":
    main()
