import sys, torch
sys.path.insert(0, ".")
import decaf377_b200 as d
from decaf377_b200 import device as dev
from tools.quick_bench import timeit
d.init(0)
g = torch.Generator(device="cuda").manual_seed(1)
for logn in (18, 20, 22):
    n = 1 << logn
    r1 = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device="cuda", generator=g)
    r2 = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device="cuda", generator=g)
    for _ in range(2):
        ms = timeit(lambda: dev.hash_to_curve(r1, r2, d.OUT_ENCODING), iters=5)
        ms2 = timeit(lambda: dev.encode_to_curve(r1, d.OUT_ENCODING), iters=5)
        ms3 = timeit(lambda: dev.hash_to_curve(r1, r2, d.OUT_ELEMENT), iters=5)
        print("n=2^%d hash+enc %.3f ms %.1f Melem/s | encode+enc %.3f ms %.1f | hash element %.3f ms %.1f" % (
            logn, ms, n / ms / 1e3, ms2, n / ms2 / 1e3, ms3, n / ms3 / 1e3), flush=True)
