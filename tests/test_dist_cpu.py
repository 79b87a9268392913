"""world_size-2 gloo test of the sharded-MSM host logic (no GPU): slices, the
128-byte all-gather and the combine step, with the oracle standing in for the
per-rank MSM kernel."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from decaf377_b200.dist import gather_partials, msm_sharded, shard_range
from oracle import c_oracle as co
from oracle import decaf377_ref as o
from tests.util import canon, oracle_points, oracle_scalars, wire


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 8, 1000, 1 << 20):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        S = canon(oracle_scalars("dist_sc", n))
        P = wire(oracle_points("dist_pt", n))
        lo, hi = shard_range(n, rank, world)

        def local_msm(s, p, fmt):
            el, _ = co.msm_pippenger(s.numpy(), p.numpy(), threads=1)
            return torch.from_numpy(el.copy())

        def local_sum(g):
            acc = o.IDENTITY
            for i in range(g.shape[0]):
                acc = o.point_add(acc, o.point_from_wire(g[i].numpy().tobytes()))
            return (torch.from_numpy(np.frombuffer(o.point_to_wire(acc), np.uint8).copy()),
                    torch.from_numpy(np.frombuffer(o.compress(acc), np.uint8).copy()))

        el, enc = msm_sharded(torch.from_numpy(S[lo:hi].copy()), torch.from_numpy(P[lo:hi].copy()),
                              0, local_msm=local_msm, local_sum=local_sum)
        g = gather_partials(el)
        assert g.shape == (world, 128)
        out_q.put((rank, enc.numpy().tobytes()))
    finally:
        dist.destroy_process_group()


def test_sharded_msm_world2_gloo():
    n, world = 37, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = o.compress(o.vartime_multiscalar_mul(oracle_scalars("dist_sc", n), oracle_points("dist_pt", n)))
    assert res[0] == res[1] == want
