"""Multi-GPU MSM: shard (scalar, point) pairs across ranks, one Pippenger per GPU,
combine the 128-byte partial sums.

SURVEY.md section 8e: sum_i s_i P_i is associative and commutative, so each rank
owns a contiguous slice and runs a complete MSM on it; the only exchange step is
one all-gather of 128 bytes per rank (NCCL over NVLink on GPUs, gloo in the CPU
tests), followed by world_size - 1 point additions.  Elliptic-curve addition is
not an NCCL reduction operator, hence all-gather + local sum instead of
all-reduce.  One process per GPU; ``torch.distributed`` is plumbing only.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of n items owned by `rank` (sizes differ by <= 1)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_partials(partial: torch.Tensor, group=None) -> torch.Tensor:
    """All-gather one 128-byte partial sum per rank -> [world, 128] (rank order)."""
    if partial.dtype != torch.uint8 or partial.numel() != 128:
        raise ValueError("partial must be 128 uint8")
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return partial.reshape(1, 128).clone()
    world = dist.get_world_size(group)
    out = torch.empty((world, 128), dtype=torch.uint8, device=partial.device)
    dist.all_gather_into_tensor(out, partial.reshape(1, 128).contiguous(), group=group)
    return out


def msm_sharded_async(scalars_local: torch.Tensor, points_local: torch.Tensor, point_format: int = 0,
                      group=None, inputs_ready: bool = False):
    """Throughput form of ``msm_sharded`` on the CUDA engine: nothing waits on the host and
    the engine stream never waits for the exchange.  The local Pippenger is enqueued with
    ``d377_msm_dev_async``; its partial sum becomes complete on the engine's result stream,
    the all-gather (torch's current stream) is ordered behind that, and the final sum +
    compress runs on the result stream again -- so rank-local MSM k+1 starts right behind
    the last bucket accumulation of MSM k while tail, all-gather and sum of MSM k run under
    it.  Returns (element, encoding) tensors that are complete on ``device.result_stream()``
    (``api.join()`` / ``api.sync()`` order the engine stream / the host behind them)."""
    from . import device
    partial, _ = device.msm_async(scalars_local, points_local, point_format, want_encoding=False,
                                  inputs_ready=inputs_ready)
    torch.cuda.current_stream().wait_stream(device.result_stream())
    gathered = gather_partials(partial, group)
    return device.element_sum_result(gathered)


def msm_sharded(scalars_local: torch.Tensor, points_local: torch.Tensor, point_format: int = 0,
                group=None, want_encoding: bool = True,
                local_msm: Optional[Callable] = None, local_sum: Optional[Callable] = None):
    """Every rank passes its own slice; every rank returns the same (element, encoding).

    ``local_msm`` / ``local_sum`` default to the CUDA engine; the CPU (gloo) tests
    inject the oracle there to exercise the sharding and exchange logic without a GPU.
    """
    if local_msm is None or local_sum is None:
        from . import api, device
        local_msm = local_msm or (lambda s, p, f: device.msm(s, p, f, want_encoding=False)[0])
        local_sum = local_sum or device.element_sum
        partial = local_msm(scalars_local, points_local, point_format)
        # the collective runs on torch's stream: order it after the engine stream
        torch.cuda.current_stream().wait_stream(device.engine_stream())
    else:
        partial = local_msm(scalars_local, points_local, point_format)
    gathered = gather_partials(partial, group)
    el, enc = local_sum(gathered)
    return el, (enc if want_encoding else None)
