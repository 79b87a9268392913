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
