"""Multi-GPU partitioning of the hot path (no counterpart in the reference, which solves problems one after another on
the CPU: benchmarking.py:73-91).

* Problems are independent -> `shard_range` gives each rank a contiguous slice of the problem batch; no collective.
* Single-problem large-sample mode -> every rank holds a slice of the S Monte-Carlo samples, all ranks use the same
  Fourier basis (omega, tau are keyed by problem/latent only) and their own (w, eps) slices; the packed gradient and
  the ELBO are summed with ONE all-reduce per iteration (`allreduce_packed`), after which every rank applies the
  identical Adam update.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) of `n` items for `rank`; the first n % world ranks get one extra item."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of size {world}")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def packed_layout(Bp: int, M: int, D: int) -> Dict[str, Tuple[int, Tuple[int, ...]]]:
    """Offsets / shapes of (d_q_mu, d_q_sqrt, d_lengthscales, d_variances, elbo) inside one flat float64 buffer."""
    shapes = [("d_q_mu", (Bp, M, D)), ("d_q_sqrt", (Bp, D, M, M)), ("d_lengthscales", (Bp, D)), ("d_variances", (Bp, D)),
              ("elbo", (Bp,))]
    out, off = {}, 0
    for name, shp in shapes:
        n = 1
        for v in shp:
            n *= v
        out[name] = (off, shp)
        off += n
    out["_total"] = (off, ())
    return out


def packed_views(flat: torch.Tensor, Bp: int, M: int, D: int) -> Dict[str, torch.Tensor]:
    lay = packed_layout(Bp, M, D)
    if flat.numel() != lay["_total"][0]:
        raise ValueError("flat buffer has the wrong size")
    views = {}
    for name, (off, shp) in lay.items():
        if name.startswith("_"):
            continue
        n = 1
        for v in shp:
            n *= v
        views[name] = flat[off:off + n].view(*shp)
    return views


def allreduce_packed(flat: torch.Tensor, group=None) -> torch.Tensor:
    """SUM over the ranks of the packed (gradient || ELBO) buffer: NCCL over NVLink on GPUs, gloo on CPU tensors."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


def partition_problems(queries: List, rank: int, world: int) -> List:
    lo, hi = shard_range(len(queries), rank, world)
    return list(queries[lo:hi])
