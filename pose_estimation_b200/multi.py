"""Multi-GPU host logic of the batched multi-hypothesis refinement.

Hypotheses are independent, so the batch shards across ranks (one process per GPU, contiguous
blocks); every rank holds a replica of the scene grid.  The only exchange is one all_gather of
the 96-byte peb_icp_result records, after which every rank can pick the best pose locally.
This module holds no arithmetic of the path: shard bookkeeping, the collective (NCCL on GPUs,
gloo in the CPU tests) and the unpacking of the gathered records.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

RESULT_DTYPE = np.dtype([("T", "<f4", (16,)), ("fitness", "<f8"), ("last_mse", "<f8"), ("iterations", "<i4"),
                         ("converged", "<i4"), ("state", "<i4"), ("n_correspondences", "<i4")])
RECORD_BYTES = RESULT_DTYPE.itemsize  # == sizeof(peb_icp_result) == 96


def shard_size(n_items: int, world: int) -> int:
    """Records every rank contributes to the all_gather (the last shard is padded)."""
    return (n_items + world - 1) // world


def shard_range(n_items: int, world: int, rank: int) -> tuple[int, int]:
    """[lo, hi) of the hypotheses rank `rank` refines: contiguous blocks, in rank order."""
    per = shard_size(n_items, world)
    return min(rank * per, n_items), min((rank + 1) * per, n_items)


def gather_results(local, n_items: int, world: int, rank: int, out=None, pad=None):
    """all_gather of this rank's result records (a uint8 tensor of (hi - lo) * 96 bytes, on the
    device the process group communicates on).  Returns a uint8 tensor of world * per * 96 bytes
    in rank order; with world == 1 the input itself."""
    import torch
    import torch.distributed as dist

    if world == 1:
        return local
    per = shard_size(n_items, world)
    lo, hi = shard_range(n_items, world, rank)
    if pad is None:
        pad = torch.zeros(per * RECORD_BYTES, dtype=torch.uint8, device=local.device)
    if out is None:
        out = torch.empty(world * per * RECORD_BYTES, dtype=torch.uint8, device=local.device)
    pad[: (hi - lo) * RECORD_BYTES].copy_(local[: (hi - lo) * RECORD_BYTES])
    dist.all_gather_into_tensor(out, pad)
    return out


def unpack_results(gathered_bytes: bytes | np.ndarray, n_items: int, world: int) -> np.ndarray:
    """Gathered buffer -> structured array of the n_items records in hypothesis order."""
    per = shard_size(n_items, world) if world > 1 else n_items
    recs = np.frombuffer(bytes(gathered_bytes), dtype=RESULT_DTYPE)
    out = np.empty(n_items, RESULT_DTYPE)
    for r in range(world):
        lo, hi = shard_range(n_items, world, r) if world > 1 else (0, n_items)
        out[lo:hi] = recs[r * per: r * per + (hi - lo)]
    return out


def best_hypothesis(records: np.ndarray) -> int:
    """Index of the converged hypothesis with the lowest fitness score (-1 if none converged)."""
    ok = records["converged"] != 0
    if not ok.any():
        return -1
    fit = np.where(ok, records["fitness"], np.inf)
    return int(np.argmin(fit))
