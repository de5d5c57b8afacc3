"""Multi-GPU sharding of the path: independent units, zero exchange.

One process per GPU; rank g of G owns the contiguous index range [g*N/G, (g+1)*N/G).  The only
collective is the final gather of the (small) per-instance error vector -- and, on request, of the
result shards -- over torch.distributed (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np


def shard_range(n: int, rank: int, world: int):
    """Contiguous range of rank `rank` out of `world` over n units (ragged sizes allowed)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return (n * rank) // world, (n * (rank + 1)) // world


def run_sharded(n: int, compute, dist=None, gather_results: bool = False, device=None):
    """Runs compute(lo, hi) -> (out [rows][hi-lo] uint64, err [hi-lo] uint32) on this rank's shard.

    Returns (out_local, err_global, (lo, hi)); with gather_results the full [rows][n] matrix is
    returned on every rank instead of the local shard.  `dist` is torch.distributed (initialised)
    or None for a single process.
    """
    import torch
    rank = dist.get_rank() if dist is not None else 0
    world = dist.get_world_size() if dist is not None else 1
    lo, hi = shard_range(n, rank, world)
    out, err = compute(lo, hi)
    if dist is None or world == 1:
        return out, err, (lo, hi)
    sizes = [shard_range(n, r, world) for r in range(world)]
    width = max(b - a for a, b in sizes)
    dev = device if device is not None else torch.device("cpu")
    # ragged shards: pad to the widest, all_gather, cut
    pad = np.zeros(width, dtype=np.int64)
    pad[: hi - lo] = err.astype(np.int64)
    mine = torch.from_numpy(pad).to(dev)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    err_all = np.concatenate([p.cpu().numpy()[: b - a] for p, (a, b) in zip(parts, sizes)]).astype(np.uint32)
    if not gather_results:
        return out, err_all, (lo, hi)
    rows = out.shape[0]
    padm = np.zeros((rows, width), dtype=np.int64)
    padm[:, : hi - lo] = out.view(np.int64)
    minem = torch.from_numpy(padm).to(dev)
    partsm = [torch.empty_like(minem) for _ in range(world)]
    dist.all_gather(partsm, minem)
    full = np.concatenate([p.cpu().numpy()[:, : b - a] for p, (a, b) in zip(partsm, sizes)], axis=1).view(np.uint64)
    return full, err_all, (lo, hi)


class RootGather:
    """The final result gather of the path (north_star: "only a final result gather, NCCL over NVLink only for that"):
    the column shards [rows][hi-lo] of every rank into rank `dst`'s memory as one [rows][n] matrix.  Buffers are
    allocated once (padded to the widest shard: ragged sizes allowed) and reused by every call; `dist` is
    torch.distributed (NCCL with CUDA tensors, gloo with CPU tensors)."""

    def __init__(self, n: int, rows: int, dtype, device, dist, dst: int = 0):
        import torch
        self.dist, self.dst, self.n, self.rows = dist, dst, n, rows
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.sizes = [shard_range(n, r, self.world) for r in range(self.world)]
        self.width = max(b - a for a, b in self.sizes)
        self.mine = torch.zeros((rows, self.width), dtype=dtype, device=device)
        self.parts = [torch.empty((rows, self.width), dtype=dtype, device=device) for _ in range(self.world)] if self.rank == dst else None
        self.bytes_into_root = sum((b - a) for r, (a, b) in enumerate(self.sizes) if r != dst) * rows * self.mine.element_size()

    def __call__(self, local):
        """local: this rank's [rows][hi-lo] tensor.  Enqueues the copy + gather; on `dst` the shards are then in
        `self.parts` (padded); `assemble()` returns the [rows][n] matrix."""
        lo, hi = self.sizes[self.rank]
        self.mine[:, : hi - lo].copy_(local, non_blocking=True)
        self.dist.gather(self.mine, self.parts, dst=self.dst)

    def assemble(self):
        import torch
        assert self.rank == self.dst
        return torch.cat([p[:, : b - a] for p, (a, b) in zip(self.parts, self.sizes)], dim=1)
