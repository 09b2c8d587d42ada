"""Window-level sharding of ONE volume's sliding-window inference over several GPUs (partition (2);
the loop being split is utils.py:86-134 of the reference).

The x-positions of the window grid (the longest axis: 13 positions for 320 voxels) are dealt to the ranks in contiguous
runs.  Rank r
  1. runs the network on the windows of its own x-positions (all z, y),
  2. receives the predictions of the few earlier x-positions whose windows reach into its voxel slab (the "seam": with
     50 % overlap one position, two at the irregular tail of the grid) from the ranks that own them -- point-to-point,
     25 windows x 442 KB per position for a 128x128x320 volume,
  3. stitches the voxel slab [x0, x1) that starts at its first own position with l3d_stitch_slab: the candidates of a
     voxel are added in the same z -> y -> x window order as on one GPU, so the slab is bit-identical to the same voxels
     of the single-GPU map given identical predictions,
  4. sends the slab to rank 0, which assembles the map and runs threshold -> connected components -> boxes (global).

Predictions are kept as [x-position][z][y] blocks so that a position is one contiguous send / receive and the received
positions simply precede the rank's own ones.  This module is the host logic only (plan + torch.distributed
point-to-point); it is exercised with gloo on CPU tensors in tests/test_window_shard_gloo.py.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist


@dataclass
class XShard:
    rank: int
    own: List[int]                      # x-position indices this rank computes (contiguous, possibly empty)
    need: List[int]                     # x-position indices whose windows cover the slab (ascending; received ones first)
    x0: int = 0                         # voxel slab [x0, x1) this rank stitches (empty when own is empty)
    x1: int = 0
    recv: Dict[int, List[int]] = field(default_factory=dict)   # source rank -> positions to receive
    send: Dict[int, List[int]] = field(default_factory=dict)   # destination rank -> positions to send

    @property
    def n_recv(self) -> int:
        return len(self.need) - len(self.own)


def plan_x_shards(xpos: Sequence[int], pw: int, W: int, world_size: int) -> List[XShard]:
    """Deal the x-positions `xpos` (ascending window starts, utils.py:63-81) to `world_size` ranks in contiguous runs
    whose sizes differ by at most one, and derive every rank's voxel slab, the positions covering it and the
    point-to-point exchange lists."""
    nx = len(xpos)
    base, extra = divmod(nx, world_size)
    shards, start = [], 0
    for r in range(world_size):
        n = base + (1 if r < extra else 0)
        shards.append(XShard(r, list(range(start, start + n)), []))
        start += n
    owner = {c: s.rank for s in shards for c in s.own}
    active = [s for s in shards if s.own]
    for i, s in enumerate(active):
        s.x0 = 0 if i == 0 else int(xpos[s.own[0]])
        s.x1 = int(W) if i == len(active) - 1 else int(xpos[active[i + 1].own[0]])
        # windows that reach into [x0, x1): start < x1 and start + pw > x0.  Later positions start at >= x1 by construction.
        s.need = [c for c in range(nx) if xpos[c] < s.x1 and xpos[c] + pw > s.x0]
        assert all(c in s.need for c in s.own) and s.need[-len(s.own):] == s.own
        for c in s.need:
            if owner[c] != s.rank:
                s.recv.setdefault(owner[c], []).append(c)
                shards[owner[c]].send.setdefault(s.rank, []).append(c)
    return shards


def exchange_seams(local: torch.Tensor, shard: XShard, group: Optional[dist.ProcessGroup] = None) -> None:
    """`local`: [len(shard.need), block...] predictions in x-position-major layout; the rows of the rank's own positions
    (the last len(own) rows) are filled.  Sends the rows other ranks need and receives this rank's seam rows in place."""
    ops = []
    first_own = shard.n_recv
    for dst, cols in sorted(shard.send.items()):
        for c in cols:
            ops.append(dist.P2POp(dist.isend, local[first_own + shard.own.index(c)], dst, group))
    for src, cols in sorted(shard.recv.items()):
        for c in cols:
            ops.append(dist.P2POp(dist.irecv, local[shard.need.index(c)], src, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


def gather_slabs(full: Optional[torch.Tensor], slab: Optional[torch.Tensor], shards: Sequence[XShard], rank: int,
                 group: Optional[dist.ProcessGroup] = None) -> None:
    """Rank 0 assembles the [D, H, W] map: it stitched its own slab straight into `full`; every other active rank sends
    its packed [D, H, x1 - x0] slab."""
    if rank == 0:
        bufs, ops = [], []
        for s in shards:
            if s.rank != 0 and s.own:
                b = torch.empty(full.shape[0], full.shape[1], s.x1 - s.x0, dtype=full.dtype, device=full.device)
                bufs.append((s, b))
                ops.append(dist.P2POp(dist.irecv, b, s.rank, group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        for s, b in bufs:
            full[:, :, s.x0:s.x1].copy_(b)
    elif shards[rank].own:
        for req in dist.batch_isend_irecv([dist.P2POp(dist.isend, slab, 0, group)]):
            req.wait()
