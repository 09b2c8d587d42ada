"""Data-parallel training step and inference sharding helpers (torch.distributed plumbing only)."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_cases(case_ids: Sequence, rank: int, world_size: int) -> List:
    """Volume-level sharding of sliding-window inference: case i goes to rank i mod world_size."""
    return [c for i, c in enumerate(case_ids) if i % world_size == rank]


def shard_windows(nwin: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Window-level sharding of ONE volume: contiguous [start, stop) range of the z->y->x window list for `rank`
    (sizes differ by at most one).  Every rank stitches from the gathered predictions, or rank 0 does."""
    base, extra = divmod(nwin, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def flatten_grads(params: Sequence[torch.nn.Parameter]) -> torch.Tensor:
    """One contiguous fp32 bucket holding every gradient (217,228 floats = 0.87 MB for the shipped model)."""
    return torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).float() for p in params])


def unflatten_into_grads(flat: torch.Tensor, params: Sequence[torch.nn.Parameter]) -> None:
    off = 0
    for p in params:
        n = p.numel()
        g = flat[off:off + n].view_as(p)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += n


class DataParallelStep:
    """One optimiser step of the reference training loop body (trainer.py:223-232): forward, loss, zero_grad,
    backward, step -- data parallel when world_size > 1.

    With `world_size > 1` the loss module must expose `reduce_group` (FocalTverskyLoss does): the Tversky sums
    are all-reduced inside the loss so every rank holds the batch-global loss and local gradients that are the
    local part of the global gradient; parameter gradients are then summed over ranks in one bucket.
    """

    def __init__(self, model: torch.nn.Module, loss_fn: torch.nn.Module, optimizer: torch.optim.Optimizer,
                 world_size: int = 1, group: Optional[dist.ProcessGroup] = None):
        self.model, self.loss_fn, self.optimizer = model, loss_fn, optimizer
        self.world_size, self.group = int(world_size), group
        self.params = [p for p in model.parameters() if p.requires_grad]
        if self.world_size > 1:
            if not dist.is_initialized():
                raise RuntimeError("DataParallelStep: torch.distributed is not initialised")
            if not hasattr(loss_fn, "reduce_group"):
                raise TypeError("DataParallelStep needs a loss with batch-global sums (FocalTverskyLoss.reduce_group)")
            loss_fn.reduce_group = group if group is not None else True
            self.broadcast_parameters()

    def broadcast_parameters(self, src: int = 0) -> None:
        """Same initial weights on every rank (rank `src`'s)."""
        flat = torch.cat([p.detach().reshape(-1) for p in self.model.parameters()])
        dist.broadcast(flat, src=src, group=self.group)
        off = 0
        with torch.no_grad():
            for p in self.model.parameters():
                p.copy_(flat[off:off + p.numel()].view_as(p))
                off += p.numel()

    def _shared_flat_grad(self) -> Optional[torch.Tensor]:
        """The engine hands out every gradient as a view of one flat fp32 buffer, in parameter order: when that is what
        .grad holds, the bucket IS that buffer (no gather / scatter copies around the all-reduce)."""
        g0 = self.params[0].grad
        if g0 is None or g0.dtype != torch.float32:
            return None
        base, off = g0.untyped_storage().data_ptr(), g0.storage_offset()
        for p in self.params:
            g = p.grad
            if (g is None or g.dtype != torch.float32 or not g.is_contiguous() or g.untyped_storage().data_ptr() != base
                    or g.storage_offset() != off):
                return None
            off += g.numel()
        total = off - g0.storage_offset()
        return torch.empty(0, dtype=torch.float32, device=g0.device).set_(g0.untyped_storage(), g0.storage_offset(), (total,), (1,))

    def reduce_gradients(self) -> None:
        shared = self._shared_flat_grad()
        if shared is not None:
            dist.all_reduce(shared, op=dist.ReduceOp.SUM, group=self.group)
            return
        flat = flatten_grads(self.params)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        unflatten_into_grads(flat, self.params)

    def step(self, images: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        images = images.float()                                    # trainer.py:225
        outputs = self.model(images)                               # trainer.py:227
        loss = self.loss_fn(outputs, labels)                       # trainer.py:228
        self.optimizer.zero_grad()                                 # trainer.py:230 (torch default: set_to_none=True)
        loss.backward()                                            # trainer.py:231
        if self.world_size > 1:
            self.reduce_gradients()
        self.optimizer.step()                                      # trainer.py:232
        return loss.detach()
