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
                 world_size: int = 1, group: Optional[dist.ProcessGroup] = None, use_graph: bool = False):
        """use_graph: capture the whole step (forward, loss, backward, gradient all-reduce, optimiser) in one CUDA graph
        per input shape and replay it -- the ~100 launches of a step then cost one host call.  Needs an optimiser built
        with capturable=True; the returned loss is a static tensor that the next step overwrites."""
        self.model, self.loss_fn, self.optimizer = model, loss_fn, optimizer
        self.world_size, self.group = int(world_size), group
        self.use_graph, self._graph, self._static, self._staged, self._copy_stream = bool(use_graph), None, None, None, None
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

    # ------------------------------------------------------------------ CUDA-graph step
    def _capture(self, images: torch.Tensor, labels: torch.Tensor, key) -> None:
        dev = images.device
        sx, st = images.float().clone(), labels.clone()
        params = list(self.model.parameters())
        p_snap = [p.detach().clone() for p in params]
        s_snap = {p: {k: v.clone() for k, v in self.optimizer.state.get(p, {}).items() if torch.is_tensor(v)} for p in params}
        rng = torch.cuda.get_rng_state(dev)
        try:        # warm-up runs on a side stream, the capture on another: the AccumulateGrad stream note does not apply
            torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
        except AttributeError:
            pass
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                              # warm-up: workspaces, optimiser state, NCCL channels
            for _ in range(2):
                self._eager_step(sx, st)
        torch.cuda.current_stream(dev).wait_stream(side)
        from .. import _native as nv
        graph = torch.cuda.CUDAGraph()
        self.optimizer.zero_grad(set_to_none=True)
        l0 = nv.launch_count()
        with torch.cuda.graph(graph):
            loss = self._eager_step(sx, st)
        self._graph_launches = nv.launch_count() - l0
        # the warm-up steps were real steps: put parameters, optimiser state and the RNG back where the caller left them
        with torch.no_grad():
            for p, q in zip(params, p_snap):
                p.copy_(q)
            for p in params:
                for k, v in self.optimizer.state.get(p, {}).items():
                    if torch.is_tensor(v):
                        if k in s_snap[p]:
                            v.copy_(s_snap[p][k])
                        else:
                            v.zero_()
        torch.cuda.set_rng_state(rng, dev)
        self._graph, self._static = graph, {"key": key, "x": sx, "t": st, "loss": loss}

    def _graph_step(self, images: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        key = (tuple(images.shape), tuple(labels.shape), labels.dtype, str(images.device))
        if self._graph is None or self._static["key"] != key:
            dev = next(self.model.parameters()).device
            self._capture(images.to(dev), labels.to(dev), key)
        self._static["x"].copy_(images, non_blocking=True)
        self._static["t"].copy_(labels, non_blocking=True)
        self._graph.replay()
        from .. import _native as nv
        nv.add_launch_count(self._graph_launches)
        return self._static["loss"]

    def prefetch(self, images: torch.Tensor, labels: torch.Tensor) -> None:
        """Start the host-to-device copy of the NEXT batch (pinned host tensors) on a side stream, under the current step;
        `step()` without arguments consumes it."""
        dev = next(self.model.parameters()).device
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(self._copy_stream):
            x = images.to(dev, non_blocking=True)
            t = labels.to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        self._staged = (x, t, ev)

    def step(self, images: Optional[torch.Tensor] = None, labels: Optional[torch.Tensor] = None) -> torch.Tensor:
        if images is None:
            if self._staged is None:
                raise RuntimeError("DataParallelStep.step(): no batch given and none prefetched")
            images, labels, ev = self._staged
            self._staged = None
            cur = torch.cuda.current_stream(images.device)
            cur.wait_event(ev)
            images.record_stream(cur)
            labels.record_stream(cur)
        if self.use_graph:
            try:
                return self._graph_step(images, labels)
            except RuntimeError as e:                              # e.g. a kernel path that synchronises (dense / grouped backward)
                if self._graph is not None:
                    raise
                import warnings
                warnings.warn(f"DataParallelStep: CUDA-graph capture failed ({e}); running eagerly")
                self.use_graph = False
        return self._eager_step(images, labels)

    def _eager_step(self, images: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        images = images.float()                                    # trainer.py:225
        outputs = self.model(images)                               # trainer.py:227
        loss = self.loss_fn(outputs, labels)                       # trainer.py:228
        self.optimizer.zero_grad()                                 # trainer.py:230 (torch default: set_to_none=True)
        loss.backward()                                            # trainer.py:231
        if self.world_size > 1:
            self.reduce_gradients()
        self.optimizer.step()                                      # trainer.py:232
        return loss.detach()
