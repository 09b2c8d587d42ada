"""world_size-2 gloo tests (CPU) of the data-parallel training step and of the inference sharding helpers.

The CUDA kernels cannot run here, so the step is exercised with a small plain-torch module and a CPU restatement
of the batch-global Focal Tversky loss (same contract as FocalTverskyLoss.reduce_group: all-reduce the three sums,
local gradient = local part of the global gradient).  What is checked is the host logic that the GPU path shares:
parameter broadcast, one-bucket SUM all-reduce of the gradients, and that a dp-2 step on two half batches equals
the single-process step on the concatenated batch."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

from light_unet.parallel import DataParallelStep, shard_cases, shard_windows
from oracle import loss_ref


class _GlobalTverskyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, alpha, beta, gamma, smooth, group_on):
        p, t = pred.reshape(-1).double(), target.reshape(-1).double()
        sums = torch.stack([(p * t).sum(), p.sum(), t.sum()])
        if group_on:
            dist.all_reduce(sums, op=dist.ReduceOp.SUM)
        tp, sp, st = sums.tolist()
        num = tp + smooth
        den = tp + alpha * (st - tp) + beta * (sp - tp) + smooth
        ti = num / den
        k = -gamma * (1.0 - ti) ** (gamma - 1.0)
        ctx.coef = (k * (-num * beta) / den ** 2, k / den)        # dL/dp_i = coef0 + coef1 * t_i  (alpha + beta = 1)
        ctx.save_for_backward(target)
        return torch.tensor((1.0 - ti) ** gamma, dtype=torch.float32)

    @staticmethod
    def backward(ctx, g):
        (t,) = ctx.saved_tensors
        return (g * (ctx.coef[0] + ctx.coef[1] * t)).float(), None, None, None, None, None, None


class GlobalTversky(nn.Module):
    def __init__(self):
        super().__init__()
        self.reduce_group = None

    def forward(self, pred, target):
        return _GlobalTverskyFn.apply(pred, target, 0.7, 0.3, 0.75, 1e-6, self.reduce_group is not None)


def make_model(seed):
    torch.manual_seed(seed)
    return nn.Sequential(nn.Conv3d(1, 4, 3, padding=1), nn.InstanceNorm3d(4, affine=True), nn.LeakyReLU(0.01),
                         nn.Conv3d(4, 1, 1), nn.Sigmoid())


def make_batch():
    rng = np.random.default_rng(0)
    x = torch.from_numpy(rng.random((4, 1, 8, 8, 8), dtype=np.float32))
    t = torch.from_numpy((rng.random((4, 1, 8, 8, 8)) > 0.9).astype(np.float32))
    return x, t


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    model = make_model(seed=100 + rank)                         # different init per rank: broadcast must fix it
    opt = torch.optim.SGD(model.parameters(), lr=0.5)   # (Adam would turn the round-off-only gradient of the pre-norm conv bias into +-lr noise)
    stepper = DataParallelStep(model, GlobalTversky(), opt, world_size=world)
    x, t = make_batch()
    lo, hi = rank * 2, rank * 2 + 2
    losses = [float(stepper.step(x[lo:hi], t[lo:hi])) for _ in range(2)]
    flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    if rank == 0:
        torch.save({"losses": losses, "params": gathered}, out)
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_dp2_step_equals_single_process_step_on_the_concatenated_batch(tmp_path):
    out = str(tmp_path / "dp.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    # single process, whole batch, rank 0's initial weights
    model = make_model(seed=100)
    opt = torch.optim.SGD(model.parameters(), lr=0.5)   # (Adam would turn the round-off-only gradient of the pre-norm conv bias into +-lr noise)
    stepper = DataParallelStep(model, GlobalTversky(), opt, world_size=1)
    x, t = make_batch()
    ref_losses = [float(stepper.step(x, t)) for _ in range(2)]
    ref = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    assert torch.equal(got["params"][0], got["params"][1])                  # ranks stay in lock step
    assert np.allclose(got["losses"], ref_losses, atol=1e-6)                # batch-global loss on every rank
    assert (got["params"][0] - ref).abs().max().item() < 1e-5
    # the CPU loss used above is the reference's loss (oracle) when not distributed
    p = torch.rand(2, 1, 4, 4, 4)
    tt = (torch.rand(2, 1, 4, 4, 4) > 0.8).float()
    assert abs(float(GlobalTversky()(p, tt)) - float(loss_ref.focal_tversky(p, tt))) < 1e-6


def test_mean_of_local_losses_is_not_the_reference_loss():
    """Why the sums are all-reduced: the Tversky index is a ratio of batch-global sums (losses.py:44-49)."""
    x, t = make_batch()
    p = torch.sigmoid(x * 3 - 1)
    whole = float(loss_ref.focal_tversky(p, t))
    halves = 0.5 * (float(loss_ref.focal_tversky(p[:2], t[:2])) + float(loss_ref.focal_tversky(p[2:], t[2:])))
    assert abs(whole - halves) > 1e-6


def test_sharding_helpers_cover_everything_once():
    cases = [f"case{i}" for i in range(11)]
    for world in (1, 2, 4, 8):
        parts = [shard_cases(cases, r, world) for r in range(world)]
        assert sorted(sum(parts, [])) == sorted(cases)
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
        for nwin in (1, 24, 81, 325):
            spans = [shard_windows(nwin, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == nwin
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
