"""Loss functions -- drop-in for light_unet/models/losses.py.

FocalTverskyLoss (the configured loss, ref :11-54) runs on libl3d kernels: one
reduction pass producing {sum p*t, sum p, sum t}, a scalar finish, and an
elementwise gradient kernel; the scalar stays on the device (no host sync).
CombinedLoss / DiceLoss are the reference's fallback / debugging options
(ref :57-113) and are not on the accelerated path: they are composed from
FocalTverskyLoss plus plain torch ops.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _native as nv


class _FocalTverskyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, alpha, beta, gamma, smooth, reduce_group):
        nv.require_cuda(pred, "FocalTverskyLoss")
        p = pred.reshape(-1)            # losses.py:40 (.view(-1); requires a contiguous prediction)
        t = target.reshape(-1)
        if p.dtype != torch.float32:
            p = p.float()
        if t.dtype != torch.float32 or t.device != p.device:
            t = t.to(device=p.device, dtype=torch.float32)
        p = p.contiguous()
        t = t.contiguous()
        st = nv.stream_ptr(p.device)
        sums = torch.zeros(3, dtype=torch.float64, device=p.device)
        with torch.cuda.device(p.device):
            nv.call("l3d_ftl_sums", nv.ptr(p), nv.ptr(t), p.numel(), nv.ptr(sums), st, algo_bytes=8 * p.numel())
        if reduce_group is not None:
            # data-parallel: the Tversky index is a ratio of batch-global sums (losses.py:44-49), so the three
            # sums are all-reduced before the ratio
            import torch.distributed as dist
            dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=reduce_group if reduce_group is not True else None)
        loss = torch.empty((), dtype=torch.float32, device=p.device)
        coef = torch.empty(2, dtype=torch.float32, device=p.device)
        with torch.cuda.device(p.device):
            nv.call("l3d_ftl_finish", nv.ptr(sums), alpha, beta, gamma, smooth, nv.ptr(loss), nv.ptr(coef), st)
        ctx.save_for_backward(t, coef)
        ctx.shape = pred.shape
        ctx.sums = sums
        return loss

    @staticmethod
    def backward(ctx, g_loss):
        t, coef = ctx.saved_tensors
        g = g_loss.reshape(1).float().contiguous()
        grad = torch.empty(t.numel(), dtype=torch.float32, device=t.device)
        with torch.cuda.device(t.device):
            nv.call("l3d_ftl_grad", nv.ptr(t), t.numel(), nv.ptr(coef), nv.ptr(g), nv.ptr(grad), nv.stream_ptr(t.device),
                    algo_bytes=8 * t.numel())
        return grad.view(ctx.shape), None, None, None, None, None, None


class FocalTverskyLoss(nn.Module):
    """Focal Tversky loss; signature and defaults of the reference (losses.py:21-28)."""

    def __init__(self, alpha=0.7, beta=0.3, gamma=0.75, smooth=1e-6):
        super().__init__()
        self.alpha = alpha
        self.beta = beta
        self.gamma = gamma
        self.smooth = smooth
        self.reduce_group = None   # set to a process group (or True for the default group) for data-parallel steps
        assert abs(alpha + beta - 1.0) < 1e-6, f"alpha + beta must equal 1.0, got {alpha + beta}"

    def forward(self, pred, target):
        return _FocalTverskyFn.apply(pred, target, float(self.alpha), float(self.beta), float(self.gamma),
                                     float(self.smooth), self.reduce_group)


class CombinedLoss(nn.Module):
    """0.8 * FocalTversky + 0.2 * BCE fallback (losses.py:57-86)."""

    def __init__(self, ftl_weight=0.8, bce_weight=0.2, alpha=0.7, beta=0.3, gamma=0.75):
        super().__init__()
        self.ftl_weight = ftl_weight
        self.bce_weight = bce_weight
        self.focal_tversky = FocalTverskyLoss(alpha=alpha, beta=beta, gamma=gamma)
        self.bce = nn.BCELoss()
        assert abs(ftl_weight + bce_weight - 1.0) < 1e-6, f"Weights must sum to 1.0, got {ftl_weight + bce_weight}"

    def forward(self, pred, target):
        ftl = self.focal_tversky(pred, target)
        bce = self.bce(pred.reshape(-1), target.reshape(-1).to(pred.dtype))
        return self.ftl_weight * ftl + self.bce_weight * bce


class DiceLoss(nn.Module):
    """Plain Dice loss kept for comparison/debugging (losses.py:88-113)."""

    def __init__(self, smooth=1e-6):
        super().__init__()
        self.smooth = smooth

    def forward(self, pred, target):
        p = pred.reshape(-1)
        t = target.reshape(-1).to(p.dtype)
        inter = (p * t).sum()
        return 1.0 - (2.0 * inter + self.smooth) / (p.sum() + t.sum() + self.smooth)


def get_loss_function(config):
    """Config dict -> loss module (losses.py:116-147); unknown names raise ValueError."""
    name = config.get("name", "FocalTverskyLoss")
    if config.get("use_combined_loss", False):
        w = config.get("combined_loss_weights", {"focal_tversky": 0.8, "bce": 0.2})
        return CombinedLoss(ftl_weight=w["focal_tversky"], bce_weight=w["bce"], alpha=config.get("alpha", 0.7),
                            beta=config.get("beta", 0.3), gamma=config.get("gamma", 0.75))
    if name == "FocalTverskyLoss":
        return FocalTverskyLoss(alpha=config.get("alpha", 0.7), beta=config.get("beta", 0.3),
                                gamma=config.get("gamma", 0.75))
    if name == "DiceLoss":
        return DiceLoss()
    raise ValueError(f"Unknown loss function: {name}")
