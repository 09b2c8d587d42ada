"""Lightweight3DUNet -- B200-native drop-in.

Public contract mirrored from the reference (light_unet/models/unet3d.py):
constructor kwargs and defaults (:155-158), attributes in_channels /
out_channels / encoder_channels (:161-163), count_parameters() (:225-229), the
93-entry state_dict (same key names and PyTorch-layout shapes, so reference
checkpoints load unchanged), forward(x fp32 [B,Cin,D,H,W]) -> probabilities
fp32 [B,Cout,D,H,W] (:204-223), autograd support and train/eval semantics
(Dropout3d active only in training; InstanceNorm has no running stats).

The sub-modules below are *parameter containers*: they create the same
torch.nn layers in the same order as the reference, so default initialisation
under a given torch.manual_seed and the state_dict are identical -- but nothing
is ever computed by those layers.  forward() hands all parameters to the kernel
plan in light_unet/engine.py (libl3d.so).  There is no eager fallback.
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence

import torch
import torch.nn as nn

from .. import _native as nv
from ..engine import UNetPlan

_DTYPES = {"f16": torch.float16, "fp16": torch.float16, "float16": torch.float16, "half": torch.float16,
           "f32": torch.float32, "fp32": torch.float32, "float32": torch.float32}


class DepthwiseSeparableConv3d(nn.Module):
    """Container for depthwise(3x3x3, groups=C) + pointwise(1x1x1) weights (ref :12-23)."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=1, bias=False):
        super().__init__()
        if kernel_size != 3 or stride != 1 or padding != 1 or bias:
            raise NotImplementedError("the native path implements the 3x3x3/stride 1/pad 1/no-bias case the U-Net uses")
        self.depthwise = nn.Conv3d(in_channels, in_channels, 3, 1, 1, groups=in_channels, bias=False)
        self.pointwise = nn.Conv3d(in_channels, out_channels, 1, bias=False)


class GroupedConv3d(nn.Module):
    """Container for a grouped 3x3x3 conv weight (ref :26-34)."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=1, groups=8, bias=False):
        super().__init__()
        if kernel_size != 3 or stride != 1 or padding != 1 or bias:
            raise NotImplementedError("the native path implements the 3x3x3/stride 1/pad 1/no-bias case the U-Net uses")
        self.conv = nn.Conv3d(in_channels, out_channels, 3, 1, 1, groups=groups, bias=False)


def _make_conv(cin, cout, dws, grouped_ok, groups):
    if dws:
        return DepthwiseSeparableConv3d(cin, cout)
    if grouped_ok:
        return GroupedConv3d(cin, cout, groups=groups)
    return nn.Conv3d(cin, cout, 3, padding=1, bias=False)


class ResidualBlock(nn.Module):
    """Parameters of conv1/norm1/conv2/norm2/shortcut (ref :37-75); registration order matches the
    reference so that seeded default init is identical."""

    def __init__(self, in_channels, out_channels, use_depthwise_separable=True, use_grouped=True, groups=8,
                 dropout_p=0.1):
        super().__init__()
        g_ok1 = use_grouped and groups > 1 and in_channels >= groups and out_channels >= groups
        g_ok2 = use_grouped and groups > 1 and out_channels >= groups
        self.conv1 = _make_conv(in_channels, out_channels, use_depthwise_separable, g_ok1, groups)
        self.norm1 = nn.InstanceNorm3d(out_channels, affine=True)
        self.conv2 = _make_conv(out_channels, out_channels, use_depthwise_separable, g_ok2, groups)
        self.norm2 = nn.InstanceNorm3d(out_channels, affine=True)
        self.dropout_p = float(dropout_p)
        if in_channels != out_channels:
            self.shortcut = nn.Sequential(nn.Conv3d(in_channels, out_channels, 1, bias=False),
                                          nn.InstanceNorm3d(out_channels, affine=True))
        else:
            self.shortcut = nn.Identity()


class DownBlock(nn.Module):
    """MaxPool3d(2) + ResidualBlock parameters (ref :96-111); the pool is fused into the previous block's writer."""

    def __init__(self, in_channels, out_channels, **kw):
        super().__init__()
        self.res_block = ResidualBlock(in_channels, out_channels, **kw)


class UpBlock(nn.Module):
    """ConvTranspose3d(C, C/2, 2, 2) + ResidualBlock parameters (ref :114-143)."""

    def __init__(self, in_channels, out_channels, **kw):
        super().__init__()
        self.up = nn.ConvTranspose3d(in_channels, in_channels // 2, kernel_size=2, stride=2)
        self.res_block = ResidualBlock(in_channels, out_channels, **kw)


class _UNetFunction(torch.autograd.Function):
    """Whole-network autograd node: forward and backward are kernel sequences in libl3d."""

    @staticmethod
    def forward(ctx, model, masks, training, x, *params):
        names = model._param_names
        P = dict(zip(names, params))
        B, C, D, H, W = x.shape
        dt = model.compute_dtype
        if C == 1:
            x_cl = x.reshape(B, D, H, W, 1).to(dt)
        else:
            x_cl = x.permute(0, 2, 3, 4, 1).contiguous().to(dt)
        with torch.cuda.device(x.device):          # libl3d launches on the CURRENT device: make it the tensors' device
            ws = model._plan.forward(P, x_cl, training, masks)
        ctx.model, ctx.ws, ctx.P = model, ws, P
        ctx.generation = ws.generation
        return ws.prob_out

    @staticmethod
    def backward(ctx, g_prob):
        ws = ctx.ws
        if ws.generation != ctx.generation or not ws.training:
            raise RuntimeError("Lightweight3DUNet: the activation workspace of this forward pass was overwritten by a "
                               "later forward of the same shape before backward() ran (or the pass ran without grad)")
        with torch.cuda.device(g_prob.device):
            grads = ctx.model._plan.backward(ctx.P, ws, g_prob.contiguous())
        return (None, None, None, None) + tuple(grads[n] for n in ctx.model._param_names)


class Lightweight3DUNet(nn.Module):
    """Lightweight 3D U-Net (16 -> 32 -> 64 -> 128) on hand-written sm_100a kernels."""

    def __init__(self, in_channels=1, out_channels=1, start_channels=16, encoder_channels=[16, 32, 64, 128],
                 use_depthwise_separable=True, use_grouped=True, groups=8, dropout_p=0.1):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.encoder_channels = encoder_channels
        e = list(encoder_channels)
        kw = dict(use_depthwise_separable=use_depthwise_separable, groups=groups, dropout_p=dropout_p)
        self.init_conv = ResidualBlock(in_channels, e[0], use_grouped=False, **kw)   # ref :168
        self.down1 = DownBlock(e[0], e[1], use_grouped=use_grouped, **kw)
        self.down2 = DownBlock(e[1], e[2], use_grouped=use_grouped, **kw)
        self.down3 = DownBlock(e[2], e[3], use_grouped=use_grouped, **kw)
        self.bottleneck = ResidualBlock(e[3], e[3], use_grouped=use_grouped, **kw)
        self.up1 = UpBlock(e[3], e[2], use_grouped=use_grouped, **kw)
        self.up2 = UpBlock(e[2], e[1], use_grouped=use_grouped, **kw)
        self.up3 = UpBlock(e[1], e[0], use_grouped=use_grouped, **kw)
        self.out_conv = nn.Conv3d(e[0], out_channels, kernel_size=1)
        self.dropout_p = float(dropout_p)
        self._plan = UNetPlan(in_channels, out_channels, e, use_depthwise_separable, use_grouped, groups)
        self._param_names: List[str] = [n for n, _ in self.named_parameters()]
        self.compute_dtype = _DTYPES[os.environ.get("L3D_DTYPE", "f16").lower()]

    # ------------------------------------------------------------------ knobs
    def set_compute_dtype(self, dtype):
        """Activation storage type in HBM: torch.float16 (default; the tensor-core operands are fp16 too) or
        torch.float32.  Accumulation, InstanceNorm statistics, parameters and the returned probabilities are always
        fp32.  bfloat16 is not offered: at the same 2 B / element its 8-bit significand misses the 1e-2 logit bar."""
        if isinstance(dtype, str):
            if dtype.lower() not in _DTYPES:
                raise ValueError(f"compute dtype must be one of {sorted(_DTYPES)} (got {dtype!r})")
            dtype = _DTYPES[dtype.lower()]
        if dtype not in (torch.float16, torch.float32):
            raise ValueError("compute dtype must be float16 or float32")
        self.compute_dtype = dtype
        return self

    def draw_dropout_masks(self, batch: int, device) -> List[Optional[torch.Tensor]]:
        """Channel-dropout keep masks for one training forward, drawn the way F.dropout3d draws them
        (ATen feature_dropout: empty(N,C,1,1,1).bernoulli_(1-p).div_(1-p)) in forward order, so a
        reference forward under the same torch seed on the same device drops the same channels
        (ref :66,84-85)."""
        p = self.dropout_p
        masks = []
        for b in self._plan.blocks:
            if p > 0:
                masks.append(torch.empty(batch, b.cout, 1, 1, 1, device=device).bernoulli_(1 - p).div_(1 - p))
            else:
                masks.append(None)
        return masks

    # ---------------------------------------------------------------- forward
    def forward(self, x, dropout_masks=None):
        if x.dim() != 5 or x.shape[1] != self.in_channels:
            raise ValueError(f"expected input [B, {self.in_channels}, D, H, W], got {tuple(x.shape)}")
        nv.require_cuda(x, "Lightweight3DUNet.forward")
        params = [p for _, p in self.named_parameters()]
        for p in params:
            nv.require_cuda(p, "Lightweight3DUNet parameters")
            if p.dtype != torch.float32 or not p.is_contiguous() or p.device != x.device:
                raise nv.NativeError(f"Lightweight3DUNet: parameters must be contiguous float32 tensors on the input's device (got {p.dtype} on "
                                     f"{p.device}); the activation storage type is set with set_compute_dtype, not model.half()")
        masks = dropout_masks
        if masks is None and self.training and self.dropout_p > 0:
            masks = self.draw_dropout_masks(x.shape[0], x.device)
        if masks is not None:
            masks = [None if m is None else m.reshape(m.shape[0], -1).float().contiguous() for m in masks]
        # activations are only kept for a backward pass when autograd will actually ask for one
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        return _UNetFunction.apply(self, masks, need_grad, x.float().contiguous(), *params)

    def count_parameters(self):
        total = sum(p.numel() for p in self.parameters())
        trainable = sum(p.numel() for p in self.parameters() if p.requires_grad)
        return {"total": total, "trainable": trainable}
