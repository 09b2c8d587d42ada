"""Functional torch-CPU-fp32 restatement of the reference 3D U-Net.
(Test infrastructure -- see oracle/__init__.py.  Pinned by tests/golden/.)

Follows /root/reference/light_unet/models/unet3d.py:
  DepthwiseSeparableConv3d :12-23, GroupedConv3d :26-34, ResidualBlock :37-93,
  DownBlock :96-111, UpBlock :114-143, Lightweight3DUNet :146-229.

The network is written as plain functions over a ``state_dict`` so that it can
serve both as the parity oracle (exact reference semantics, autograd for the
gradients) and as the timed CPU baseline (same ATen CPU kernels the reference's
``nn.Module`` dispatches to).
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

LEAKY_SLOPE = 0.01  # unet3d.py:52,63
IN_EPS = 1e-5       # nn.InstanceNorm3d default eps (unet3d.py:51,62,72)


@dataclass
class UNetCfg:
    """Constructor arguments of Lightweight3DUNet (unet3d.py:155-158)."""
    in_channels: int = 1
    out_channels: int = 1
    encoder_channels: Sequence[int] = (16, 32, 64, 128)
    use_depthwise_separable: bool = True
    use_grouped: bool = True
    groups: int = 8
    dropout_p: float = 0.1


def conv_kind(cfg: UNetCfg, cin: int, cout: int, which: int, allow_grouped: bool) -> str:
    """Branch selection of ResidualBlock.__init__ (unet3d.py:44-49 for conv1,
    :55-60 for conv2).  ``allow_grouped`` is False for init_conv (:168)."""
    if cfg.use_depthwise_separable:
        return "dws"
    g = cfg.groups
    if which == 1:
        ok = allow_grouped and g > 1 and cin >= g and cout >= g
    else:
        ok = allow_grouped and g > 1 and cout >= g
    return "grouped" if ok else "dense"


def block_specs(cfg: UNetCfg):
    """(name, prefix, cin, cout, allow_grouped) for the 8 residual blocks in
    forward order (unet3d.py:165-198)."""
    e = list(cfg.encoder_channels)
    ug = cfg.use_grouped
    return [
        ("init_conv", "init_conv", cfg.in_channels, e[0], False),
        ("down1", "down1.res_block", e[0], e[1], ug),
        ("down2", "down2.res_block", e[1], e[2], ug),
        ("down3", "down3.res_block", e[2], e[3], ug),
        ("bottleneck", "bottleneck", e[3], e[3], ug),
        ("up1", "up1.res_block", e[3], e[2], ug),
        ("up2", "up2.res_block", e[2], e[1], ug),
        ("up3", "up3.res_block", e[1], e[0], ug),
    ]


def param_shapes(cfg: UNetCfg) -> "OrderedDict[str, tuple]":
    """state_dict keys and shapes in the reference's registration order."""
    shapes: "OrderedDict[str, tuple]" = OrderedDict()

    def conv(prefix, cin, cout, kind):
        if kind == "dws":
            shapes[f"{prefix}.depthwise.weight"] = (cin, 1, 3, 3, 3)
            shapes[f"{prefix}.pointwise.weight"] = (cout, cin, 1, 1, 1)
        elif kind == "grouped":
            shapes[f"{prefix}.conv.weight"] = (cout, cin // cfg.groups, 3, 3, 3)
        else:
            shapes[f"{prefix}.weight"] = (cout, cin, 3, 3, 3)

    def res_block(prefix, cin, cout, allow_grouped):
        conv(f"{prefix}.conv1", cin, cout, conv_kind(cfg, cin, cout, 1, allow_grouped))
        shapes[f"{prefix}.norm1.weight"] = (cout,)
        shapes[f"{prefix}.norm1.bias"] = (cout,)
        conv(f"{prefix}.conv2", cout, cout, conv_kind(cfg, cout, cout, 2, allow_grouped))
        shapes[f"{prefix}.norm2.weight"] = (cout,)
        shapes[f"{prefix}.norm2.bias"] = (cout,)
        if cin != cout:
            shapes[f"{prefix}.shortcut.0.weight"] = (cout, cin, 1, 1, 1)
            shapes[f"{prefix}.shortcut.1.weight"] = (cout,)
            shapes[f"{prefix}.shortcut.1.bias"] = (cout,)

    for name, prefix, cin, cout, ag in block_specs(cfg):
        if name.startswith("up"):
            shapes[f"{name}.up.weight"] = (cin, cin // 2, 2, 2, 2)
            shapes[f"{name}.up.bias"] = (cin // 2,)
        res_block(prefix, cin, cout, ag)
    e0 = cfg.encoder_channels[0]
    shapes["out_conv.weight"] = (cfg.out_channels, e0, 1, 1, 1)
    shapes["out_conv.bias"] = (cfg.out_channels,)
    return shapes


def draw_dropout_masks(cfg: UNetCfg, batch: int, device="cpu") -> List[Optional[torch.Tensor]]:
    """One [B, Cout, 1, 1, 1] keep-mask (already scaled by 1/(1-p)) per residual
    block, drawn in forward order from torch's global generator exactly the way
    ``F.dropout3d`` draws it (ATen feature_dropout: ``empty(N,C,1,1,1)
    .bernoulli_(1-p).div_(1-p)``), so that a reference ``model.train()`` forward
    under the same ``torch.manual_seed`` sees the same masks (unet3d.py:66,84-85).
    """
    p = cfg.dropout_p
    masks = []
    for _, _, _, cout, _ in block_specs(cfg):
        if p > 0:
            m = torch.empty(batch, cout, 1, 1, 1, device=device).bernoulli_(1 - p).div_(1 - p)
        else:
            m = None
        masks.append(m)
    return masks


def _conv3(sd, prefix, x, kind, groups):
    if kind == "dws":                       # unet3d.py:20-23
        c = x.shape[1]
        x = F.conv3d(x, sd[f"{prefix}.depthwise.weight"], None, 1, 1, 1, c)
        return F.conv3d(x, sd[f"{prefix}.pointwise.weight"])
    if kind == "grouped":                   # unet3d.py:33-34
        return F.conv3d(x, sd[f"{prefix}.conv.weight"], None, 1, 1, 1, groups)
    return F.conv3d(x, sd[f"{prefix}.weight"], None, 1, 1)


def _inorm(sd, prefix, x):
    return F.instance_norm(x, None, None, sd[f"{prefix}.weight"], sd[f"{prefix}.bias"],
                           use_input_stats=True, eps=IN_EPS)


def f16_storage(x: torch.Tensor) -> torch.Tensor:
    """Round to fp16 with a straight-through gradient: models a tensor that the CUDA path *stores* in fp16
    between kernels (all arithmetic stays fp32).  Used only to separate "kernel bug" from "precision policy"
    when checking the 16-bit storage mode: the CUDA backward is the exact gradient of this rounded forward."""
    return x + (x.detach().to(torch.float16).to(torch.float32) - x.detach())


def residual_block(sd, cfg, prefix, x, cin, cout, allow_grouped, mask, taps=None, quant=None):
    """ResidualBlock.forward (unet3d.py:77-93).  ``quant`` (None = the reference's plain fp32) is applied to
    the tensors the CUDA path materialises in HBM: both raw conv outputs, the raw shortcut and the block output."""
    q = quant if quant is not None else (lambda v: v)
    if cin != cout:
        r = _inorm(sd, f"{prefix}.shortcut.1", q(F.conv3d(x, sd[f"{prefix}.shortcut.0.weight"])))
    else:
        r = x
    t1 = q(_conv3(sd, f"{prefix}.conv1", x, conv_kind(cfg, cin, cout, 1, allow_grouped), cfg.groups))
    a = F.leaky_relu(_inorm(sd, f"{prefix}.norm1", t1), LEAKY_SLOPE)
    if mask is not None:
        a = a * mask
    t2 = q(_conv3(sd, f"{prefix}.conv2", a, conv_kind(cfg, cout, cout, 2, allow_grouped), cfg.groups))
    out = q(F.leaky_relu(_inorm(sd, f"{prefix}.norm2", t2) + r, LEAKY_SLOPE))
    if taps is not None:
        taps[prefix + ".t1"] = t1
        taps[prefix + ".t2"] = t2
        taps[prefix + ".out"] = out
    return out


def up_merge(sd, name, x, skip, quant=None):
    """UpBlock.forward up to the concat (unet3d.py:126-141): transposed conv,
    centre pad to the skip's size, concat [upsampled, skip]."""
    x = F.conv_transpose3d(x, sd[f"{name}.up.weight"], sd[f"{name}.up.bias"], stride=2)
    if quant is not None:
        x = quant(x)
    if x.shape != skip.shape:
        dd = skip.size(2) - x.size(2)
        dh = skip.size(3) - x.size(3)
        dw = skip.size(4) - x.size(4)
        x = F.pad(x, [dw // 2, dw - dw // 2, dh // 2, dh - dh // 2, dd // 2, dd - dd // 2])
    return torch.cat([x, skip], dim=1)


def forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, cfg: UNetCfg,
            masks: Optional[List[Optional[torch.Tensor]]] = None,
            taps: Optional[dict] = None, return_logits: bool = False, quant=None) -> torch.Tensor:
    """Lightweight3DUNet.forward (unet3d.py:204-223).  Returns probabilities
    (sigmoid already applied, :220-221) unless ``return_logits``."""
    specs = block_specs(cfg)
    if masks is None:
        masks = [None] * len(specs)
    feats = []
    h = x if quant is None else quant(x)
    for i, (name, prefix, cin, cout, ag) in enumerate(specs):
        if name.startswith("down"):
            h = F.max_pool3d(h, 2, 2)                       # unet3d.py:109
        elif name.startswith("up"):
            h = up_merge(sd, name, h, feats[3 - int(name[-1])], quant)  # up1<-x3, up2<-x2, up3<-x1
        h = residual_block(sd, cfg, prefix, h, cin, cout, ag, masks[i], taps, quant)
        if name in ("init_conv", "down1", "down2"):
            feats.append(h)
    logits = F.conv3d(h, sd["out_conv.weight"], sd["out_conv.bias"])
    if taps is not None:
        taps["logits"] = logits
    return logits if return_logits else torch.sigmoid(logits)


def to_torch(sd_np) -> "OrderedDict[str, torch.Tensor]":
    return OrderedDict((k, torch.from_numpy(v.copy())) for k, v in sd_np.items())


def count_parameters(cfg: UNetCfg) -> int:
    n = 0
    for s in param_shapes(cfg).values():
        k = 1
        for d in s:
            k *= d
        n += k
    return n


def forward_flops(cfg: UNetCfg, size) -> dict:
    """Multiply-accumulate counts per sample for one forward pass, split by
    layer type (used for the roofline arithmetic; reproduces SURVEY.md
    section 8(a)/(d) figures: 0.701 GMAC for the 217K model at 48^3)."""
    if isinstance(size, int):
        size = (size, size, size)
    dims = [tuple(size)]
    for _ in range(3):
        dims.append(tuple(d // 2 for d in dims[-1]))
    level = {"init_conv": 0, "down1": 1, "down2": 2, "down3": 3, "bottleneck": 3, "up1": 2, "up2": 1, "up3": 0}
    out = {"dw": 0, "pw": 0, "conv3": 0, "convt": 0, "head": 0}
    for name, prefix, cin, cout, ag in block_specs(cfg):
        d = dims[level[name]]
        vox = d[0] * d[1] * d[2]
        if name.startswith("up"):
            lower = dims[level[name] + 1]
            out["convt"] += lower[0] * lower[1] * lower[2] * cin * (cin // 2) * 8
        for which, (ci, co) in ((1, (cin, cout)), (2, (cout, cout))):
            kind = conv_kind(cfg, ci, co, which, ag)
            if kind == "dws":
                out["dw"] += vox * ci * 27
                out["pw"] += vox * ci * co
            elif kind == "grouped":
                out["conv3"] += vox * 27 * (ci // cfg.groups) * co
            else:
                out["conv3"] += vox * 27 * ci * co
        if cin != cout:
            out["pw"] += vox * cin * cout
    out["head"] = dims[0][0] * dims[0][1] * dims[0][2] * cfg.encoder_channels[0] * cfg.out_channels
    out["total"] = sum(out.values())
    return out
