"""Focal Tversky loss restatement (test infrastructure -- see oracle/__init__.py).

Follows /root/reference/light_unet/models/losses.py:11-54 (FocalTverskyLoss),
:57-86 (CombinedLoss), :88-113 (DiceLoss).
"""
from __future__ import annotations

import numpy as np
import torch


def focal_tversky(pred: torch.Tensor, target: torch.Tensor, alpha=0.7, beta=0.3, gamma=0.75, smooth=1e-6):
    """losses.py:40-54 -- three global sums over every element of the batch,
    then (1 - TI)^gamma.  Differentiable (autograd) for the gradient oracle."""
    p = pred.reshape(-1)
    t = target.reshape(-1)
    tp = (p * t).sum()                     # losses.py:44
    fp = (p * (1 - t)).sum()               # losses.py:45
    fn = ((1 - p) * t).sum()               # losses.py:46
    ti = (tp + smooth) / (tp + alpha * fn + beta * fp + smooth)   # losses.py:49
    return (1 - ti) ** gamma               # losses.py:52


def focal_tversky_closed_form_grad(pred: np.ndarray, target: np.ndarray, alpha=0.7, beta=0.3, gamma=0.75,
                                   smooth=1e-6):
    """float64 closed form of dL/dp used to cross-check autograd:
    with N = TP+s, D = TP + a*FN + b*FP + s,
      dN/dp_i = t_i,   dD/dp_i = t_i - a*t_i + b*(1 - t_i) = (1-a-b)*t_i + b   (general a,b)
      dL/dp_i = -gamma*(1-TI)^(gamma-1) * (t_i*D - N*dD_i) / D^2."""
    p = pred.astype(np.float64).ravel()
    t = target.astype(np.float64).ravel()
    tp = (p * t).sum()
    fp = (p * (1 - t)).sum()
    fn = ((1 - p) * t).sum()
    n = tp + smooth
    d = tp + alpha * fn + beta * fp + smooth
    ti = n / d
    dd = (1.0 - alpha - beta) * t + beta
    g = -gamma * (1 - ti) ** (gamma - 1) * (t * d - n * dd) / (d * d)
    return (1 - ti) ** gamma, g.reshape(pred.shape)


def dice(pred, target, smooth=1e-6):
    """losses.py:106-113."""
    p = pred.reshape(-1)
    t = target.reshape(-1)
    inter = (p * t).sum()
    union = p.sum() + t.sum()
    return 1.0 - (2.0 * inter + smooth) / (union + smooth)


def combined(pred, target, ftl_weight=0.8, bce_weight=0.2, alpha=0.7, beta=0.3, gamma=0.75):
    """losses.py:83-86."""
    ftl = focal_tversky(pred, target, alpha, beta, gamma)
    bce = torch.nn.functional.binary_cross_entropy(pred.reshape(-1), target.reshape(-1))
    return ftl_weight * ftl + bce_weight * bce
