"""Sliding-window grid, Gaussian importance map and overlap stitching
(test infrastructure -- see oracle/__init__.py).

Follows /root/reference/light_unet/utils.py:
  window grid  :47-81, per-window loop :86-134, normalisation :137,
  _get_gaussian_importance_map :142-173.
"""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple

import numpy as np


def axis_positions(dim: int, patch: int, overlap: float) -> List[int]:
    """Start offsets along one axis (utils.py:47-49, :63-81)."""
    stride = max(1, int(patch * (1 - overlap)))
    pos = list(range(0, max(0, dim - patch + 1), stride)) if dim >= patch else []
    if dim > patch and (len(pos) == 0 or pos[-1] + patch < dim):
        pos.append(dim - patch)           # tail window flush with the far edge
    if len(pos) == 0:
        pos = [0]                          # volume shorter than the patch: one padded window
    return pos


def window_grid(shape: Sequence[int], patch: Sequence[int], overlap: float):
    return tuple(axis_positions(d, p, overlap) for d, p in zip(shape, patch))


def gaussian_1d(length: int) -> np.ndarray:
    """utils.py:153-160 -- float64; centre at length/2 (not (length-1)/2, so the
    profile is asymmetric), sigma = length/6."""
    center = length / 2.0
    sigma = length / 6.0
    x = np.arange(length)
    return np.exp(-((x - center) ** 2) / (2 * sigma ** 2))


def gaussian_importance_map(patch: Sequence[int]) -> np.ndarray:
    """utils.py:163-173 -- float64 outer product, divide by max, cast fp32."""
    gz, gy, gx = (gaussian_1d(p) for p in patch)
    m = gz[:, None, None] * gy[None, :, None] * gx[None, None, :]
    m = m / m.max()
    return m.astype(np.float32)


def extract_window(image: np.ndarray, z: int, y: int, x: int, patch) -> np.ndarray:
    """utils.py:91-112 -- slice, zero-pad at the END of each short axis."""
    pd, ph, pw = patch
    sub = image[z:z + pd, y:y + ph, x:x + pw]
    if sub.shape != tuple(patch):
        full = np.zeros(patch, dtype=image.dtype)
        full[:sub.shape[0], :sub.shape[1], :sub.shape[2]] = sub
        sub = full
    return sub


def stitch(shape, patch, positions, preds: np.ndarray, use_gaussian: bool = True) -> np.ndarray:
    """Accumulate per-window predictions in the reference's z->y->x window order
    (utils.py:86-88), fp32 ``prob += pred*w; cnt += w`` (:133-134), then
    ``prob/cnt where cnt>0`` (:137).  ``preds`` is [n_windows, pd, ph, pw] in
    the same window order."""
    d, h, w = shape
    pd, ph, pw = patch
    imp = gaussian_importance_map(patch) if use_gaussian else np.ones(patch, dtype=np.float32)
    prob = np.zeros(shape, dtype=np.float32)
    cnt = np.zeros(shape, dtype=np.float32)
    i = 0
    for z in positions[0]:
        for y in positions[1]:
            for x in positions[2]:
                ze, ye, xe = min(z + pd, d), min(y + ph, h), min(x + pw, w)
                ad, ah, aw = ze - z, ye - y, xe - x
                wts = imp[:ad, :ah, :aw]
                prob[z:ze, y:ye, x:xe] += preds[i][:ad, :ah, :aw] * wts
                cnt[z:ze, y:ye, x:xe] += wts
                i += 1
    return np.divide(prob, cnt, where=cnt > 0, out=prob)


def sliding_window(image: np.ndarray, predict: Callable[[np.ndarray], np.ndarray], patch=(48, 48, 48),
                   overlap: float = 0.5, use_gaussian: bool = True, batch: int = 1) -> np.ndarray:
    """utils.py:11-139 with the model abstracted as ``predict`` ([n,1,pd,ph,pw]
    float32 -> [n,1,pd,ph,pw] float32).  batch=1 reproduces the reference's
    per-window forward; InstanceNorm is per-sample so batching windows changes
    nothing beyond fp32 round-off."""
    if image.ndim == 4 and image.shape[0] == 1:
        image = image[0]
    if image.ndim != 3:
        raise ValueError(f"Expected 3D image [D, H, W], got shape {image.shape}")
    positions = window_grid(image.shape, patch, overlap)
    wins = [extract_window(image, z, y, x, patch)
            for z in positions[0] for y in positions[1] for x in positions[2]]
    preds = []
    for i in range(0, len(wins), batch):
        chunk = np.stack(wins[i:i + batch])[:, None].astype(np.float32)
        preds.append(predict(chunk)[:, 0])
    preds = np.concatenate(preds)
    return stitch(image.shape, patch, positions, preds, use_gaussian)
