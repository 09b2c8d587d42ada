"""Patch extraction and augmentation of the training data pipeline (test infrastructure -- see oracle/__init__.py).

Follows /root/reference/light_unet/datasets/patch_dataset.py: _extract_patch :136-154, _augment :156-220.  The two
resampling steps go through scipy.ndimage there (rotate, order 1 / 0, reshape=False, mode='constant'; zoom, order 1 / 0,
mode='constant'); their arithmetic is restated here in numpy, operation for operation in float64 (scipy 1.18:
ni_interpolation.c NI_GeometricTransform / NI_ZoomShift, spline order <= 1 so no prefilter), and pinned against scipy itself
in tests/test_augment_ref.py.
"""
from __future__ import annotations

import math

import numpy as np


def extract_patch(image, label, center, patch_size):
    """patch_dataset.py:136-154: window starting at max(0, c - p // 2), clipped at the far edge, zero-padded at the END."""
    pz, py, px = patch_size
    z, y, x = (int(c) for c in center)
    zs, ys, xs = max(0, z - pz // 2), max(0, y - py // 2), max(0, x - px // 2)
    ze, ye, xe = min(image.shape[0], zs + pz), min(image.shape[1], ys + py), min(image.shape[2], xs + px)
    ip, lp = image[zs:ze, ys:ye, xs:xe], label[zs:ze, ys:ye, xs:xe]
    if ip.shape != tuple(patch_size):
        pad = [(0, pz - ip.shape[0]), (0, py - ip.shape[1]), (0, px - ip.shape[2])]
        ip, lp = np.pad(ip, pad), np.pad(lp, pad)
    return ip, lp


def cos_sin_degrees(angle):
    """scipy.ndimage.rotate uses special.cosdg / sindg (cephes, exact at multiples of 90 degrees)."""
    try:
        from scipy import special
        return float(special.cosdg(angle)), float(special.sindg(angle))
    except ImportError:                                           # pragma: no cover
        return math.cos(math.radians(angle)), math.sin(math.radians(angle))


def _interp_axis(cc, n, order):
    """Per-axis sampling of scipy's 'constant' mode for spline order 0 / 1: coordinates outside [0, n - 1] select cval;
    inside, order 1 blends floor(cc) and floor(cc) + 1 (mirrored back at the far edge, where its weight is 0) with
    weights (1 - f, f), order 0 takes floor(cc + 0.5).  Returns (inside, i0, i1, w0, w1)."""
    inside = ~((cc < 0) | (cc > n - 1))
    c = np.where(inside, cc, 0.0)
    if order == 1:
        fl = np.floor(c)
        f = c - fl
        i0 = fl.astype(np.int64)
        i1 = i0 + 1
        i1 = np.where(i1 >= n, np.maximum(2 * n - 2 - i1, 0), i1)
        return inside, i0, i1, 1.0 - f, f
    i0 = np.floor(c + 0.5).astype(np.int64)
    i0 = np.where(i0 >= n, np.maximum(2 * n - 2 - i0, 0), i0)
    return inside, i0, i0, np.ones_like(c), np.zeros_like(c)


def rotate(arr, angle, axes, order):
    """scipy.ndimage.rotate(arr, angle, axes=axes, reshape=False, order=order, mode='constant', cval=0) for a 3-D array."""
    a0, a1 = sorted(int(a) % 3 for a in axes)
    c, s = cos_sin_degrees(angle)
    m = np.array([[c, s], [-s, c]])
    shp = np.array([arr.shape[a0], arr.shape[a1]], dtype=np.float64)
    out_center = m @ ((shp - 1) / 2)
    offset = (shp - 1) / 2 - out_center
    o0, o1 = np.meshgrid(np.arange(arr.shape[a0], dtype=np.float64), np.arange(arr.shape[a1], dtype=np.float64), indexing="ij")
    cc0 = offset[0] + o0 * m[0, 0]
    cc0 = cc0 + o1 * m[0, 1]
    cc1 = offset[1] + o0 * m[1, 0]
    cc1 = cc1 + o1 * m[1, 1]
    in0, i00, i01, w00, w01 = _interp_axis(cc0, arr.shape[a0], order)
    in1, i10, i11, w10, w11 = _interp_axis(cc1, arr.shape[a1], order)
    inside = in0 & in1
    src = np.moveaxis(arr, (a0, a1), (0, 1)).astype(np.float64)           # [n0, n1, rest]
    if order == 1:
        t = src[i00, i10] * w00[..., None] * w10[..., None]
        t = t + src[i00, i11] * w00[..., None] * w11[..., None]
        t = t + src[i01, i10] * w01[..., None] * w10[..., None]
        t = t + src[i01, i11] * w01[..., None] * w11[..., None]
    else:
        t = src[i00, i10]
    t = np.where(inside[..., None], t, 0.0)
    return np.moveaxis(t, (0, 1), (a0, a1)).astype(arr.dtype)


def zoom(arr, factor, order):
    """scipy.ndimage.zoom(arr, factor, order=order, mode='constant', cval=0) for a 3-D array and a scalar factor."""
    out_shape = tuple(int(round(n * factor)) for n in arr.shape)
    parts = []
    for n, on in zip(arr.shape, out_shape):
        z = (n - 1) / (on - 1) if on > 1 else 1.0                          # grid_mode=False: corners map onto corners
        cc = np.arange(on, dtype=np.float64) * z
        parts.append(_interp_axis(cc, n, order))
    src = arr.astype(np.float64)
    (inz, z0, z1, wz0, wz1), (iny, y0, y1, wy0, wy1), (inx, x0, x1, wx0, wx1) = parts
    ix = lambda a, b, c: src[a[:, None, None], b[None, :, None], c[None, None, :]]
    W = lambda a, b, c: a[:, None, None] * b[None, :, None] * c[None, None, :]
    if order == 1:
        t = ix(z0, y0, x0) * wz0[:, None, None] * wy0[None, :, None] * wx0[None, None, :]
        for (zi, wz), (yi, wy), (xi, wx) in [((z0, wz0), (y0, wy0), (x1, wx1)), ((z0, wz0), (y1, wy1), (x0, wx0)), ((z0, wz0), (y1, wy1), (x1, wx1)),
                                             ((z1, wz1), (y0, wy0), (x0, wx0)), ((z1, wz1), (y0, wy0), (x1, wx1)), ((z1, wz1), (y1, wy1), (x0, wx0)),
                                             ((z1, wz1), (y1, wy1), (x1, wx1))]:
            t = t + ix(zi, yi, xi) * wz[:, None, None] * wy[None, :, None] * wx[None, None, :]
    else:
        t = ix(z0, y0, x0)
    inside = inz[:, None, None] & iny[None, :, None] & inx[None, None, :]
    return np.where(inside, t, 0.0).astype(arr.dtype)


def fit_to_patch(arr, patch_size):
    """patch_dataset.py:186-208: centre-crop axes that grew, zero-pad (at the end) axes that shrank."""
    for ax, p in enumerate(patch_size):
        n = arr.shape[ax]
        if n > p:
            st = (n - p) // 2
            arr = np.take(arr, np.arange(st, st + p), axis=ax)
    pad = [(0, max(0, p - n)) for n, p in zip(arr.shape, patch_size)]
    if any(b for _, b in pad):
        arr = np.pad(arr, pad)
    return arr


def apply(image, label, ops, patch_size):
    """Apply a recorded list of augmentation decisions in the reference's order (patch_dataset.py:160-218).
    ops: dict with optional keys flip=axis, rotate=(angle, axes), scale=factor, shift=value, noise=float64 array."""
    if "flip" in ops:
        image, label = np.flip(image, axis=ops["flip"]).copy(), np.flip(label, axis=ops["flip"]).copy()
    if "rotate" in ops:
        ang, axes = ops["rotate"]
        image, label = rotate(image, ang, axes, 1), rotate(label, ang, axes, 0)
    if "scale" in ops:
        image, label = zoom(image, ops["scale"], 1), zoom(label, ops["scale"], 0)
        if image.shape != tuple(patch_size):
            image, label = fit_to_patch(image, patch_size), fit_to_patch(label, patch_size)
    if "shift" in ops:
        image = np.clip(image + ops["shift"], 0, 1)
    if "noise" in ops:
        image = np.clip(image + ops["noise"], 0, 1)
    return image, label


PATCH_AUG = {"random_flip": {"enabled": True, "prob": 0.5, "axes": [0, 1, 2]},
             "random_rotation": {"enabled": True, "prob": 0.5, "angle_range": [-15, 15], "axes": [[0, 1], [0, 2], [1, 2]]},
             "random_scale": {"enabled": True, "prob": 0.3, "scale_range": [0.9, 1.1]},
             "intensity_shift": {"enabled": True, "prob": 0.5, "shift_range": [-0.1, 0.1]},
             "gaussian_noise": {"enabled": True, "prob": 0.3, "sigma": 0.01}}


def synth_cases():
    """Seeded synthetic cases (image in [0, 1], sparse blob labels, a body mask for one of them): the volumes behind
    tests/golden/patches.json."""
    from . import metrics_ref
    vols = []
    for i, shape in enumerate([(40, 44, 52), (36, 50, 41), (30, 30, 30)]):
        prob, label = metrics_ref.synth_case(shape, 70 + i)
        rng = np.random.default_rng(80 + i)
        image = np.clip(0.6 * prob + 0.4 * rng.random(shape, dtype=np.float32), 0, 1).astype(np.float32)
        body = np.ones(shape, dtype=np.float32)
        body[:, :, : shape[2] // 5] = 0
        vols.append((image, label.astype(np.float32), body if i == 1 else None))
    return vols


class RefSampler:
    """PatchDataset (patch_dataset.py:17-220) over in-memory volumes, with private generators seeded the way the reference
    seeds the global ones (np.random.seed(seed) / random.seed(seed)): the same draws in the same order."""

    def __init__(self, volumes, patch_size, lesion_patch_ratio=0.5, augmentation=None, seed=42):
        import random
        self.volumes, self.patch_size = volumes, tuple(patch_size)
        self.ratio, self.aug = lesion_patch_ratio, augmentation
        self.rs, self.pr = np.random.RandomState(seed), random.Random(seed)
        self.lesion, self.background = [], []
        for ci, vol in enumerate(volumes):                         # :72-99
            label = vol[1].astype(np.float64)
            body = vol[2].astype(bool) if len(vol) > 2 and vol[2] is not None else None
            lc = np.argwhere(label > 0)
            if len(lc) > 0:
                for idx in self.rs.randint(len(lc), size=max(10, len(lc) // 1000)):
                    self.lesion.append((ci, lc[idx]))
            bc = np.argwhere((label == 0) & body) if body is not None else np.argwhere(label == 0)
            if len(bc) > 0:
                for idx in self.rs.randint(len(bc), size=max(10, len(bc) // 5000)):
                    self.background.append((ci, bc[idx]))

    def item(self):
        """One __getitem__ (:114-134, :156-220) -> (image patch, label patch) as float32."""
        rs, pr, aug = self.rs, self.pr, self.aug
        if rs.rand() < self.ratio and len(self.lesion) > 0:
            ci, center = self.lesion[rs.randint(len(self.lesion))]
        elif len(self.background) > 0:
            ci, center = self.background[rs.randint(len(self.background))]
        else:
            ci, center = self.lesion[rs.randint(len(self.lesion))]
        image, label = self.volumes[ci][0].astype(np.float32), self.volumes[ci][1].astype(np.float32)
        ip, lp = extract_patch(image, label, center, self.patch_size)
        ops = {}
        if aug:
            if aug.get("random_flip", {}).get("enabled", False) and rs.rand() < aug["random_flip"].get("prob", 0.5):
                ops["flip"] = pr.choice(aug["random_flip"].get("axes", [0, 1, 2]))
            if aug.get("random_rotation", {}).get("enabled", False) and rs.rand() < aug["random_rotation"].get("prob", 0.5):
                lo, hi = aug["random_rotation"].get("angle_range", [-15, 15])
                ang = rs.uniform(lo, hi)
                ops["rotate"] = (ang, pr.choice(aug["random_rotation"].get("axes", [[0, 1], [0, 2], [1, 2]])))
            if aug.get("random_scale", {}).get("enabled", False) and rs.rand() < aug["random_scale"].get("prob", 0.3):
                lo, hi = aug["random_scale"].get("scale_range", [0.9, 1.1])
                ops["scale"] = rs.uniform(lo, hi)
            if aug.get("intensity_shift", {}).get("enabled", False) and rs.rand() < aug["intensity_shift"].get("prob", 0.5):
                lo, hi = aug["intensity_shift"].get("shift_range", [-0.1, 0.1])
                ops["shift"] = rs.uniform(lo, hi)
            if aug.get("gaussian_noise", {}).get("enabled", False) and rs.rand() < aug["gaussian_noise"].get("prob", 0.3):
                ops["noise"] = rs.normal(0, aug["gaussian_noise"].get("sigma", 0.01), ip.shape)
        ip, lp = apply(ip, lp, ops, self.patch_size)
        return np.ascontiguousarray(ip, dtype=np.float32), np.ascontiguousarray(lp, dtype=np.float32), ops
