"""Deterministic synthetic inputs shared by the golden generator, the tests,
``smoke()`` and ``bench.py``.  (Test infrastructure -- see oracle/__init__.py.)

Everything is drawn from numpy's PCG64 (``np.random.default_rng``), whose stream
is stable across numpy versions and machines, so a fixture only has to store the
*outputs*; weights and inputs are regenerated from the seed on both sides.
"""
from __future__ import annotations

import zlib
from collections import OrderedDict

import numpy as np


def _key_seed(seed: int, key: str) -> int:
    return (int(seed) * 1000003 + zlib.crc32(key.encode())) & 0x7FFFFFFF


def synth_state_dict(shapes: "OrderedDict[str, tuple]", seed: int = 0) -> "OrderedDict[str, np.ndarray]":
    """Fill every parameter of ``shapes`` (see ``unet_ref.param_shapes``) with
    non-trivial values.  Conv weights ~ N(0, 1/fan_in); InstanceNorm gamma ~
    1 + 0.2 N, beta ~ 0.2 N; conv biases ~ 0.1 N.  Non-identity affine
    parameters matter: the reference's default init (gamma=1, beta=0) would hide
    scale/shift bugs."""
    out = OrderedDict()
    for key, shape in shapes.items():
        rng = np.random.default_rng(_key_seed(seed, key))
        if len(shape) == 5:  # conv / transposed-conv weight
            fan_in = int(np.prod(shape[1:]))
            if ".up.weight" in key:  # ConvTranspose3d weight is (Cin, Cout, 2,2,2)
                fan_in = int(shape[0])
            w = rng.standard_normal(shape) / np.sqrt(max(fan_in, 1))
        elif ("norm" in key or "shortcut.1" in key) and key.endswith(".weight"):
            w = 1.0 + 0.2 * rng.standard_normal(shape)
        elif ("norm" in key or "shortcut.1" in key) and key.endswith(".bias"):
            w = 0.2 * rng.standard_normal(shape)
        else:  # conv biases
            w = 0.1 * rng.standard_normal(shape)
        out[key] = w.astype(np.float32)
    return out


def synth_patches(batch: int, size, seed: int = 42, lesion_frac: float = 0.02):
    """Images in [0,1] (preprocessed PET range, reference
    scripts/preprocess_data.py:43-45) and ~2 % positive binary targets
    (SURVEY.md section 8(d), config C1)."""
    if isinstance(size, int):
        size = (size, size, size)
    rng = np.random.default_rng(seed)
    x = rng.random((batch, 1) + tuple(size), dtype=np.float32)
    t = (rng.random((batch, 1) + tuple(size), dtype=np.float32) > (1.0 - lesion_frac)).astype(np.float32)
    return x, t


def synth_volume(shape=(128, 128, 320), seed: int = 42, n_blobs: int = 6):
    """Whole-body-like synthetic PET volume: uniform noise in [0, 0.3] plus a
    few bright Gaussian blobs."""
    rng = np.random.default_rng(seed)
    vol = 0.3 * rng.random(shape, dtype=np.float32)
    zz, yy, xx = np.meshgrid(*[np.arange(s, dtype=np.float32) for s in shape], indexing="ij")
    for _ in range(n_blobs):
        c = [rng.uniform(0.15 * s, 0.85 * s) for s in shape]
        sig = rng.uniform(2.0, 5.0)
        amp = rng.uniform(0.4, 0.7)
        vol += (amp * np.exp(-((zz - c[0]) ** 2 + (yy - c[1]) ** 2 + (xx - c[2]) ** 2) / (2 * sig * sig))).astype(np.float32)
    return np.clip(vol, 0.0, 1.0).astype(np.float32)


def synth_prob_map(shape=(40, 48, 56), seed: int = 7, n_blobs: int = 12):
    """Synthetic probability map with blobs of assorted sizes (some below the
    min-volume filter, some touching, some on the border) for the
    threshold -> connected-components -> bounding-box stage."""
    rng = np.random.default_rng(seed)
    prob = (0.25 * rng.random(shape, dtype=np.float32)).astype(np.float32)
    zz, yy, xx = np.meshgrid(*[np.arange(s, dtype=np.float32) for s in shape], indexing="ij")
    for _ in range(n_blobs):
        c = [rng.uniform(0, s - 1) for s in shape]
        r = rng.uniform(0.8, 4.5)
        amp = rng.uniform(0.35, 0.95)
        d2 = (zz - c[0]) ** 2 + (yy - c[1]) ** 2 + (xx - c[2]) ** 2
        prob = np.maximum(prob, (amp * (d2 <= r * r)).astype(np.float32))
    # isolated speckle above threshold (removed by the min-size filter)
    speck = rng.random(shape) > 0.997
    prob[speck] = 0.6
    return prob.astype(np.float32)
