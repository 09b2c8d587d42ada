"""Threshold -> connected components -> bounding boxes restatement
(test infrastructure -- see oracle/__init__.py).

Follows /root/reference/light_unet/models/metrics.py:38-63
(get_connected_components) and /root/reference/light_unet/core/inferencer.py:62-111
(Inferencer.extract_bboxes).  The labelling itself is oracle/ccl_ref.c.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import List, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libccl_ref.so")
_lib = None


def build(force: bool = False) -> str:
    """gcc -O2 -shared oracle/ccl_ref.c -> oracle/libccl_ref.so"""
    src = os.path.join(_HERE, "ccl_ref.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", _LIB_PATH, src])
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.ccl6_label.restype = ctypes.c_int
        _lib.ccl6_label.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    return _lib


def label6(mask: np.ndarray) -> Tuple[np.ndarray, int]:
    """scipy.ndimage.label(mask) with the default (6-connected) structure."""
    m = np.ascontiguousarray(mask != 0, dtype=np.int32)
    if m.ndim != 3:
        raise ValueError("label6 expects a 3-D array")
    out = np.zeros(m.shape, dtype=np.int32)
    n = _load().ccl6_label(m.ctypes.data, out.ctypes.data, *m.shape)
    if n < 0:
        raise MemoryError("ccl6_label")
    return out, int(n)


def connected_components(mask: np.ndarray, min_size: int = 0) -> Tuple[np.ndarray, int]:
    """metrics.py:50-63 -- label; if min_size > 0 zero every component smaller
    than min_size voxels and label again (numbering = raster order of the
    survivors)."""
    labeled, n = label6(mask)
    if min_size > 0:
        sizes = np.bincount(labeled.ravel())
        small = sizes < min_size
        small[0] = False
        labeled[small[labeled]] = 0
        labeled, n = label6(labeled > 0)
    return labeled, n


def extract_bboxes(prob_map: np.ndarray, threshold=0.3, min_volume_cc=0.5, spacing=(4.0, 4.0, 4.0),
                   expansion_voxels: int = 3) -> List[dict]:
    """inferencer.py:62-111.  ``expansion_voxels`` is
    config["data"]["bbox_expansion_voxels"] (:83)."""
    binary = (prob_map >= threshold).astype(np.int32)            # :64 (float32 array vs python float)
    voxel_cc = spacing[0] * spacing[1] * spacing[2] / 1000.0      # :66-67
    min_voxels = int(np.ceil(min_volume_cc / voxel_cc))            # :68
    labeled, n = connected_components(binary, min_voxels)          # :70
    out = []
    for cid in range(1, n + 1):
        comp = labeled == cid
        coords = np.argwhere(comp)
        if len(coords) == 0:
            continue
        lo = coords.min(axis=0)
        hi = coords.max(axis=0)
        box = []
        for ax in range(3):                                        # :83-89
            box.append(max(0, lo[ax] - expansion_voxels))
            box.append(min(prob_map.shape[ax] - 1, hi[ax] + expansion_voxels))
        mm = [box[2 * ax + k] * spacing[ax] for ax in range(3) for k in range(2)]   # :91-96
        vol_cc = comp.sum() * voxel_cc                              # :98-99
        conf = prob_map[comp].max()                                 # :100
        out.append({
            "mask_id": cid,
            "bbox_voxel": [int(v) for v in box],
            "bbox_mm": [float(v) for v in mm],
            "volume_cc": float(vol_cc),
            "confidence": float(conf),
        })
    return out
