"""Connected components on the GPU -- drop-in for get_connected_components
(light_unet/models/metrics.py:38-63).

The lesion-matching metrics of the reference file (:66-404) are validation-time
CPU code outside this path and are not provided.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _native as nv

DEFAULT_SPACING = (4.0, 4.0, 4.0)


def label_device(mask: torch.Tensor, min_size: int = 0):
    """mask: CUDA int32 [D,H,W] (non-zero = foreground).  Returns (labels int32 [D,H,W] CUDA, n as a
    1-element CUDA int32 tensor).  6-connectivity, components smaller than `min_size` voxels removed,
    ids 1..n in raster order of each component's first voxel (scipy.ndimage.label numbering)."""
    nv.require_cuda(mask, "label_device")
    D, H, W = mask.shape
    m = mask.contiguous()
    if m.dtype != torch.int32:
        m = (m != 0).to(torch.int32)
    nvox = D * H * W
    work = torch.empty(int(nv.lib().l3d_ccl_workspace_elems(nvox)), dtype=torch.int32, device=m.device)
    labels = torch.empty(D, H, W, dtype=torch.int32, device=m.device)
    n_out = torch.zeros(1, dtype=torch.int32, device=m.device)
    nv.call("l3d_ccl_label", nv.ptr(m), D, H, W, int(min_size), nv.ptr(labels), nv.ptr(n_out), nv.ptr(work),
            nv.stream_ptr(m.device), algo_bytes=8 * nvox)          # mask read once + labels written once
    return labels, n_out


def get_connected_components(mask, min_size=0):
    """Reference signature: binary mask (ndarray) -> (labeled int32 ndarray, num_components)."""
    if not torch.cuda.is_available():
        raise nv.NativeError("get_connected_components: the B200-native path needs a CUDA device (no CPU fallback)")
    arr = np.asarray(mask)
    shape = arr.shape
    if arr.ndim > 3 or arr.ndim == 0:
        raise ValueError("get_connected_components supports 1-D, 2-D and 3-D masks")
    arr3 = arr.reshape((1,) * (3 - arr.ndim) + shape)   # face connectivity is unchanged by singleton axes
    dev = torch.device("cuda", torch.cuda.current_device())
    m = torch.from_numpy(np.ascontiguousarray(arr3 != 0).astype(np.int32)).to(dev)
    labels, n = label_device(m, int(min_size) if min_size > 0 else 0)
    return labels.cpu().numpy().reshape(shape), int(n.item())
