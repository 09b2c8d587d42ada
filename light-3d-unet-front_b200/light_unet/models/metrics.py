"""Evaluation metrics -- drop-in for light_unet/models/metrics.py.

get_connected_components (:38-63) and the lesion-wise metrics (:66-404) run their O(volume) parts on
the GPU: thresholding, 6-connected labelling (l3d_ccl_label) and ONE pass over the two label maps that yields the
pair-intersection histogram, the component sizes and the coordinate sums (l3d_label_pair_stats) -- what the reference
gets from scipy.ndimage.label, three np.bincount calls and ndimage.center_of_mass.  All of these are integers; the
remaining arithmetic (IoU in float32, centres / distances in float64, the greedy matching, the ratios) is the reference's
own numpy code on (n_pred x n_target)-sized arrays, so every returned number is bit-identical to the reference's.
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


SMOOTH = 1e-6
SPATIAL_DIMENSIONS = 3


def calculate_dsc(pred, target, smooth=SMOOTH):
    """metrics.py:15-35 (host arithmetic on the caller's arrays, kept for API compatibility; calculate_metrics derives the
    Dice terms from the device-side pair statistics instead)."""
    pred = np.ravel(pred)
    target = np.ravel(target)
    intersection = (pred * target).sum()
    union = pred.sum() + target.sum()
    return (2.0 * intersection + smooth) / (union + smooth)


def calculate_iou(pred_component, target_component):
    """metrics.py:66-74."""
    intersection = np.logical_and(pred_component, target_component).sum()
    union = np.logical_or(pred_component, target_component).sum()
    if union == 0:
        return 0.0
    return intersection / union


def _device():
    if not torch.cuda.is_available():
        raise nv.NativeError("light_unet.models.metrics: the B200-native path needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _as_device_int32(a, dev):
    if isinstance(a, torch.Tensor):
        return a.to(device=dev, dtype=torch.int32).contiguous()
    return torch.from_numpy(np.ascontiguousarray(a).astype(np.int32, copy=False)).to(dev)


def _vol3(t):
    while t.dim() > 3 and t.shape[0] == 1:
        t = t[0]
    if t.dim() < 3:
        t = t.reshape((1,) * (3 - t.dim()) + tuple(t.shape))
    if t.dim() != 3:
        raise ValueError(f"expected a [D, H, W] volume, got shape {tuple(t.shape)}")
    return t


def pair_stats_device(la: torch.Tensor, na: int, lb: torch.Tensor = None, nb: int = 0):
    """One pass over label map(s) on the device.  Returns host arrays: counts int64 [(na+1), (nb+1)] (None without lb),
    mom_a int64 [na+1, 4] = {voxels, sum z, sum y, sum x} per component id, mom_b likewise (None without lb)."""
    la = _vol3(la)
    D, H, W = la.shape
    dev = la.device
    mom_a = torch.zeros(na + 1, 4, dtype=torch.int64, device=dev)
    counts = mom_b = None
    if lb is not None:
        lb = _vol3(lb)
        if tuple(lb.shape) != (D, H, W):
            raise ValueError(f"label maps differ in shape: {tuple(la.shape)} vs {tuple(lb.shape)}")
        if (na + 1) * (nb + 1) > (1 << 28):
            raise ValueError(f"{na} x {nb} component pairs: intersection histogram too large")
        counts = torch.zeros(na + 1, nb + 1, dtype=torch.int32, device=dev)
        mom_b = torch.zeros(nb + 1, 4, dtype=torch.int64, device=dev)
    nv.call("l3d_label_pair_stats", nv.ptr(la), nv.ptr(lb), D, H, W, int(na), int(nb), nv.ptr(counts), nv.ptr(mom_a), nv.ptr(mom_b),
            nv.stream_ptr(dev), algo_bytes=(8 if lb is not None else 4) * D * H * W)
    return (None if counts is None else counts.cpu().numpy().astype(np.int64), mom_a.cpu().numpy(),
            None if mom_b is None else mom_b.cpu().numpy())


def _centers_from_moments(mom):
    """ndimage.center_of_mass with unit weights (metrics.py:111-124): coordinate sums / voxel count, in float64."""
    n = mom.shape[0] - 1
    if n <= 0:
        return np.empty((0, 3), dtype=np.float64)
    return mom[1:, 1:4].astype(np.float64) / mom[1:, 0:1].astype(np.float64)


def _compute_component_centers(labeled):
    """metrics.py:107-124 for a host (or device) labelled volume."""
    dev = _device()
    la = _vol3(_as_device_int32(labeled, dev))
    if la.numel() == 0:
        return np.empty((0, 3), dtype=np.float64)
    n = int(la.max().item())
    if n == 0:
        return np.empty((0, 3), dtype=np.float64)
    return _centers_from_moments(pair_stats_device(la, n)[1])


def calculate_center_distance(pred_component, target_component, spacing=(1.0, 1.0, 1.0)):
    """metrics.py:77-104."""
    dev = _device()
    a = (_vol3(_as_device_int32(np.asarray(pred_component) != 0, dev)))
    b = (_vol3(_as_device_int32(np.asarray(target_component) != 0, dev)))
    _, ma, mb = pair_stats_device(a, 1, b, 1)
    with np.errstate(invalid="ignore", divide="ignore"):
        ca, cb = ma[1, 1:4] / np.float64(ma[1, 0]), mb[1, 1:4] / np.float64(mb[1, 0])
    return np.linalg.norm(ca * np.array(spacing) - cb * np.array(spacing))


def _match_from_stats(counts, mom_p, mom_t, iou_threshold, distance_threshold_mm, spacing):
    """The host part of match_components (metrics.py:150-229) on the integer statistics."""
    num_pred, num_target = counts.shape[0] - 1, counts.shape[1] - 1
    intersection = counts.copy()
    intersection[0, :] = 0
    intersection[:, 0] = 0
    pred_sizes, target_sizes = mom_p[:, 0].copy(), mom_t[:, 0].copy()
    union = pred_sizes[:, None] + target_sizes[None, :] - intersection
    iou_matrix = np.divide(intersection, union, out=np.zeros_like(intersection, dtype=np.float32), where=union > 0)
    spacing_arr = np.asarray(spacing, dtype=np.float64)
    pred_centers = _centers_from_moments(mom_p) * spacing_arr
    target_centers = _centers_from_moments(mom_t) * spacing_arr
    if pred_centers.size and target_centers.size:
        diff = pred_centers[:, None, :] - target_centers[None, :, :]
        distance_matrix = np.linalg.norm(diff, axis=2)
    else:
        distance_matrix = np.full((num_pred, num_target), np.inf, dtype=np.float64)
    matches, matched_pred = [], set()
    matched_target_mask = np.zeros(num_target, dtype=bool)
    for pred_id in range(1, num_pred + 1):
        iou_row = iou_matrix[pred_id, 1:]
        if distance_matrix.size > 0:
            distance_criteria = distance_matrix[pred_id - 1] <= distance_threshold_mm
            valid_mask = ~matched_target_mask & ((iou_row >= iou_threshold) | distance_criteria)
        else:
            valid_mask = ~matched_target_mask & (iou_row >= iou_threshold)
        if not np.any(valid_mask):
            continue
        candidate_ious = np.where(valid_mask, iou_row, -np.inf)
        best_target_idx = int(np.argmax(candidate_ious))
        matches.append((pred_id, best_target_idx + 1))
        matched_pred.add(pred_id)
        matched_target_mask[best_target_idx] = True
    unmatched_pred = [i for i in range(1, num_pred + 1) if i not in matched_pred]
    unmatched_target = [i for i in range(1, num_target + 1) if not matched_target_mask[i - 1]]
    return matches, unmatched_pred, unmatched_target


def match_components(pred_labeled, target_labeled, iou_threshold=0.1, distance_threshold_mm=10.0, spacing=(4.0, 4.0, 4.0)):
    """metrics.py:127-229: (matches, unmatched_pred, unmatched_target) for two labelled volumes (host arrays or CUDA
    tensors)."""
    dev = _device()
    lp, lt = _vol3(_as_device_int32(pred_labeled, dev)), _vol3(_as_device_int32(target_labeled, dev))
    num_pred, num_target = int(lp.max().item()), int(lt.max().item())
    if num_pred == 0 or num_target == 0:
        return [], list(range(1, num_pred + 1)), list(range(1, num_target + 1))
    counts, mp, mt = pair_stats_device(lp, num_pred, lt, num_target)
    return _match_from_stats(counts, mp, mt, iou_threshold, distance_threshold_mm, spacing)


def _threshold_device(vol: torch.Tensor, threshold) -> torch.Tensor:
    """(vol >= threshold).astype(int32) with the reference's numpy promotion (metrics.py:263-264): a float32 volume is
    compared with the python float in float32."""
    v = _vol3(vol)
    if v.dtype != torch.float32:
        v = v.to(torch.float32)
    v = v.contiguous()
    mask = torch.empty(v.shape, dtype=torch.int32, device=v.device)
    nv.call("l3d_threshold", nv.ptr(v), v.numel(), float(np.float32(threshold)), nv.ptr(mask), nv.stream_ptr(v.device),
            algo_bytes=8 * v.numel())
    return mask


def _to_device_f32(a, dev):
    if isinstance(a, torch.Tensor):
        return a.to(device=dev, dtype=torch.float32)
    arr = np.asarray(a)
    if arr.dtype == np.float64:
        # the reference compares float64 arrays in float64; float32 holds thresholds like 0.5 exactly but not every value:
        # keep the comparison exact by thresholding on the host for this (non-hot) case
        return arr
    return torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float32)).to(dev)


def _binarize(a, threshold, dev) -> torch.Tensor:
    v = _to_device_f32(a, dev)
    if isinstance(v, np.ndarray):
        return _vol3(torch.from_numpy((v >= threshold).astype(np.int32)).to(dev))
    return _threshold_device(v, threshold)


def lesion_stats_device(pred_mask: torch.Tensor, target_labeled: torch.Tensor, num_target: int, min_size_voxels=0):
    """Device core shared by calculate_lesion_metrics and the validation sweep: label the prediction mask and gather the
    pair statistics against an already labelled target.  Returns (num_pred, counts, mom_p, mom_t)."""
    pred_labeled, n_d = label_device(pred_mask, int(min_size_voxels) if min_size_voxels > 0 else 0)
    num_pred = int(n_d.item())
    counts, mp, mt = pair_stats_device(pred_labeled, num_pred, target_labeled, num_target)
    return num_pred, counts, mp, mt


def _lesion_counts(num_pred, num_target, counts, mp, mt, iou_threshold, distance_threshold_mm, spacing):
    """tp / fp / fn and the ratios of calculate_lesion_metrics (metrics.py:271-308)."""
    if num_target == 0:
        if num_pred == 0:
            return {"recall": 1.0, "precision": 1.0, "f1": 1.0, "tp": 0, "fp": 0, "fn": 0}
        return {"recall": 0.0, "precision": 0.0, "f1": 0.0, "tp": 0, "fp": num_pred, "fn": 0}
    if num_pred == 0:
        return {"recall": 0.0, "precision": 0.0, "f1": 0.0, "tp": 0, "fp": 0, "fn": num_target}
    matches, unmatched_pred, unmatched_target = _match_from_stats(counts, mp, mt, iou_threshold, distance_threshold_mm, spacing)
    tp, fp, fn = len(matches), len(unmatched_pred), len(unmatched_target)
    recall = tp / (tp + fn) if (tp + fn) > 0 else 0.0
    precision = tp / (tp + fp) if (tp + fp) > 0 else 0.0
    f1 = 2 * (precision * recall) / (precision + recall) if (precision + recall) > 0 else 0.0
    return {"recall": recall, "precision": precision, "f1": f1, "tp": tp, "fp": fp, "fn": fn}


def _squeeze_case(a):
    """Shape handling of calculate_lesion_metrics (metrics.py:252-261)."""
    if len(a.shape) == 5:
        a = a[:, 0]
    if len(a.shape) == 4 and a.shape[0] == 1:
        a = a[0]
    return a


def calculate_lesion_metrics(pred, target, threshold=0.5, min_size_voxels=0, iou_threshold=0.1, distance_threshold_mm=10.0,
                             spacing=(4.0, 4.0, 4.0)):
    """metrics.py:232-308: {"recall", "precision", "f1", "tp", "fp", "fn"} for one case."""
    dev = _device()
    pred, target = _squeeze_case(pred), _squeeze_case(target)
    pred_mask = _binarize(pred, threshold, dev)
    target_mask = _binarize(target, 0.5, dev)
    target_labeled, nt_d = label_device(target_mask, int(min_size_voxels) if min_size_voxels > 0 else 0)
    num_target = int(nt_d.item())
    num_pred, counts, mp, mt = lesion_stats_device(pred_mask, target_labeled, num_target, min_size_voxels)
    return _lesion_counts(num_pred, num_target, counts, mp, mt, iou_threshold, distance_threshold_mm, spacing)


def _normalize_spacing_per_case(spacing, num_cases):
    """metrics.py:291-307."""
    if num_cases == 0:
        return []
    if isinstance(spacing, np.ndarray):
        spacing = spacing.tolist()
    if isinstance(spacing, (list, tuple)):
        if len(spacing) == 0:
            return [tuple(map(float, DEFAULT_SPACING)) for _ in range(num_cases)]
        if len(spacing) == num_cases and isinstance(spacing[0], (list, tuple, np.ndarray)):
            return [tuple(map(float, s)) for s in spacing]
        if len(spacing) == SPATIAL_DIMENSIONS and all(isinstance(s, (int, float, np.floating)) for s in spacing):
            return [tuple(map(float, spacing)) for _ in range(num_cases)]
    return [tuple(map(float, DEFAULT_SPACING)) for _ in range(num_cases)]


class CaseAccumulator:
    """Running totals of calculate_metrics (metrics.py:355-404) for ONE threshold, fed with per-case integer statistics."""

    def __init__(self):
        self.tp = self.fp = self.fn = 0
        self.intersection_sum = 0.0
        self.union_sum = 0.0
        self.per_case_dsc = []
        self.num_cases = 0

    def add(self, num_pred, num_target, counts, mp, mt, pred_voxels, target_voxels, spacing):
        inter = int(counts[1:, 1:].sum()) if counts is not None else 0
        # metrics.py:366-372 -- (pred_binary * target_binary).sum() etc. are integer sums; the Dice ratios are float64
        self.intersection_sum += inter
        self.union_sum += pred_voxels + target_voxels
        self.per_case_dsc.append((2.0 * inter + SMOOTH) / ((pred_voxels + target_voxels) + SMOOTH))
        m = _lesion_counts(num_pred, num_target, counts, mp, mt, 0.1, 10.0, spacing)
        self.tp += m["tp"]; self.fp += m["fp"]; self.fn += m["fn"]
        self.num_cases += 1

    def result(self):
        voxel_dsc_micro = (2.0 * self.intersection_sum + SMOOTH) / (self.union_sum + SMOOTH)
        voxel_dsc_macro = np.mean(self.per_case_dsc) if self.per_case_dsc else 0.0
        tp, fp, fn = self.tp, self.fp, self.fn
        lesion_recall = tp / (tp + fn) if (tp + fn) > 0 else 0.0
        lesion_precision = tp / (tp + fp) if (tp + fp) > 0 else 0.0
        lesion_f1 = (2 * lesion_precision * lesion_recall) / (lesion_precision + lesion_recall) if (lesion_precision + lesion_recall) > 0 else 0.0
        fp_per_case = fp / self.num_cases if self.num_cases > 0 else 0.0
        return {"lesion_wise_recall": lesion_recall, "lesion_wise_precision": lesion_precision, "lesion_wise_f1": lesion_f1,
                "voxel_wise_dsc_micro": voxel_dsc_micro, "voxel_wise_dsc_macro": voxel_dsc_macro, "fp_per_case": fp_per_case,
                "tp": tp, "fp": fp, "fn": fn,
                "dsc": voxel_dsc_micro, "recall": lesion_recall, "precision": lesion_precision}


def case_stats_device(pred_mask: torch.Tensor, target_labeled: torch.Tensor, num_target: int):
    """Everything CaseAccumulator.add needs for one (case, threshold): the prediction is labelled with min_size 0, as
    calculate_metrics does (metrics.py:376-383), so every foreground voxel belongs to a component and the voxel counts of
    the Dice terms are the component sizes."""
    num_pred, counts, mp, mt = lesion_stats_device(pred_mask, target_labeled, num_target, 0)
    return num_pred, num_target, counts, mp, mt, int(mp[1:, 0].sum()), int(mt[1:, 0].sum())


def calculate_metrics(predictions, labels, threshold=0.5, spacing=DEFAULT_SPACING):
    """metrics.py:311-404: lesion-wise recall / precision / F1, micro and macro voxel Dice, FP per case over a list (or
    batch array) of cases, with the reference's keys (incl. the dsc / recall / precision aliases)."""
    for name, obj in (("predictions", predictions), ("labels", labels)):
        if not isinstance(obj, (list, tuple)) and not (hasattr(obj, "shape") and hasattr(obj, "__getitem__")):
            raise TypeError(f"{name} must be a list/tuple or array-like object with shape and indexing support")
    pred_list = list(predictions) if isinstance(predictions, (list, tuple)) else [predictions[i] for i in range(predictions.shape[0])]
    label_list = list(labels) if isinstance(labels, (list, tuple)) else [labels[i] for i in range(labels.shape[0])]
    spacing_list = _normalize_spacing_per_case(spacing, len(pred_list))
    dev = _device()
    acc = CaseAccumulator()
    for pred, target, sp in zip(pred_list, label_list, spacing_list):
        pred, target = _squeeze_case(pred), _squeeze_case(target)
        pred_mask = _binarize(pred, threshold, dev)
        target_mask = _binarize(target, 0.5, dev)
        target_labeled, nt_d = label_device(target_mask, 0)
        acc.add(*case_stats_device(pred_mask, target_labeled, int(nt_d.item())), sp)
    return acc.result()
