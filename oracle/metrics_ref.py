"""Lesion-wise evaluation metrics and the validation threshold sweep (test infrastructure -- see oracle/__init__.py).

Follows /root/reference/light_unet/models/metrics.py (calculate_dsc :15-35, match_components :127-229,
calculate_lesion_metrics :232-308, calculate_metrics :311-404) and the selection logic of Trainer.validate
(/root/reference/light_unet/core/trainer.py:183-189, :423-445).  Labelling goes through the oracle's own 6-connectivity
restatement (bbox_ref.connected_components == scipy.ndimage.label); centres of mass are coordinate means in float64
(scipy.ndimage.center_of_mass with unit weights).  Pinned against the reference itself by tests/golden/make_golden.py.
"""
from __future__ import annotations

import numpy as np

from . import bbox_ref

SMOOTH = 1e-6
DEFAULT_SPACING = (4.0, 4.0, 4.0)
EPS = 1e-8                       # Trainer.EPS (trainer.py)


def calculate_dsc(pred, target, smooth=SMOOTH):
    pred, target = np.ravel(pred), np.ravel(target)
    return (2.0 * (pred * target).sum() + smooth) / (pred.sum() + target.sum() + smooth)


def component_centers(labeled):
    n = int(labeled.max()) if labeled.size else 0
    if n == 0:
        return np.empty((0, 3), dtype=np.float64)
    flat = labeled.ravel().astype(np.int64)
    cnt = np.bincount(flat, minlength=n + 1).astype(np.float64)
    zz, yy, xx = np.meshgrid(*[np.arange(s, dtype=np.float64) for s in labeled.shape], indexing="ij")
    cols = [np.bincount(flat, weights=c.ravel(), minlength=n + 1) / np.where(cnt > 0, cnt, 1.0) for c in (zz, yy, xx)]
    return np.stack(cols, axis=1)[1:]


def match_components(pred_labeled, target_labeled, iou_threshold=0.1, distance_threshold_mm=10.0, spacing=DEFAULT_SPACING):
    num_pred, num_target = int(pred_labeled.max()), int(target_labeled.max())
    if num_pred == 0 or num_target == 0:
        return [], list(range(1, num_pred + 1)), list(range(1, num_target + 1))
    pf, tf = np.ravel(pred_labeled).astype(np.int64), np.ravel(target_labeled).astype(np.int64)
    off = np.int64(num_target + 1)
    inter = np.bincount(pf * off + tf, minlength=(num_pred + 1) * off).reshape(num_pred + 1, num_target + 1)
    inter[0, :] = 0
    inter[:, 0] = 0
    ps, ts = np.bincount(pf, minlength=num_pred + 1), np.bincount(tf, minlength=num_target + 1)
    union = ps[:, None] + ts[None, :] - inter
    iou = np.divide(inter, union, out=np.zeros_like(inter, dtype=np.float32), where=union > 0)
    sp = np.asarray(spacing, dtype=np.float64)
    pc, tcn = component_centers(pred_labeled) * sp, component_centers(target_labeled) * sp
    dist = np.linalg.norm(pc[:, None, :] - tcn[None, :, :], axis=2)
    matches, matched_pred = [], set()
    taken = np.zeros(num_target, dtype=bool)
    for pid in range(1, num_pred + 1):
        row = iou[pid, 1:]
        valid = ~taken & ((row >= iou_threshold) | (dist[pid - 1] <= distance_threshold_mm))
        if not np.any(valid):
            continue
        best = int(np.argmax(np.where(valid, row, -np.inf)))
        matches.append((pid, best + 1))
        matched_pred.add(pid)
        taken[best] = True
    return (matches, [i for i in range(1, num_pred + 1) if i not in matched_pred],
            [i for i in range(1, num_target + 1) if not taken[i - 1]])


def calculate_lesion_metrics(pred, target, threshold=0.5, min_size_voxels=0, iou_threshold=0.1, distance_threshold_mm=10.0,
                             spacing=DEFAULT_SPACING):
    pb = (pred >= threshold).astype(np.int32)
    tb = (target >= 0.5).astype(np.int32)
    pl, npred = bbox_ref.connected_components(pb, min_size_voxels)
    tl, ntar = bbox_ref.connected_components(tb, min_size_voxels)
    if ntar == 0:
        if npred == 0:
            return {"recall": 1.0, "precision": 1.0, "f1": 1.0, "tp": 0, "fp": 0, "fn": 0}
        return {"recall": 0.0, "precision": 0.0, "f1": 0.0, "tp": 0, "fp": npred, "fn": 0}
    if npred == 0:
        return {"recall": 0.0, "precision": 0.0, "f1": 0.0, "tp": 0, "fp": 0, "fn": ntar}
    m, up, ut = match_components(pl, tl, iou_threshold, distance_threshold_mm, spacing)
    tp, fp, fn = len(m), len(up), len(ut)
    recall = tp / (tp + fn) if (tp + fn) > 0 else 0.0
    precision = tp / (tp + fp) if (tp + fp) > 0 else 0.0
    f1 = 2 * (precision * recall) / (precision + recall) if (precision + recall) > 0 else 0.0
    return {"recall": recall, "precision": precision, "f1": f1, "tp": tp, "fp": fp, "fn": fn}


def calculate_metrics(predictions, labels, threshold=0.5, spacing=DEFAULT_SPACING):
    n = len(predictions)
    spacings = [tuple(map(float, s)) for s in spacing] if (len(spacing) == n and isinstance(spacing[0], (list, tuple, np.ndarray))) \
        else [tuple(map(float, spacing))] * n
    tp = fp = fn = 0
    inter_sum = union_sum = 0.0
    dscs = []
    for pred, target, sp in zip(predictions, labels, spacings):
        pred, target = np.asarray(pred), np.asarray(target)
        pb, tb = (pred >= threshold).astype(np.int32), (target >= 0.5).astype(np.int32)
        inter_sum += (pb * tb).sum()
        union_sum += pb.sum() + tb.sum()
        dscs.append(calculate_dsc(pb, tb))
        m = calculate_lesion_metrics(pred, target, threshold, 0, 0.1, 10.0, sp)
        tp += m["tp"]; fp += m["fp"]; fn += m["fn"]
    micro = (2.0 * inter_sum + SMOOTH) / (union_sum + SMOOTH)
    macro = np.mean(dscs) if dscs else 0.0
    rec = tp / (tp + fn) if (tp + fn) > 0 else 0.0
    prec = tp / (tp + fp) if (tp + fp) > 0 else 0.0
    f1 = (2 * prec * rec) / (prec + rec) if (prec + rec) > 0 else 0.0
    return {"lesion_wise_recall": rec, "lesion_wise_precision": prec, "lesion_wise_f1": f1, "voxel_wise_dsc_micro": micro,
            "voxel_wise_dsc_macro": macro, "fp_per_case": fp / n if n > 0 else 0.0, "tp": tp, "fp": fp, "fn": fn,
            "dsc": micro, "recall": rec, "precision": prec}


def is_better_metric(recall, dsc, best_recall, best_dsc, tie_threshold):
    """trainer.py:183-189."""
    if recall > best_recall + EPS:
        return True
    return abs(recall - best_recall) <= tie_threshold + EPS and dsc > best_dsc + EPS


def select_threshold(predictions, labels, spacings, thresholds, tie_threshold=0.0):
    """The threshold sweep of Trainer.validate (trainer.py:423-445)."""
    best_t = thresholds[0]
    best = calculate_metrics(predictions, labels, best_t, spacings)
    best_recall, best_dsc = best["lesion_wise_recall"], best["voxel_wise_dsc_macro"]
    for t in thresholds[1:]:
        m = calculate_metrics(predictions, labels, t, spacings)
        if is_better_metric(m["lesion_wise_recall"], m["voxel_wise_dsc_macro"], best_recall, best_dsc, tie_threshold):
            best_recall, best_dsc, best_t, best = m["lesion_wise_recall"], m["voxel_wise_dsc_macro"], t, m
    best = dict(best)
    best["best_threshold"], best["best_recall"], best["best_dsc_macro"] = best_t, best_recall, best_dsc
    return best


def synth_case(shape, seed):
    """Seeded synthetic (probability map, label) pair: the label is a set of blobs, the prediction finds most of them
    slightly shifted / shrunk, misses some and adds false positives -- every branch of the matcher is exercised."""
    rng = np.random.default_rng(seed)
    zz, yy, xx = np.meshgrid(*[np.arange(s, dtype=np.float32) for s in shape], indexing="ij")
    label = np.zeros(shape, dtype=np.float32)
    prob = (0.2 * rng.random(shape, dtype=np.float32)).astype(np.float32)
    for i in range(int(rng.integers(3, 9))):
        c = np.array([rng.uniform(2, s - 3) for s in shape])
        r = rng.uniform(1.2, 4.0)
        label[((zz - c[0]) ** 2 + (yy - c[1]) ** 2 + (xx - c[2]) ** 2) <= r * r] = 1.0
        kind = rng.integers(0, 4)
        if kind == 0:
            continue                                              # missed lesion
        shift = rng.uniform(-1.5, 1.5, size=3) if kind < 3 else rng.uniform(-6.0, 6.0, size=3)
        rr = r * rng.uniform(0.6, 1.2)
        d2 = (zz - c[0] - shift[0]) ** 2 + (yy - c[1] - shift[1]) ** 2 + (xx - c[2] - shift[2]) ** 2
        prob = np.maximum(prob, (rng.uniform(0.35, 0.95) * np.exp(-d2 / (2 * rr * rr))).astype(np.float32))
    for i in range(int(rng.integers(0, 4))):                      # false positives
        c = np.array([rng.uniform(0, s - 1) for s in shape])
        d2 = (zz - c[0]) ** 2 + (yy - c[1]) ** 2 + (xx - c[2]) ** 2
        prob = np.maximum(prob, (rng.uniform(0.3, 0.9) * (d2 <= rng.uniform(1.0, 9.0))).astype(np.float32))
    return prob.astype(np.float32), label
