"""Device-resident validation -- Trainer.validate (trainer.py:349-445) without its host round trips.

The reference pulls every validation volume to the host, runs the per-window sliding window (H2D + ~100 launches +
blocking D2H + NumPy read-modify-write per window, utils.py:86-134), keeps all probability maps and labels in host
lists and then recomputes scipy labelling + bincount matching once per threshold of the sensitivity sweep
(trainer.py:423-439 -> metrics.py:311-404).  Here a case stays on the GPU from the volume upload to the integer
statistics: batched sliding window + Gaussian stitch (body mask fused), the label volume labelled ONCE, and per threshold
of the sweep threshold -> 6-connected labelling -> one pass of pair statistics; only (n_pred x n_target)-sized integer
tables come back, and the reference's own arithmetic on them (models/metrics.py) gives bit-identical metrics for
identical probability maps.  Nothing per-case is kept: the sweep is accumulated case by case, so validation memory
does not grow with the size of the validation split.
"""
from __future__ import annotations

from typing import Iterable, Optional, Sequence

import numpy as np
import torch

from .. import _native as nv
from ..models.metrics import (DEFAULT_SPACING, CaseAccumulator, _binarize, _threshold_device, _vol3, case_stats_device,
                              label_device)
from ..utils import sliding_window_device

EPS = 1e-8      # Trainer.EPS


def is_better_metric(recall, dsc, best_recall, best_dsc, tie_threshold):
    """trainer.py:183-189: higher lesion recall wins; within the tie margin the higher macro Dice does.
    Returns (is_better, recall_improved)."""
    tie_margin = tie_threshold + EPS
    if recall > best_recall + EPS:
        return True, True
    if abs(recall - best_recall) <= tie_margin and dsc > best_dsc + EPS:
        return True, False
    return False, False


def resolve_spacing_value(spacings, index, target_spacing):
    """trainer.py:194-206."""
    if spacings is None:
        value = target_spacing
    elif isinstance(spacings, (torch.Tensor, list, tuple)):
        value = target_spacing if len(spacings) <= index else spacings[index]
    else:
        value = spacings
    if isinstance(value, torch.Tensor):
        value = value.tolist()
    return tuple(float(s) for s in value)


class DeviceValidator:
    def __init__(self, model, config: dict, device: Optional[torch.device] = None):
        if not torch.cuda.is_available():
            raise nv.NativeError("DeviceValidator: the B200-native path needs a CUDA device (no CPU fallback)")
        self.model, self.config = model, config
        self.device = torch.device(device) if device is not None else next(model.parameters()).device
        v = config["validation"]
        self.default_threshold = v["default_threshold"]
        self.thresholds = list(v.get("threshold_sensitivity_range", [self.default_threshold]))
        self.tie_threshold = config.get("metrics", {}).get("model_selection", {}).get("tie_threshold", 0.0)
        self.patch_size = tuple(config["data"]["patch_size"])
        self.target_spacing = tuple(config.get("data", {}).get("spacing", {}).get("target", DEFAULT_SPACING))
        bm = config.get("data", {}).get("body_mask", {})
        self.apply_body_mask = bm.get("apply_to_validation", False) and bm.get("enabled", False)
        self.reset()

    def reset(self):
        self.acc = {t: CaseAccumulator() for t in self.thresholds}
        self.num_cases = 0

    # ------------------------------------------------------------------ per case
    def add_probability_map(self, prob, label, spacing=None):
        """One case whose probability map is already known (host array or CUDA tensor): the threshold sweep only."""
        dev = self.device
        with torch.cuda.device(dev):
            if isinstance(prob, torch.Tensor):
                prob_d = _vol3(prob.to(device=dev, dtype=torch.float32)).contiguous()
            else:
                prob_d = _vol3(torch.from_numpy(np.ascontiguousarray(prob, dtype=np.float32)).to(dev))
            target_mask = _binarize(label, 0.5, dev)                      # metrics.py:264 / :364
            if tuple(target_mask.shape) != tuple(prob_d.shape):
                raise ValueError(f"label shape {tuple(target_mask.shape)} does not match the probability map {tuple(prob_d.shape)}")
            target_labeled, nt_d = label_device(target_mask, 0)
            num_target = int(nt_d.item())
            sp = tuple(float(s) for s in (spacing if spacing is not None else self.target_spacing))
            for t in self.thresholds:
                self.acc[t].add(*case_stats_device(_threshold_device(prob_d, t), target_labeled, num_target), sp)
            self.num_cases += 1

    @torch.no_grad()
    def add_case(self, image, label, spacing=None, body_mask=None):
        """One validation volume: sliding window on the device (trainer.py:392-399), body mask (:401-402), sweep."""
        dev = self.device
        vol = image if isinstance(image, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(image, dtype=np.float32))
        vol = _vol3(vol.to(device=dev, dtype=torch.float32, non_blocking=True))
        bm = None
        if self.apply_body_mask and body_mask is not None:
            bm = body_mask if isinstance(body_mask, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(body_mask))
            bm = _vol3(bm != 0)
        prob_d, _ = sliding_window_device(vol, self.model, self.patch_size, 0.5, True, body_mask=bm)
        self.add_probability_map(prob_d, label, spacing)
        return prob_d

    # ------------------------------------------------------------------ result
    def result(self):
        """(0.0, best_metrics) exactly as Trainer.validate returns them (trainer.py:415-445)."""
        if self.num_cases == 0:
            return 0.0, {"lesion_wise_recall": 0.0, "lesion_wise_precision": 0.0, "voxel_wise_dsc_macro": 0.0,
                         "voxel_wise_dsc_micro": 0.0, "fp_per_case": 0.0, "best_threshold": self.default_threshold,
                         "best_recall": 0.0, "best_dsc_macro": 0.0}
        best_threshold = self.thresholds[0]
        best_metrics = self.acc[best_threshold].result()
        best_recall, best_dsc = best_metrics["lesion_wise_recall"], best_metrics["voxel_wise_dsc_macro"]
        for t in self.thresholds[1:]:
            m = self.acc[t].result()
            better, _ = is_better_metric(m["lesion_wise_recall"], m["voxel_wise_dsc_macro"], best_recall, best_dsc, self.tie_threshold)
            if better:
                best_recall, best_dsc, best_threshold, best_metrics = m["lesion_wise_recall"], m["voxel_wise_dsc_macro"], t, m
        best_metrics = dict(best_metrics)
        best_metrics["best_threshold"] = best_threshold
        best_metrics["best_recall"] = best_recall
        best_metrics["best_dsc_macro"] = best_dsc
        return 0.0, best_metrics

    def sweep(self):
        """{threshold: metrics} of every threshold of the sensitivity range (what scripts/evaluate-style reports print)."""
        return {t: self.acc[t].result() for t in self.thresholds}

    # ------------------------------------------------------------------ loader level
    def validate(self, val_loader: Iterable):
        """Drop-in body of Trainer.validate: consumes the reference's validation batches -- (images, labels[, case_ids,
        spacings[, body_masks]]) with images [B, 1, D, H, W] or [B, D, H, W] (trainer.py:363-404)."""
        self.model.eval()
        self.reset()
        for batch in val_loader:
            spacings = body_masks = None
            if isinstance(batch, (list, tuple)) and len(batch) >= 5:
                images, labels, _, spacings, body_masks = batch[:5]
            elif isinstance(batch, (list, tuple)) and len(batch) >= 4:
                images, labels, _, spacings = batch[:4]
            else:
                images, labels = batch[0], batch[1]
            for b in range(images.shape[0]):
                if images.ndim == 5:
                    image, label = images[b, 0], labels[b, 0]
                    bmask = body_masks[b, 0] if body_masks is not None else None
                elif images.ndim == 4:
                    image, label = images[b], labels[b]
                    bmask = body_masks[b] if body_masks is not None else None
                else:
                    raise ValueError(f"Unexpected image shape: {images.shape}")
                sp = resolve_spacing_value(spacings, b, self.target_spacing) if spacings is not None else self.target_spacing
                self.add_case(image, label, sp, bmask)
        return self.result()


def use_device_validation(trainer):
    """Swap the reference Trainer's validate() (trainer.py:349-445) for the device-resident sweep, in place.

        trainer = Trainer(config)                      # the reference's own class (through l3d_overlay)
        use_device_validation(trainer)
        trainer.train()                                # every validation epoch now stays on the GPU

    The replacement keeps the contract -- validate(epoch) -> (0.0, metrics dict with best_threshold / best_recall /
    best_dsc_macro) -- and reads the same attributes (model, config, val_loader, device)."""
    validator = DeviceValidator(trainer.model, trainer.config, getattr(trainer, "device", None))

    def validate(epoch):
        return validator.validate(trainer.val_loader)
    trainer.validate = validate
    return validator
