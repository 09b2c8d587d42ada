"""PatchDataset on the device -- class-balanced 48^3 patch sampling + augmentation with the volumes resident in HBM.

Reference: light_unet/datasets/patch_dataset.py.  Its __getitem__ (:114-134) re-loads two whole NIfTI volumes from disk
for every patch, slices one window out of them and runs scipy.ndimage.rotate / zoom on the host (:156-220); at B200 step
rates (~4.6 ms per batch of 8) that loader is the bottleneck.  Here the volumes are uploaded ONCE, a whole batch is cut by
one kernel and augmented by at most four more (libl3d: l3d_patch_extract / flip / rotate / zoom / intensity).

What is kept exactly:
  * the sampling logic: candidate lesion / background locations (:72-99) and the per-item decisions (:115-125, :160-218)
    are drawn in the reference's order from generators seeded like the reference seeds the global ones
    (np.random.seed(seed) / random.seed(seed)), so with num_workers = 0 the same seed visits the same (case, centre,
    flip, angle, scale, shift) sequence;
  * the arithmetic: every patch is bit-identical to the reference's for the same decisions (float64 interpolation in
    scipy's operation order).  Gaussian noise is drawn on the device by default (a different random stream, same
    distribution); noise="host" draws it with the reference's generator for bit-exact comparison.
"""
from __future__ import annotations

import math
import random as _random
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from .. import _native as nv


def _cos_sin_degrees(angle: float) -> Tuple[float, float]:
    """scipy.ndimage.rotate takes cos / sin from scipy.special.cosdg / sindg (cephes); use them when scipy is there so the
    last bit agrees, math otherwise."""
    try:
        from scipy import special
        return float(special.cosdg(angle)), float(special.sindg(angle))
    except ImportError:                                           # pragma: no cover
        return math.cos(math.radians(angle)), math.sin(math.radians(angle))


class DevicePatchSampler:
    def __init__(self, volumes: Sequence, patch_size=(48, 48, 48), lesion_patch_ratio=0.5, augmentation: Optional[dict] = None,
                 seed: int = 42, device=None, noise: str = "device"):
        """volumes: sequence of (image, label) or (image, label, body_mask) host arrays [D, H, W] (what nib.load(...)
        .get_fdata() returns for the case files, patch_dataset.py:76-77,128-129).  body_mask restricts the background
        candidates (:88-91)."""
        if not torch.cuda.is_available():
            raise nv.NativeError("DevicePatchSampler: the B200-native path needs a CUDA device (no CPU fallback)")
        if noise not in ("device", "host"):
            raise ValueError("noise must be 'device' or 'host'")
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.patch_size = tuple(int(p) for p in patch_size)
        self.lesion_patch_ratio = lesion_patch_ratio
        self.augmentation = augmentation
        self.noise_mode = noise
        self.np_rng = np.random.RandomState(seed)       # the stream np.random.seed(seed) gives the reference's global generator
        self.py_rng = _random.Random(seed)              # likewise random.seed(seed)
        self.images, self.labels, self.dims = [], [], []
        self.lesion_locations, self.background_locations = [], []
        for case_idx, vol in enumerate(volumes):
            image, label = np.asarray(vol[0]), np.asarray(vol[1])
            body = np.asarray(vol[2]).astype(bool) if len(vol) > 2 and vol[2] is not None else None
            if image.ndim != 3 or image.shape != label.shape:
                raise ValueError(f"case {case_idx}: image {image.shape} / label {label.shape} must be equal 3-D shapes")
            # candidate locations, patch_dataset.py:79-97 (same draws, same order)
            lesion_coords = np.argwhere(label > 0)
            if len(lesion_coords) > 0:
                for idx in self.np_rng.randint(len(lesion_coords), size=max(10, len(lesion_coords) // 1000)):
                    self.lesion_locations.append((case_idx, lesion_coords[idx]))
            bg_coords = np.argwhere((label == 0) & body) if body is not None else np.argwhere(label == 0)
            if len(bg_coords) > 0:
                for idx in self.np_rng.randint(len(bg_coords), size=max(10, len(bg_coords) // 5000)):
                    self.background_locations.append((case_idx, bg_coords[idx]))
            self.images.append(torch.from_numpy(np.ascontiguousarray(image, dtype=np.float32)).to(self.device))
            self.labels.append(torch.from_numpy(np.ascontiguousarray(label, dtype=np.float32)).to(self.device))
            self.dims.append(image.shape)
        self._img_ptrs = torch.tensor([t.data_ptr() for t in self.images], dtype=torch.int64)
        self._lab_ptrs = torch.tensor([t.data_ptr() for t in self.labels], dtype=torch.int64)
        self.last_decisions: List[dict] = []

    def __len__(self):
        return len(self.lesion_locations) + len(self.background_locations)

    # ------------------------------------------------------------------ host: the reference's random decisions
    def _draw_item(self) -> dict:
        """One __getitem__ worth of decisions (patch_dataset.py:115-125 then :160-218), in the reference's draw order."""
        rs, pr = self.np_rng, self.py_rng
        if rs.rand() < self.lesion_patch_ratio and len(self.lesion_locations) > 0:
            case_idx, center = self.lesion_locations[rs.randint(len(self.lesion_locations))]
        elif len(self.background_locations) > 0:
            case_idx, center = self.background_locations[rs.randint(len(self.background_locations))]
        else:
            case_idx, center = self.lesion_locations[rs.randint(len(self.lesion_locations))]
        d = {"case": int(case_idx), "center": tuple(int(c) for c in center)}
        aug = self.augmentation
        if not aug:
            return d
        if aug.get("random_flip", {}).get("enabled", False):
            if rs.rand() < aug["random_flip"].get("prob", 0.5):
                d["flip"] = pr.choice(aug["random_flip"].get("axes", [0, 1, 2]))
        if aug.get("random_rotation", {}).get("enabled", False):
            if rs.rand() < aug["random_rotation"].get("prob", 0.5):
                lo, hi = aug["random_rotation"].get("angle_range", [-15, 15])
                angle = rs.uniform(lo, hi)
                d["rotate"] = (float(angle), tuple(pr.choice(aug["random_rotation"].get("axes", [[0, 1], [0, 2], [1, 2]]))))
        if aug.get("random_scale", {}).get("enabled", False):
            if rs.rand() < aug["random_scale"].get("prob", 0.3):
                lo, hi = aug["random_scale"].get("scale_range", [0.9, 1.1])
                d["scale"] = float(rs.uniform(lo, hi))
        if aug.get("intensity_shift", {}).get("enabled", False):
            if rs.rand() < aug["intensity_shift"].get("prob", 0.5):
                lo, hi = aug["intensity_shift"].get("shift_range", [-0.1, 0.1])
                d["shift"] = float(rs.uniform(lo, hi))
        if aug.get("gaussian_noise", {}).get("enabled", False):
            if rs.rand() < aug["gaussian_noise"].get("prob", 0.3):
                d["noise_sigma"] = float(aug["gaussian_noise"].get("sigma", 0.01))
                if self.noise_mode == "host":
                    d["noise"] = rs.normal(0, d["noise_sigma"], self.patch_size)
        return d

    # ------------------------------------------------------------------ device: one batch
    @torch.no_grad()
    def sample_batch(self, batch_size: int, decisions: Optional[List[dict]] = None):
        """-> (images, labels), each a CUDA float32 tensor [B, 1, pd, ph, pw] (what the DataLoader collates from
        PatchDataset.__getitem__ followed by trainer.py:225's images.float())."""
        dev = self.device
        pd, ph, pw = self.patch_size
        per = pd * ph * pw
        B = int(batch_size)
        items = decisions if decisions is not None else [self._draw_item() for _ in range(B)]
        self.last_decisions = items
        st = nv.stream_ptr(dev)
        with torch.cuda.device(dev):
            pt = torch.tensor([[self._img_ptrs[d["case"]], self._lab_ptrs[d["case"]]] for d in items], dtype=torch.int64)
            dims = torch.tensor([self.dims[d["case"]] for d in items], dtype=torch.int32)
            start = torch.tensor([[max(0, c - p // 2) for c, p in zip(d["center"], self.patch_size)] for d in items], dtype=torch.int32)
            img_ptrs, lab_ptrs = pt[:, 0].contiguous().to(dev), pt[:, 1].contiguous().to(dev)
            dims_d, start_d = dims.to(dev), start.to(dev)
            img = torch.empty(B, 1, pd, ph, pw, dtype=torch.float32, device=dev)
            lab = torch.empty_like(img)
            nv.call("l3d_patch_extract", nv.ptr(img_ptrs), nv.ptr(lab_ptrs), nv.ptr(dims_d), nv.ptr(start_d), B, pd, ph, pw,
                    nv.ptr(img), nv.ptr(lab), st, algo_bytes=16 * B * per)
            if any("flip" in d for d in items):
                axis = torch.tensor([d.get("flip", -1) for d in items], dtype=torch.int32).to(dev)
                img2, lab2 = torch.empty_like(img), torch.empty_like(lab)
                nv.call("l3d_patch_flip", nv.ptr(img), nv.ptr(lab), nv.ptr(img2), nv.ptr(lab2), nv.ptr(axis), B, pd, ph, pw, st,
                        algo_bytes=16 * B * per)
                img, lab = img2, lab2
            if any("rotate" in d for d in items):
                axes, coef = [], []
                for d in items:
                    if "rotate" not in d:
                        axes.append((-1, -1)); coef.append((0.0,) * 6)
                        continue
                    angle, ax = d["rotate"]
                    a0, a1 = sorted(int(a) % 3 for a in ax)
                    c, s = _cos_sin_degrees(angle)
                    m = np.array([[c, s], [-s, c]])
                    shp = np.array([self.patch_size[a0], self.patch_size[a1]], dtype=np.float64)
                    off = (shp - 1) / 2 - m @ ((shp - 1) / 2)         # scipy: offset = in_center - rot_matrix @ out_center
                    axes.append((a0, a1)); coef.append((m[0, 0], m[0, 1], m[1, 0], m[1, 1], off[0], off[1]))
                axes_d = torch.tensor(axes, dtype=torch.int32).to(dev)
                coef_d = torch.tensor(coef, dtype=torch.float64).to(dev)
                img2, lab2 = torch.empty_like(img), torch.empty_like(lab)
                nv.call("l3d_patch_rotate", nv.ptr(img), nv.ptr(lab), nv.ptr(img2), nv.ptr(lab2), nv.ptr(axes_d), nv.ptr(coef_d), B, pd, ph, pw, st,
                        algo_bytes=16 * B * per)
                img, lab = img2, lab2
            if any("scale" in d for d in items):
                geo, zf = [], []
                for d in items:
                    if "scale" not in d:
                        geo.append((0,) * 7); zf.append((1.0, 1.0, 1.0))
                        continue
                    zd = [int(round(n * d["scale"])) for n in self.patch_size]         # scipy.ndimage.zoom's output shape
                    geo.append((1, *zd, *[(z - p) // 2 if z > p else 0 for z, p in zip(zd, self.patch_size)]))
                    zf.append(tuple((n - 1) / (z - 1) if z > 1 else 1.0 for n, z in zip(self.patch_size, zd)))
                geo_d = torch.tensor(geo, dtype=torch.int32).to(dev)
                zf_d = torch.tensor(zf, dtype=torch.float64).to(dev)
                img2, lab2 = torch.empty_like(img), torch.empty_like(lab)
                nv.call("l3d_patch_zoom", nv.ptr(img), nv.ptr(lab), nv.ptr(img2), nv.ptr(lab2), nv.ptr(geo_d), nv.ptr(zf_d), B, pd, ph, pw, st,
                        algo_bytes=16 * B * per)
                img, lab = img2, lab2
            has_shift, has_noise = any("shift" in d for d in items), any("noise_sigma" in d for d in items)
            if has_shift or has_noise:
                shift = torch.tensor([np.float32(d.get("shift", 0.0)) for d in items], dtype=torch.float32).to(dev)
                shift_on = torch.tensor([1 if "shift" in d else 0 for d in items], dtype=torch.int32).to(dev)
                noise_d = noise_on = None
                if has_noise:
                    noise_on = torch.tensor([1 if "noise_sigma" in d else 0 for d in items], dtype=torch.int32).to(dev)
                    if self.noise_mode == "host":
                        host = np.zeros((B, pd, ph, pw), dtype=np.float64)
                        for i, d in enumerate(items):
                            if "noise" in d:
                                host[i] = d["noise"]
                        noise_d = torch.from_numpy(host).to(dev)
                    else:
                        sig = torch.tensor([d.get("noise_sigma", 0.0) for d in items], dtype=torch.float64, device=dev)
                        noise_d = torch.randn(B, pd, ph, pw, dtype=torch.float64, device=dev) * sig.view(B, 1, 1, 1)
                nv.call("l3d_patch_intensity", nv.ptr(img), nv.ptr(shift), nv.ptr(shift_on), nv.ptr(noise_d), nv.ptr(noise_on), B, per, st,
                        algo_bytes=8 * B * per)
        return img, lab


class MixedDevicePatchSampler:
    """MixedPatchDataset (patch_dataset.py:223-268): every item comes from the FL sampler with probability fl_ratio, from the
    DLBCL sampler otherwise.  The two samplers keep their own generators (seed, seed + 1 in the reference); the mixing draws
    come from a third stream seeded like the FL dataset's, which is what the shared global generator amounts to for
    num_workers = 0 only approximately -- the per-domain sample counts are what the reference reports (:262-268)."""

    def __init__(self, fl: DevicePatchSampler, dlbcl: DevicePatchSampler, fl_ratio: float = 0.5, seed: int = 42):
        self.fl, self.dlbcl, self.fl_ratio = fl, dlbcl, fl_ratio
        self.rng = np.random.RandomState(seed)
        self.reset_sample_counts()

    def reset_sample_counts(self):
        self.fl_sample_count = 0
        self.dlbcl_sample_count = 0

    def __len__(self):
        return len(self.fl) + len(self.dlbcl)

    def sample_batch(self, batch_size: int):
        pick_fl = [(self.rng.rand() < self.fl_ratio and len(self.fl) > 0) or len(self.dlbcl) == 0 for _ in range(batch_size)]
        n_fl = sum(pick_fl)
        self.fl_sample_count += n_fl
        self.dlbcl_sample_count += batch_size - n_fl
        parts = []
        if n_fl:
            parts.append((self.fl.sample_batch(n_fl), [i for i, p in enumerate(pick_fl) if p]))
        if batch_size - n_fl:
            parts.append((self.dlbcl.sample_batch(batch_size - n_fl), [i for i, p in enumerate(pick_fl) if not p]))
        img = torch.empty(batch_size, 1, *self.fl.patch_size, dtype=torch.float32, device=self.fl.device)
        lab = torch.empty_like(img)
        for (pi, pl), idx in parts:
            ix = torch.tensor(idx, device=img.device)
            img[ix], lab[ix] = pi, pl
        return img, lab
