"""GPU parity of the device-side training-patch pipeline (reference: datasets/patch_dataset.py) against
the patches the reference's own PatchDataset produced (tests/golden/patches.json) and against the oracle sampler:
bit-identical patches -- sampling decisions, extraction, flip, rotate, zoom + crop / pad, intensity shift, noise."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN
from oracle import augment_ref

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _golden():
    with open(os.path.join(GOLDEN, "patches.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("batch", [1, 4])
@pytest.mark.parametrize("tag", ["plain16", "aug16", "aug24", "edge48"])
def test_device_sampler_reproduces_reference_patch_dataset(tag, batch):
    from light_unet.datasets import DevicePatchSampler
    rec = _golden()[tag]
    aug = augment_ref.PATCH_AUG if rec["aug"] else None
    s = DevicePatchSampler(augment_ref.synth_cases(), tuple(rec["patch"]), 0.5, aug, rec["seed"], torch.device(DEV), noise="host")
    assert (len(s.lesion_locations), len(s.background_locations)) == (rec["n_lesion"], rec["n_background"])
    n = len(rec["items"]) // batch * batch
    k = 0
    while k < n:
        img, lab = s.sample_batch(batch)
        assert img.shape == lab.shape == (batch, 1) + tuple(rec["patch"]) and img.dtype == lab.dtype == torch.float32 and img.is_cuda
        ih, lh = img.cpu().numpy(), lab.cpu().numpy()
        for b in range(batch):
            want = rec["items"][k + b]
            assert hashlib.sha256(ih[b, 0].tobytes()).hexdigest() == want["img_sha"], (tag, k + b, s.last_decisions[b].keys(),
                                                                                         float(ih[b, 0].astype(np.float64).sum()), want["img_sum"])
            assert hashlib.sha256(lh[b, 0].tobytes()).hexdigest() == want["lab_sha"], (tag, k + b)
        k += batch


def test_each_augmentation_against_oracle_at_48():
    """One decision per kernel on 48^3 patches (the configured patch size), including rotation about every axis pair and both
    zoom directions, against oracle/augment_ref.py (== scipy.ndimage, tests/test_augment_ref.py)."""
    from light_unet.datasets import DevicePatchSampler
    vols = [(np.random.default_rng(1).random((60, 64, 70), dtype=np.float32), (np.random.default_rng(2).random((60, 64, 70)) > 0.97).astype(np.float32))]
    s = DevicePatchSampler(vols, (48, 48, 48), 0.5, None, 0, torch.device(DEV), noise="host")
    rng = np.random.default_rng(5)
    base = {"case": 0, "center": (30, 32, 35)}
    decisions = [dict(base), dict(base, flip=0), dict(base, flip=1), dict(base, flip=2), dict(base, rotate=(-15.0, (0, 1))),
                 dict(base, rotate=(7.77, (0, 2))), dict(base, rotate=(14.2, (2, 1))), dict(base, scale=0.9), dict(base, scale=1.1),
                 dict(base, scale=1.0417), dict(base, shift=-0.1), dict(base, shift=0.0731), dict(base, noise_sigma=0.01, noise=rng.normal(0, 0.01, (48, 48, 48))),
                 dict(base, center=(2, 63, 69), flip=2, rotate=(3.3, (1, 2)), scale=0.93, shift=0.05, noise_sigma=0.01, noise=rng.normal(0, 0.01, (48, 48, 48)))]
    img, lab = s.sample_batch(len(decisions), decisions=decisions)
    ih, lh = img.cpu().numpy(), lab.cpu().numpy()
    for b, d in enumerate(decisions):
        ip, lp = augment_ref.extract_patch(vols[0][0], vols[0][1], d["center"], (48, 48, 48))
        ops = {k: v for k, v in d.items() if k in ("flip", "rotate", "scale", "shift", "noise")}
        ip, lp = augment_ref.apply(ip, lp, ops, (48, 48, 48))
        assert np.array_equal(ih[b, 0], ip.astype(np.float32)), (b, list(ops), float(np.abs(ih[b, 0] - ip).max()))
        assert np.array_equal(lh[b, 0], lp.astype(np.float32)), (b, list(ops))


def test_device_noise_and_mixed_sampler():
    from light_unet.datasets import DevicePatchSampler, MixedDevicePatchSampler
    cases = augment_ref.synth_cases()
    aug = {"gaussian_noise": {"enabled": True, "prob": 1.0, "sigma": 0.05}}
    s = DevicePatchSampler(cases, (16, 16, 16), 0.5, aug, 1, torch.device(DEV))           # noise drawn on the device
    s0 = DevicePatchSampler(cases, (16, 16, 16), 0.5, None, 1, torch.device(DEV))
    b, _ = s0.sample_batch(64)
    a, _ = s.sample_batch(64, decisions=[dict(d, noise_sigma=0.05) for d in s0.last_decisions])    # same patches + device noise
    d = (a - b).double()
    inner = (b > 0.2) & (b < 0.8)                                                          # away from the clip at 0 / 1
    assert abs(float(d[inner].std()) - 0.05) < 5e-3 and abs(float(d[inner].mean())) < 5e-3
    assert float(a.min()) >= 0.0 and float(a.max()) <= 1.0
    m = MixedDevicePatchSampler(DevicePatchSampler(cases[:2], (16, 16, 16), 0.5, None, 42, torch.device(DEV)),
                                DevicePatchSampler(cases[2:], (16, 16, 16), 0.5, None, 43, torch.device(DEV)), fl_ratio=0.7, seed=42)
    for _ in range(20):
        x, y = m.sample_batch(8)
        assert x.shape == y.shape == (8, 1, 16, 16, 16)
    assert m.fl_sample_count + m.dlbcl_sample_count == 160 and 0.55 < m.fl_sample_count / 160 < 0.85
