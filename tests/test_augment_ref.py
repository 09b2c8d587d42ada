"""Pins oracle/augment_ref.py (the numpy restatement of the reference's patch extraction / augmentation,
patch_dataset.py:136-220) against scipy.ndimage itself -- the third-party arithmetic the reference calls -- bit for bit."""
import numpy as np
import pytest

from oracle import augment_ref

ndimage = pytest.importorskip("scipy.ndimage")


@pytest.mark.parametrize("shape", [(12, 14, 16), (48, 48, 48), (9, 9, 9)])
@pytest.mark.parametrize("axes", [(0, 1), (0, 2), (1, 2), (2, 0)])
def test_rotate_matches_scipy(shape, axes):
    rng = np.random.default_rng(hash((shape, axes)) % 1000)
    img = rng.random(shape, dtype=np.float32)
    lab = (rng.random(shape) > 0.7).astype(np.float32)
    for ang in (-15.0, -7.3, 0.0, 3.141, 14.99, 90.0, 45.0):
        want_i = ndimage.rotate(img, ang, axes=axes, reshape=False, order=1, mode="constant", cval=0)
        want_l = ndimage.rotate(lab, ang, axes=axes, reshape=False, order=0, mode="constant", cval=0)
        assert np.array_equal(augment_ref.rotate(img, ang, axes, 1), want_i), (shape, axes, ang)
        assert np.array_equal(augment_ref.rotate(lab, ang, axes, 0), want_l), (shape, axes, ang)


@pytest.mark.parametrize("shape", [(12, 14, 16), (48, 48, 48), (7, 20, 11)])
def test_zoom_matches_scipy(shape):
    rng = np.random.default_rng(sum(shape))
    img = rng.random(shape, dtype=np.float32)
    lab = (rng.random(shape) > 0.7).astype(np.float32)
    for f in (0.9, 0.95, 1.0, 1.0417, 1.1, 0.9001, 1.0999):
        want_i = ndimage.zoom(img, f, order=1, mode="constant", cval=0)
        want_l = ndimage.zoom(lab, f, order=0, mode="constant", cval=0)
        got_i, got_l = augment_ref.zoom(img, f, 1), augment_ref.zoom(lab, f, 0)
        assert got_i.shape == want_i.shape and np.array_equal(got_i, want_i), (shape, f)
        assert np.array_equal(got_l, want_l), (shape, f)
        p = (12, 12, 12)
        assert augment_ref.fit_to_patch(got_i, p).shape == p


def test_extract_patch_edges():
    img = np.arange(10 * 11 * 12, dtype=np.float32).reshape(10, 11, 12)
    lab = (img % 7 == 0).astype(np.float32)
    for c in [(0, 0, 0), (9, 10, 11), (5, 5, 5), (2, 9, 1)]:
        ip, lp = augment_ref.extract_patch(img, lab, c, (8, 8, 8))
        assert ip.shape == lp.shape == (8, 8, 8)
        zs, ys, xs = (max(0, c[i] - 4) for i in range(3))
        blk = img[zs:zs + 8, ys:ys + 8, xs:xs + 8]
        assert np.array_equal(ip[:blk.shape[0], :blk.shape[1], :blk.shape[2]], blk)
        assert ip[blk.shape[0]:].sum() == 0 and ip[:, blk.shape[1]:].sum() == 0 and ip[:, :, blk.shape[2]:].sum() == 0


def _golden():
    import json
    import os
    from helpers import GOLDEN
    with open(os.path.join(GOLDEN, "patches.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("tag", ["plain16", "aug16", "aug24", "edge48"])
def test_oracle_sampler_reproduces_reference_patch_dataset(tag):
    """oracle/augment_ref.RefSampler against the patches the reference's own PatchDataset produced from the same in-memory
    volumes (tests/golden/patches.json, written by make_golden.py): same candidate lists, same item sequence, and every
    patch bit-identical (SHA-256 of the float32 bytes) -- sampling, extraction and all five augmentations."""
    import hashlib
    rec = _golden()[tag]
    s = augment_ref.RefSampler(augment_ref.synth_cases(), tuple(rec["patch"]), 0.5, augment_ref.PATCH_AUG if rec["aug"] else None, rec["seed"])
    assert (len(s.lesion), len(s.background)) == (rec["n_lesion"], rec["n_background"])
    seen = set()
    for k, want in enumerate(rec["items"]):
        ip, lp, ops = s.item()
        seen |= set(ops)
        assert hashlib.sha256(ip.tobytes()).hexdigest() == want["img_sha"], (tag, k, ops.keys(), float(ip.astype(np.float64).sum()), want["img_sum"])
        assert hashlib.sha256(lp.tobytes()).hexdigest() == want["lab_sha"], (tag, k)
    if rec["aug"] and len(rec["items"]) >= 24:
        assert seen == {"flip", "rotate", "scale", "shift", "noise"}
