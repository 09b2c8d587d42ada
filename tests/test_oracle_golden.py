"""CPU tests: the oracle restatement reproduces the fixtures generated from the
reference (tests/golden/make_golden.py).  This is what pins the oracle."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import bbox_ref, loss_ref, stitch_ref, synth, unet_ref
from helpers import GOLDEN, UNET_CASES, load_unet_case, sub


@pytest.mark.parametrize("name", UNET_CASES)
def test_unet_forward_matches_reference_fixture(name):
    z, meta, cfg, sd_np, x, t = load_unet_case(name)
    assert unet_ref.count_parameters(cfg) == meta["n_params"]
    sd = unet_ref.to_torch(sd_np)
    with torch.no_grad():
        y = unet_ref.forward(sd, torch.from_numpy(x), cfg).numpy()
    s = meta["subsample"]
    ref = z["prob_eval_sub"] if s else z["prob_eval"]
    assert np.abs(sub(y, s) - ref).max() < 5e-6
    loss = loss_ref.focal_tversky(torch.from_numpy(y), torch.from_numpy(t)).item()
    assert abs(loss - float(z["loss_eval"])) < 1e-5


@pytest.mark.parametrize("name", ["dws_16", "dws_20_pad", "grouped_16", "dense_16"])
def test_unet_train_forward_backward_matches_reference_fixture(name):
    z, meta, cfg, sd_np, x, t = load_unet_case(name)
    sd = {k: v.requires_grad_(True) for k, v in unet_ref.to_torch(sd_np).items()}
    torch.manual_seed(int(z["train_seed"]))
    masks = unet_ref.draw_dropout_masks(cfg, meta["batch"])
    y = unet_ref.forward(sd, torch.from_numpy(x), cfg, masks)
    assert np.abs(y.detach().numpy() - z["prob_train"]).max() < 5e-6
    loss = loss_ref.focal_tversky(y, torch.from_numpy(t))
    assert abs(loss.item() - float(z["loss_train"])) < 1e-5
    loss.backward()
    for k, p in sd.items():
        g = p.grad.numpy()
        gn = float(z[f"gnorm::{k}"])
        assert abs(np.sqrt((g.astype(np.float64) ** 2).sum()) - gn) <= 2e-3 * gn + 1e-8, k
        head = z[f"ghead::{k}"]
        assert np.abs(g.ravel()[:512] - head).max() <= 2e-3 * (np.abs(head).max() + 1e-9) + 1e-9, k


def test_param_counts():
    C = unet_ref.UNetCfg
    assert unet_ref.count_parameters(C()) == 217228
    assert unet_ref.count_parameters(C(use_depthwise_separable=False)) == 391521
    assert unet_ref.count_parameters(C(use_depthwise_separable=False, use_grouped=False)) == 2308737
    assert len(unet_ref.param_shapes(C())) == 93
    macs = unet_ref.forward_flops(C(), 48)
    assert abs(macs["total"] / 1e9 - 0.701) < 0.002
    macs = unet_ref.forward_flops(C(use_depthwise_separable=False, use_grouped=False), 48)
    assert abs(macs["total"] / 1e9 - 6.23) < 0.02


def test_loss_fixture():
    z = np.load(os.path.join(GOLDEN, "loss.npz"))
    rng = np.random.default_rng(5)
    p = rng.random((2, 1, 12, 10, 14), dtype=np.float32)
    t = (rng.random((2, 1, 12, 10, 14)) > 0.9).astype(np.float32)
    for i in range(3):
        a, b, g = z[f"abg{i}"]
        l64, g64 = loss_ref.focal_tversky_closed_form_grad(p, t, a, b, g)
        assert abs(l64 - float(z[f"loss{i}"])) < 1e-6
        assert np.abs(g64 - z[f"grad{i}"]).max() < 1e-5 * np.abs(g64).max()
    for j, tv in enumerate((0.0, 1.0)):
        pp = torch.from_numpy(z[f"edge_p{j}"])
        l = loss_ref.focal_tversky(pp, torch.full_like(pp, tv)).item()
        assert abs(l - float(z[f"edge_loss{j}"])) < 1e-6
    assert abs(loss_ref.dice(torch.from_numpy(p), torch.from_numpy(t)).item() - float(z["dice"])) < 1e-6
    assert abs(loss_ref.combined(torch.from_numpy(p), torch.from_numpy(t)).item() - float(z["combined"])) < 1e-6


def test_gaussian_fixture_and_known_answers():
    z = np.load(os.path.join(GOLDEN, "gaussian.npz"))
    for patch in [(48, 48, 48), (16, 16, 16), (32, 48, 64), (7, 9, 11)]:
        g = stitch_ref.gaussian_importance_map(patch)
        tag = "x".join(map(str, patch))
        assert g.dtype == np.float32
        assert np.array_equal(g[:, patch[1] // 2, patch[2] // 2], z[f"gz_{tag}"])
        assert np.array_equal(g[patch[0] // 2, :, patch[2] // 2], z[f"gy_{tag}"])
        assert np.array_equal(g[patch[0] // 2, patch[1] // 2, :], z[f"gx_{tag}"])
        assert abs(g.astype(np.float64).sum() - float(z[f"gsum_{tag}"])) < 1e-6
        assert g.min() == z[f"gmin_{tag}"]
        assert tuple(np.unravel_index(g.argmax(), g.shape)) == tuple(z[f"gargmax_{tag}"])
    # SURVEY.md section 8(c) known answers for 48^3
    g1 = stitch_ref.gaussian_1d(48)
    for idx, val in {0: .011109, 1: .016038, 23: .992218, 24: 1.0, 25: .992218, 47: .016038}.items():
        assert abs(g1[idx] - val) < 1e-6
    g = stitch_ref.gaussian_importance_map((48, 48, 48))
    assert abs(g.min() - 1.3709591e-6) < 1e-12


def test_window_grid_fixture():
    with open(os.path.join(GOLDEN, "window_grid.json")) as f:
        grids = json.load(f)
    for key, exp in grids.items():
        shape, patch, ov = key.split("|")
        got = stitch_ref.window_grid(eval(shape), eval(patch), float(ov))
        assert [list(p) for p in got] == exp, key
    g = stitch_ref.window_grid((128, 128, 320), (48, 48, 48), 0.5)
    assert list(g[0]) == [0, 24, 48, 72, 80] and len(g[2]) == 13
    assert len(g[0]) * len(g[1]) * len(g[2]) == 325


def test_sliding_window_fixture():
    z = np.load(os.path.join(GOLDEN, "sliding_window.npz"))
    cfg = unet_ref.UNetCfg(dropout_p=0.0)
    sd = unet_ref.to_torch(synth.synth_state_dict(unet_ref.param_shapes(cfg), 3))

    def predict(chunk):
        with torch.no_grad():
            return unet_ref.forward(sd, torch.from_numpy(chunk), cfg).numpy()

    for tag in "abc":
        c = z[f"cfg_{tag}"]
        shape, patch, ov, gauss = tuple(int(v) for v in c[:3]), tuple(int(v) for v in c[3:6]), float(c[6]), bool(c[7])
        vol = synth.synth_volume(shape, seed=9, n_blobs=2)
        got = stitch_ref.sliding_window(vol, predict, patch, ov, gauss, batch=4)
        assert got.shape == shape and got.dtype == np.float32
        assert np.abs(got - z[f"prob_{tag}"]).max() < 1e-5
    with pytest.raises(ValueError):
        stitch_ref.sliding_window(np.zeros((4, 4), np.float32), predict)


def test_ccl_oracle_matches_scipy():
    from scipy import ndimage
    rng = np.random.default_rng(0)
    for shape, dens in [((9, 10, 11), 0.5), ((17, 5, 23), 0.35), ((1, 1, 7), 0.6), ((4, 4, 4), 1.0),
                        ((6, 6, 6), 0.0), ((20, 21, 22), 0.45)]:
        m = (rng.random(shape) < dens).astype(np.int32)
        lab, n = bbox_ref.label6(m)
        lab_s, n_s = ndimage.label(m)
        assert n == n_s and np.array_equal(lab, lab_s)


def test_bbox_fixture():
    with open(os.path.join(GOLDEN, "bbox.json")) as f:
        fx = json.load(f)
    prob = np.zeros((20, 24, 28), dtype=np.float32)
    prob[2:5, 3:6, 4:7] = 0.9
    prob[10:12, 10:12, 10:12] = 0.31
    prob[15, 15, 15] = 0.99
    prob[18:20, 20:24, 25:28] = 0.3
    prob[6, 6, 6] = 0.5
    prob[7, 7, 7] = 0.5
    assert bbox_ref.extract_bboxes(prob, 0.3, 0.5, (4.0, 4.0, 4.0), 3) == fx["kat"]
    for tag, c in fx.items():
        if tag == "kat":
            continue
        prob = synth.synth_prob_map(tuple(c["shape"]), c["seed"])
        got = bbox_ref.extract_bboxes(prob, c["threshold"], c["min_volume_cc"], tuple(c["spacing"]), c["expansion"])
        assert got == c["bboxes"], tag
        mv = int(np.ceil(c["min_volume_cc"] / (np.prod(c["spacing"]) / 1000.0)))
        lab, n = bbox_ref.connected_components((prob >= c["threshold"]).astype(np.int32), mv)
        assert n == c["n"] and int(lab.astype(np.int64).sum()) == c["label_sum"]
        w = np.arange(lab.size, dtype=np.int64) % 1009
        assert int((lab.astype(np.int64).ravel() * w).sum()) == c["label_wsum"]


def test_metrics_oracle_against_reference_fixture():
    """oracle/metrics_ref.py (lesion matching, Dice, the validation threshold sweep) against the numbers the reference's
    own metrics.py / Trainer._is_better_metric produced (tests/golden/metrics.json)."""
    import json
    from oracle import metrics_ref
    with open(os.path.join(GOLDEN, "metrics.json")) as f:
        fx = json.load(f)
    thresholds = [0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8]
    for tag, rec in fx.items():
        pairs = [metrics_ref.synth_case(tuple(shape), seed) for shape, seed in rec["cases"]]
        if tag == "empty":
            pairs = [(np.zeros_like(p), np.zeros_like(l)) for p, l in pairs] + [(pairs[0][0], np.zeros_like(pairs[0][1]))]
        preds, labels = [p for p, _ in pairs], [l for _, l in pairs]
        spacings = [tuple(s) for s in rec["spacings"]]
        for t in thresholds:
            assert metrics_ref.calculate_metrics(preds, labels, t, spacings) == rec["per_threshold"][str(t)], (tag, t)
        for tie in (0.0, 0.05):
            best = metrics_ref.select_threshold(preds, labels, spacings, thresholds, tie)
            assert {k: best[k] for k in ("best_threshold", "best_recall", "best_dsc_macro")} == rec[f"best_tie{tie}"], (tag, tie)
