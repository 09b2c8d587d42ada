#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the REFERENCE itself.

Run in the build container only (needs /root/reference, which is absent on the
GPU box):   python -B tests/golden/make_golden.py

For every case it (1) runs the unmodified reference code, (2) runs the oracle
restatement on the same inputs and asserts they agree (this is what pins the
oracle), and (3) stores the reference's outputs.  Inputs and weights are not
stored: they are regenerated from numpy PCG64 seeds by oracle/synth.py.
nibabel is not installed; NIfTI I/O is off the hot path, so an empty stub
module is registered before importing the reference.
"""
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.dont_write_bytecode = True
sys.modules.setdefault("nibabel", types.ModuleType("nibabel"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

from light_unet.models.unet3d import Lightweight3DUNet as RefUNet          # noqa: E402
from light_unet.models.losses import FocalTverskyLoss as RefFTL            # noqa: E402
from light_unet.models.losses import CombinedLoss as RefCombined, DiceLoss as RefDice  # noqa: E402
from light_unet.models.metrics import get_connected_components as ref_cc  # noqa: E402
from light_unet.utils import sliding_window_inference_3d as ref_sw        # noqa: E402
from light_unet.utils import _get_gaussian_importance_map as ref_gauss    # noqa: E402
from light_unet.core.inferencer import Inferencer as RefInferencer        # noqa: E402

assert "/root/reference" in sys.modules["light_unet"].__file__, "must import the reference package"

from oracle import unet_ref, loss_ref, stitch_ref, bbox_ref, synth, metrics_ref, augment_ref   # noqa: E402
from light_unet.models import metrics as ref_metrics                      # noqa: E402

torch.set_num_threads(8)


def ref_model(cfg: unet_ref.UNetCfg, seed: int):
    m = RefUNet(in_channels=cfg.in_channels, out_channels=cfg.out_channels, start_channels=16,
                encoder_channels=list(cfg.encoder_channels),
                use_depthwise_separable=cfg.use_depthwise_separable, use_grouped=cfg.use_grouped,
                groups=cfg.groups, dropout_p=cfg.dropout_p)
    shapes = unet_ref.param_shapes(cfg)
    ref_sd = m.state_dict()
    assert list(ref_sd.keys()) == list(shapes.keys()), "state_dict key order differs from oracle.param_shapes"
    for k, v in ref_sd.items():
        assert tuple(v.shape) == tuple(shapes[k]), (k, v.shape, shapes[k])
    sd_np = synth.synth_state_dict(shapes, seed)
    m.load_state_dict(unet_ref.to_torch(sd_np), strict=True)
    return m, sd_np


def grads_summary(named_grads):
    out = {}
    for k, g in named_grads.items():
        g = g.detach().numpy().astype(np.float32)
        out[f"gnorm::{k}"] = np.float64(np.sqrt((g.astype(np.float64) ** 2).sum()))
        out[f"ghead::{k}"] = g.ravel()[:512].copy()
    return out


def unet_case(name, cfg, size, batch, wseed, xseed, train_seed=None, with_grad=True, subsample=None):
    m, sd_np = ref_model(cfg, wseed)
    x_np, t_np = synth.synth_patches(batch, size, xseed)
    x, t = torch.from_numpy(x_np), torch.from_numpy(t_np)
    sd = unet_ref.to_torch(sd_np)
    rec = {}
    # --- eval forward
    m.eval()
    with torch.no_grad():
        y_ref = m(x)
        y_or = unet_ref.forward(sd, x, cfg)
        lg_or = unet_ref.forward(sd, x, cfg, return_logits=True)
    err = (y_ref - y_or).abs().max().item()
    assert err < 2e-6, (name, "oracle != reference (eval)", err)
    if subsample:
        rec["prob_eval_sub"] = y_ref.numpy()[:, :, ::subsample, ::subsample, ::subsample].copy()
        rec["logit_eval_sub"] = lg_or.numpy()[:, :, ::subsample, ::subsample, ::subsample].copy()
        rec["prob_eval_stats"] = np.array([y_ref.mean().item(), y_ref.min().item(), y_ref.max().item(),
                                           y_ref.double().pow(2).sum().sqrt().item()])
    else:
        rec["prob_eval"] = y_ref.numpy().copy()
        rec["logit_eval"] = lg_or.numpy().copy()
    loss_ref_eval = RefFTL()(y_ref, t).item()
    rec["loss_eval"] = np.float64(loss_ref_eval)
    assert abs(loss_ref.focal_tversky(y_or, t).item() - loss_ref_eval) < 1e-6
    # --- train forward + backward (dropout active)
    if train_seed is not None:
        m.train()
        torch.manual_seed(train_seed)
        y_tr = m(x)
        loss = RefFTL()(y_tr, t)
        m.zero_grad()
        loss.backward()
        ref_grads = {k: p.grad.clone() for k, p in m.named_parameters()}
        torch.manual_seed(train_seed)
        masks = unet_ref.draw_dropout_masks(cfg, batch)
        sd_g = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        y_or_tr = unet_ref.forward(sd_g, x, cfg, masks)
        err = (y_tr - y_or_tr).abs().max().item()
        assert err < 2e-6, (name, "oracle != reference (train; dropout-mask recipe)", err)
        l_or = loss_ref.focal_tversky(y_or_tr, t)
        l_or.backward()
        for k in ref_grads:
            ge = (ref_grads[k] - sd_g[k].grad).abs().max().item()
            gs = ref_grads[k].abs().max().item() + 1e-12
            assert ge <= 2e-4 * gs + 1e-9, (name, k, ge, gs)
        if cfg.dropout_p > 0:
            assert any((mk == 0).any().item() for mk in masks), "fixture should actually drop a channel"
        if subsample:
            rec["prob_train_sub"] = y_tr.detach().numpy()[:, :, ::subsample, ::subsample, ::subsample].copy()
        else:
            rec["prob_train"] = y_tr.detach().numpy().copy()
        rec["loss_train"] = np.float64(loss.item())
        rec["train_seed"] = np.int64(train_seed)
        if with_grad:
            rec.update(grads_summary(ref_grads))
    meta = dict(in_channels=cfg.in_channels, out_channels=cfg.out_channels,
                encoder_channels=list(cfg.encoder_channels), dws=cfg.use_depthwise_separable,
                grouped=cfg.use_grouped, groups=cfg.groups, dropout_p=cfg.dropout_p,
                size=list(size) if not isinstance(size, int) else [size] * 3, batch=batch,
                wseed=wseed, xseed=xseed, subsample=subsample or 0,
                n_params=int(sum(p.numel() for p in m.parameters())))
    rec["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, f"unet_{name}.npz"), **rec)
    print(f"unet_{name}: ok  params={meta['n_params']}  loss_eval={loss_ref_eval:.6f}")


def loss_cases():
    rec = {}
    rng = np.random.default_rng(5)
    cases = [(0.7, 0.3, 0.75), (0.5, 0.5, 1.0), (0.3, 0.7, 2.0)]
    p_np = rng.random((2, 1, 12, 10, 14), dtype=np.float32)
    t_np = (rng.random((2, 1, 12, 10, 14)) > 0.9).astype(np.float32)
    for i, (a, b, g) in enumerate(cases):
        p = torch.from_numpy(p_np.copy()).requires_grad_(True)
        loss = RefFTL(alpha=a, beta=b, gamma=g)(p, torch.from_numpy(t_np))
        loss.backward()
        lo = loss_ref.focal_tversky(torch.from_numpy(p_np), torch.from_numpy(t_np), a, b, g).item()
        assert abs(lo - loss.item()) < 1e-7
        l64, g64 = loss_ref.focal_tversky_closed_form_grad(p_np, t_np, a, b, g)
        assert abs(l64 - loss.item()) < 1e-6
        assert np.abs(g64 - p.grad.numpy()).max() < 1e-6 * np.abs(g64).max() + 1e-10
        rec[f"loss{i}"] = np.float64(loss.item())
        rec[f"grad{i}"] = p.grad.numpy().copy()
        rec[f"abg{i}"] = np.array([a, b, g])
    # edge: all-zero target, all-one target
    for j, tv in enumerate((0.0, 1.0)):
        tt = torch.full((1, 1, 4, 4, 4), tv)
        pp = torch.from_numpy(rng.random((1, 1, 4, 4, 4), dtype=np.float32))
        rec[f"edge_p{j}"] = pp.numpy().copy()
        rec[f"edge_loss{j}"] = np.float64(RefFTL()(pp, tt).item())
    pp = torch.from_numpy(p_np)
    tt = torch.from_numpy(t_np)
    rec["dice"] = np.float64(RefDice()(pp, tt).item())
    rec["combined"] = np.float64(RefCombined()(pp, tt).item())
    assert abs(loss_ref.dice(pp, tt).item() - rec["dice"]) < 1e-7
    assert abs(loss_ref.combined(pp, tt).item() - rec["combined"]) < 1e-6
    np.savez_compressed(os.path.join(HERE, "loss.npz"), **rec)
    print("loss: ok")


def gaussian_and_grid_cases():
    rec = {}
    for patch in [(48, 48, 48), (16, 16, 16), (32, 48, 64), (7, 9, 11)]:
        g = ref_gauss(patch)
        go = stitch_ref.gaussian_importance_map(patch)
        assert g.dtype == np.float32 and np.array_equal(g, go), patch
        tag = "x".join(map(str, patch))
        rec[f"gz_{tag}"] = g[:, patch[1] // 2, patch[2] // 2].copy()
        rec[f"gy_{tag}"] = g[patch[0] // 2, :, patch[2] // 2].copy()
        rec[f"gx_{tag}"] = g[patch[0] // 2, patch[1] // 2, :].copy()
        rec[f"gsum_{tag}"] = np.float64(g.astype(np.float64).sum())
        rec[f"gmin_{tag}"] = np.float32(g.min())
        rec[f"gargmax_{tag}"] = np.array(np.unravel_index(g.argmax(), g.shape))
    np.savez_compressed(os.path.join(HERE, "gaussian.npz"), **rec)

    # window grid: record the positions the reference visits, by spying on a fake model
    class Spy(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.zeros(1))
            self.seen = []

        def forward(self, x):
            self.seen.append(float(x[0, 0, 0, 0, 0]))
            return x

    grids = {}
    for shape, patch, ov in [((128, 128, 320), (48, 48, 48), 0.5), ((144, 144, 320), (48, 48, 48), 0.5),
                             ((128, 128, 320), (64, 64, 64), 0.5), ((128, 128, 320), (96, 96, 96), 0.5),
                             ((40, 40, 40), (48, 48, 48), 0.5), ((50, 48, 100), (48, 48, 48), 0.5),
                             ((20, 28, 36), (16, 16, 16), 0.5), ((30, 30, 30), (16, 16, 16), 0.25),
                             ((33, 17, 16), (16, 16, 16), 0.75), ((10, 40, 23), (16, 16, 16), 0.0)]:
        # encode the linear voxel index in the image so the spy can recover each window origin
        n = int(np.prod(shape))
        img = (np.arange(n, dtype=np.float64)).reshape(shape)
        # float32 cannot hold indices > 2^24 exactly: use per-axis passes instead
        pos = []
        for ax in range(3):
            ramp_shape = [1, 1, 1]
            ramp_shape[ax] = shape[ax]
            img = np.broadcast_to(np.arange(shape[ax], dtype=np.float32).reshape(ramp_shape), shape).copy()
            spy = Spy()
            ref_sw(img, spy, patch_size=patch, overlap=ov, device=torch.device("cpu"), use_gaussian=True)
            pos.append(spy.seen)
        zyx = list(zip(*[[int(v) for v in p] for p in pos]))
        exp = stitch_ref.window_grid(shape, patch, ov)
        want = [(z, y, x) for z in exp[0] for y in exp[1] for x in exp[2]]
        assert zyx == want, (shape, patch, ov)
        grids[f"{shape}|{patch}|{ov}"] = [list(map(int, p)) for p in exp]
    with open(os.path.join(HERE, "window_grid.json"), "w") as f:
        json.dump(grids, f, indent=1)
    print("gaussian + grid: ok")


def sliding_window_cases():
    cfg = unet_ref.UNetCfg(dropout_p=0.0)
    m, sd_np = ref_model(cfg, 3)
    m.eval()
    sd = unet_ref.to_torch(sd_np)

    def predict(chunk):
        with torch.no_grad():
            return unet_ref.forward(sd, torch.from_numpy(chunk), cfg).numpy()

    rec = {}
    for tag, shape, patch, ov, gauss in [("a", (20, 28, 36), (16, 16, 16), 0.5, True),
                                         ("b", (12, 28, 20), (16, 16, 16), 0.5, True),
                                         ("c", (24, 24, 24), (16, 16, 16), 0.25, False)]:
        vol = synth.synth_volume(shape, seed=9, n_blobs=2)
        ref = ref_sw(vol, m, patch_size=patch, overlap=ov, device=torch.device("cpu"), use_gaussian=gauss)
        got = stitch_ref.sliding_window(vol, predict, patch, ov, gauss, batch=1)
        assert ref.dtype == np.float32 and ref.shape == shape
        err = np.abs(ref - got).max()
        assert err < 2e-6, (tag, err)
        got_b = stitch_ref.sliding_window(vol, predict, patch, ov, gauss, batch=5)
        assert np.abs(ref - got_b).max() < 5e-6
        rec[f"prob_{tag}"] = ref
        rec[f"cfg_{tag}"] = np.array(list(shape) + list(patch) + [ov, float(gauss)])
    # 4-D input accepted, 2-D rejected (utils.py:37-41)
    vol = synth.synth_volume((20, 28, 36), seed=9, n_blobs=2)
    r4 = ref_sw(vol[None], m, patch_size=(16, 16, 16), overlap=0.5, device=torch.device("cpu"))
    assert np.array_equal(r4, rec["prob_a"])
    try:
        ref_sw(vol[0], m, patch_size=(16, 16, 16))
        raise AssertionError("reference should reject 2-D input")
    except ValueError:
        pass
    np.savez_compressed(os.path.join(HERE, "sliding_window.npz"), **rec)
    print("sliding window: ok")


def bbox_cases():
    inf = RefInferencer.__new__(RefInferencer)
    inf.config = {"data": {"bbox_expansion_voxels": 3}}
    out = {}
    # hand-checked known-answer case of SURVEY.md section 8(c)
    prob = np.zeros((20, 24, 28), dtype=np.float32)
    prob[2:5, 3:6, 4:7] = 0.9
    prob[10:12, 10:12, 10:12] = 0.31
    prob[15, 15, 15] = 0.99
    prob[18:20, 20:24, 25:28] = 0.3
    prob[6, 6, 6] = 0.5
    prob[7, 7, 7] = 0.5
    kat = inf.extract_bboxes(prob, threshold=0.3, min_volume_cc=0.5, spacing=(4.0, 4.0, 4.0))
    assert [b["bbox_voxel"] for b in kat] == [[0, 7, 0, 8, 1, 9], [7, 14, 7, 14, 7, 14], [15, 19, 17, 23, 22, 27]]
    assert kat == bbox_ref.extract_bboxes(prob, 0.3, 0.5, (4.0, 4.0, 4.0), 3)
    out["kat"] = kat
    cases = [("blobs7", (40, 48, 56), 7, 0.3, 0.5, (4.0, 4.0, 4.0), 3),
             ("blobs8", (33, 21, 47), 8, 0.5, 0.25, (4.0, 4.0, 4.0), 3),
             ("blobs9", (24, 64, 30), 9, 0.3, 0.5, (2.0, 3.0, 5.0), 1),
             ("blobs10", (16, 16, 16), 10, 0.3, 0.0, (4.0, 4.0, 4.0), 0)]
    for tag, shape, seed, thr, mincc, spacing, expand in cases:
        inf.config = {"data": {"bbox_expansion_voxels": expand}}
        prob = synth.synth_prob_map(shape, seed)
        ref = inf.extract_bboxes(prob, threshold=thr, min_volume_cc=mincc, spacing=spacing)
        got = bbox_ref.extract_bboxes(prob, thr, mincc, spacing, expand)
        assert ref == got, tag
        # labelling parity with the reference's get_connected_components (scipy)
        binary = (prob >= thr).astype(np.int32)
        mv = int(np.ceil(mincc / (spacing[0] * spacing[1] * spacing[2] / 1000.0)))
        lab_ref, n_ref = ref_cc(binary.copy(), min_size=mv)
        lab_or, n_or = bbox_ref.connected_components(binary.copy(), mv)
        assert n_ref == n_or and np.array_equal(lab_ref, lab_or), tag
        out[tag] = {"shape": list(shape), "seed": seed, "threshold": thr, "min_volume_cc": mincc,
                    "spacing": list(spacing), "expansion": expand, "n": int(n_ref),
                    "label_sum": int(lab_ref.astype(np.int64).sum()),
                    "label_wsum": int((lab_ref.astype(np.int64).ravel() *
                                       (np.arange(lab_ref.size, dtype=np.int64) % 1009)).sum()),
                    "bboxes": ref}
        print(f"bbox {tag}: {n_ref} components")
    with open(os.path.join(HERE, "bbox.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("bbox: ok")


def metrics_cases():
    """Lesion-wise metrics and the validation threshold sweep: the reference's own calculate_lesion_metrics /
    calculate_metrics / match_components (metrics.py:127-404) and Trainer._is_better_metric (trainer.py:183-189) on seeded
    synthetic cases; the oracle restatement must reproduce every number exactly."""
    from light_unet.core.trainer import Trainer as RefTrainer
    out = {}
    thresholds = [0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8]
    sets = {"small": [((24, 28, 32), 3), ((20, 20, 40), 4), ((32, 24, 24), 5)],
            "spacing": [((24, 28, 32), 13), ((18, 30, 26), 14)],
            "empty": [((12, 12, 12), 21)]}
    for tag, cases in sets.items():
        pairs = [metrics_ref.synth_case(shape, seed) for shape, seed in cases]
        if tag == "empty":
            pairs = [(np.zeros_like(p), np.zeros_like(l)) for p, l in pairs] + [(pairs[0][0], np.zeros_like(pairs[0][1]))]
        preds, labels = [p for p, _ in pairs], [l for _, l in pairs]
        spacings = [(4.0, 4.0, 4.0)] * len(preds) if tag != "spacing" else [(2.0, 3.0, 5.0), (4.0, 4.0, 2.5)]
        rec = {"cases": cases, "spacings": spacings, "per_threshold": {}, "lesion": []}
        for t in thresholds:
            ref = ref_metrics.calculate_metrics(preds, labels, threshold=t, spacing=spacings)
            ora = metrics_ref.calculate_metrics(preds, labels, t, spacings)
            assert ref == ora, (tag, t, ref, ora)
            rec["per_threshold"][str(t)] = {k: (int(v) if isinstance(v, (int, np.integer)) else float(v)) for k, v in ref.items()}
        for p, l, sp in zip(preds, labels, spacings):
            for ms in (0, 8):
                ref = ref_metrics.calculate_lesion_metrics(p, l, threshold=0.4, min_size_voxels=ms, spacing=sp)
                ora = metrics_ref.calculate_lesion_metrics(p, l, 0.4, ms, spacing=sp)
                assert ref == ora, (tag, ms, ref, ora)
                rec["lesion"].append({k: (int(v) if isinstance(v, (int, np.integer)) else float(v)) for k, v in ref.items()})
        # the sweep of Trainer.validate (trainer.py:423-445) with the reference's own comparison
        tr = RefTrainer.__new__(RefTrainer)
        for tie in (0.0, 0.05):
            best_t = thresholds[0]
            best = ref_metrics.calculate_metrics(preds, labels, threshold=best_t, spacing=spacings)
            br, bd = best["lesion_wise_recall"], best["voxel_wise_dsc_macro"]
            for t in thresholds[1:]:
                m = ref_metrics.calculate_metrics(preds, labels, threshold=t, spacing=spacings)
                better, _ = tr._is_better_metric(m["lesion_wise_recall"], m["voxel_wise_dsc_macro"], br, bd, tie)
                if better:
                    br, bd, best_t, best = m["lesion_wise_recall"], m["voxel_wise_dsc_macro"], t, m
            ora = metrics_ref.select_threshold(preds, labels, spacings, thresholds, tie)
            assert ora["best_threshold"] == best_t and ora["best_recall"] == br and ora["best_dsc_macro"] == bd, (tag, tie)
            rec[f"best_tie{tie}"] = {"best_threshold": best_t, "best_recall": float(br), "best_dsc_macro": float(bd)}
        out[tag] = rec
    with open(os.path.join(HERE, "metrics.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("metrics: ok")


def patch_cases():
    """The reference's own PatchDataset (patch_dataset.py) on in-memory volumes: nibabel.load is replaced by a table lookup
    and the case files exist as empty files so that find_case_files finds them.  Stores, per configuration, the SHA-256 of
    the first items' image / label patches (as float32, which is what trainer.py:225 makes of them)."""
    import hashlib
    import tempfile
    import nibabel
    from light_unet.datasets.patch_dataset import PatchDataset as RefPatchDataset
    vols = augment_ref.synth_cases()
    PATCH_AUG = augment_ref.PATCH_AUG
    store = {}

    class FakeImg:
        def __init__(self, a):
            self.a = a

        def get_fdata(self):
            return self.a.astype(np.float64)
    nibabel.load = lambda path: FakeImg(store[str(path)])
    out = {}
    with tempfile.TemporaryDirectory() as d:
        for sub in ("images", "labels", "body_masks"):
            os.makedirs(os.path.join(d, sub))
        ids = []
        for i, (image, label, body) in enumerate(vols):
            cid = f"{i + 1:04d}"
            ids.append(cid)
            for path, arr in ((os.path.join(d, "images", f"{cid}_0000.nii.gz"), image), (os.path.join(d, "labels", f"{cid}.nii.gz"), label)):
                open(path, "wb").close()
                store[path] = arr
            if body is not None:
                path = os.path.join(d, "body_masks", f"{cid}.nii.gz")
                open(path, "wb").close()
                store[path] = body
        split = os.path.join(d, "split.txt")
        with open(split, "w") as f:
            f.write("\n".join(ids) + "\n")
        for tag, patch, aug, seed, n in [("plain16", (16, 16, 16), None, 42, 12), ("aug16", (16, 16, 16), PATCH_AUG, 7, 40),
                                         ("aug24", (24, 20, 28), PATCH_AUG, 11, 24), ("edge48", (48, 48, 48), PATCH_AUG, 3, 10)]:
            import contextlib
            import io
            with contextlib.redirect_stdout(io.StringIO()):
                ds = RefPatchDataset(d, split, patch_size=patch, lesion_patch_ratio=0.5, augmentation=aug, seed=seed,
                                     body_mask_config={"enabled": True, "apply_to_training_sampling": False})
            items = []
            for k in range(n):
                img, lab = ds[k]
                assert tuple(img.shape) == (1,) + patch and tuple(lab.shape) == (1,) + patch
                a, b = img.numpy().astype(np.float32), lab.numpy().astype(np.float32)
                items.append({"img_sha": hashlib.sha256(a.tobytes()).hexdigest(), "lab_sha": hashlib.sha256(b.tobytes()).hexdigest(),
                              "img_sum": float(a.astype(np.float64).sum()), "lab_sum": float(b.sum())})
            out[tag] = {"patch": list(patch), "seed": seed, "aug": aug is not None, "n_lesion": len(ds.lesion_locations),
                        "n_background": len(ds.background_locations), "items": items}
    with open(os.path.join(HERE, "patches.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("patches: ok", {k: (v["n_lesion"], v["n_background"]) for k, v in out.items()})


def default_init_case():
    """Seeded default initialisation of the reference module tree: the drop-in creates the same torch.nn layers
    in the same order, so its parameters must be identical under the same seed."""
    rec = {}
    for tag, kw in [("dws", {}), ("grouped", dict(use_depthwise_separable=False)),
                    ("dense", dict(use_depthwise_separable=False, use_grouped=False))]:
        torch.manual_seed(1234)
        m = RefUNet(**kw)
        rec[tag] = {k: [float(v.double().sum()), float(v.double().abs().sum())] for k, v in m.state_dict().items()}
    with open(os.path.join(HERE, "default_init.json"), "w") as f:
        json.dump(rec, f)
    print("default init: ok")


def main():
    default_init_case()
    C = unet_ref.UNetCfg
    unet_case("dws_16", C(dropout_p=0.3), 16, 2, wseed=1, xseed=11, train_seed=123)
    unet_case("dws_24", C(dropout_p=0.0), 24, 1, wseed=2, xseed=12, train_seed=5)
    unet_case("dws_20_pad", C(dropout_p=0.3), (20, 18, 22), 2, wseed=3, xseed=13, train_seed=77)
    unet_case("dws_small_enc", C(encoder_channels=(8, 16, 32, 64), dropout_p=0.3), 16, 1, wseed=4, xseed=14,
              train_seed=9)
    unet_case("grouped_16", C(use_depthwise_separable=False, use_grouped=True, dropout_p=0.3), 16, 2, wseed=5,
              xseed=15, train_seed=31)
    unet_case("dense_16", C(use_depthwise_separable=False, use_grouped=False, dropout_p=0.3), 16, 1, wseed=6,
              xseed=16, train_seed=32)
    unet_case("dws_48_c1", C(dropout_p=0.1), 48, 2, wseed=7, xseed=42, train_seed=42, with_grad=True, subsample=5)
    loss_cases()
    gaussian_and_grid_cases()
    sliding_window_cases()
    bbox_cases()
    metrics_cases()
    patch_cases()


if __name__ == "__main__":
    main()
