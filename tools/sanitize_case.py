#!/usr/bin/env python
"""Smallest end-to-end case for compute-sanitizer (one tool per gpurun call, see B200_PROFILING.md):

    compute-sanitizer --tool racecheck|initcheck|synccheck|memcheck python tools/sanitize_case.py [train|infer|patches ...]

train:   forward + Focal Tversky + backward + AdamW on 2 x 16^3 patches, fp32 and fp16 storage (every training kernel);
infer:   sliding-window inference of a 48x48x72 volume (48^3 windows) -> stitch -> threshold -> CC -> boxes, and the
         lesion-metric sweep of one case;
patches: one augmented batch from the device patch sampler.
Summaries are kept under profiles/r02_sanitizer_*.txt."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "light-3d-unet-front_b200"))
from light_unet.models import Lightweight3DUNet, FocalTverskyLoss          # noqa: E402
from oracle import augment_ref, metrics_ref, synth, unet_ref             # noqa: E402  (synthetic inputs only)

what = sys.argv[1:] or ["train", "infer", "patches"]
dev = torch.device("cuda:0")
cfg = unet_ref.UNetCfg(dropout_p=0.1)
sd = unet_ref.to_torch(synth.synth_state_dict(unet_ref.param_shapes(cfg), 1))
if "train" in what:
    for dtype in ("f32", "f16"):
        m = Lightweight3DUNet(dropout_p=0.1)
        m.load_state_dict(sd)
        m = m.to(dev).set_compute_dtype(dtype).train()
        opt = torch.optim.AdamW(m.parameters(), lr=1e-4, weight_decay=1e-5)
        x, t = synth.synth_patches(2, 16, 11)
        for _ in range(2):
            opt.zero_grad()
            loss = FocalTverskyLoss()(m(torch.from_numpy(x).to(dev)), torch.from_numpy(t).to(dev))
            loss.backward()
            opt.step()
        torch.cuda.synchronize()
        print(f"train {dtype}: loss {loss.item():.6f}")
if "infer" in what:
    from light_unet.core.inferencer import Inferencer
    from light_unet.core.validation import DeviceValidator
    m = Lightweight3DUNet(dropout_p=0.0)
    m.load_state_dict(sd)
    inf = Inferencer.__new__(Inferencer)
    inf.config = {"data": {"bbox_expansion_voxels": 3, "patch_size": [48, 48, 48], "volume_threshold": {"inference_cc": 0.5}},
                  "validation": {"default_threshold": 0.5, "threshold_sensitivity_range": [0.4, 0.5]}, "metrics": {"model_selection": {}}}
    inf.device, inf.model = dev, m.to(dev).eval()
    vol = synth.synth_volume((48, 48, 72), seed=5, n_blobs=2)
    prob, boxes = inf.infer_volume(vol, threshold=0.5)
    v = DeviceValidator(inf.model, inf.config, dev)
    v.add_probability_map(prob, metrics_ref.synth_case((48, 48, 72), 3)[1])
    torch.cuda.synchronize()
    print(f"infer: {len(boxes)} boxes, sweep {v.result()[1]['best_threshold']}")
if "patches" in what:
    from light_unet.datasets import DevicePatchSampler
    s = DevicePatchSampler(augment_ref.synth_cases(), (16, 16, 16), 0.5, augment_ref.PATCH_AUG, 7, dev)
    for _ in range(3):
        a, b = s.sample_batch(8)
    torch.cuda.synchronize()
    print(f"patches: {tuple(a.shape)} mean {float(a.mean()):.4f}")
