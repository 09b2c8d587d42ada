#!/usr/bin/env python
"""Per-parameter gradient error of the CUDA training path vs the oracle (development tool)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "light-3d-unet-front_b200"), ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from helpers import load_unet_case
from oracle import loss_ref, unet_ref
from test_gpu_parity import rel_l2
from test_gpu_train import run_ours

for name in sys.argv[2:] or ["dws_16"]:
    dtype = sys.argv[1] if len(sys.argv) > 1 else "f16"
    z, meta, cfg, sd_np, x, t = load_unet_case(name)
    torch.manual_seed(int(z["train_seed"]))
    masks = unet_ref.draw_dropout_masks(cfg, meta["batch"])
    prob, loss, grads, _ = run_ours(cfg, sd_np, x, t, masks, dtype)
    sd = {k: v.requires_grad_(True) for k, v in unet_ref.to_torch(sd_np).items()}
    ref = unet_ref.forward(sd, torch.from_numpy(x), cfg, masks, quant=unet_ref.f16_storage if os.environ.get('Q') else None)
    loss_ref.focal_tversky(ref, torch.from_numpy(t)).backward()
    print(f"== {name}/{dtype}: prob rel-L2 {rel_l2(prob, ref.detach().numpy()):.3e} loss {loss:.6f} ref {float(z['loss_train']):.6f}")
    allg, allr = [], []
    for k, g in grads.items():
        rg = sd[k].grad.numpy()
        allg.append(g.ravel()); allr.append(rg.ravel())
        print(f"   {k:45s} rel-L2 {rel_l2(g, rg):.3e}  |ref| {np.sqrt((rg.astype(np.float64)**2).sum()):.3e}")
    print(f"   ALL PARAMS rel-L2 {rel_l2(np.concatenate(allg), np.concatenate(allr)):.3e}")
