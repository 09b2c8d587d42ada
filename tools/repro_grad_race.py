#!/usr/bin/env python
"""Development aid: run the same training step (fp32 storage) several times on identical inputs and report how far the
parameter gradients of the repetitions are from the first one (per-tensor rel-L2, floor 1e-3 of the largest norm).
Atomic summation order alone gives ~1e-6; anything near 1e-2 is a race.   python tools/repro_grad_race.py [B] [S] [reps]"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "light-3d-unet-front_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import synth, unet_ref
from light_unet.models import Lightweight3DUNet, FocalTverskyLoss
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
S = int(sys.argv[2]) if len(sys.argv) > 2 else 24
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
dev = torch.device("cuda:0")
cfg = unet_ref.UNetCfg(dropout_p=0.0)
sd_np = synth.synth_state_dict(unet_ref.param_shapes(cfg), 1)
x, t = synth.synth_patches(B, (S, S, S), 42)
xs, ts = torch.from_numpy(x).to(dev), torch.from_numpy(t).to(dev)
loss_fn = FocalTverskyLoss()
first = None
for r in range(reps):
    m = Lightweight3DUNet(dropout_p=0.0)
    m.load_state_dict(unet_ref.to_torch(sd_np))
    m = m.to(dev).set_compute_dtype(os.environ.get("DT", "f32")).train()
    loss = loss_fn(m(xs.float()), ts)
    loss.backward()
    g = {k: p.grad.detach().clone() for k, p in m.named_parameters()}
    torch.cuda.synchronize()
    if first is None:
        first = g
        print(f"rep 0: loss {float(loss):.7f}")
        continue
    gmax = max(float(v.norm()) for v in first.values())
    errs = {k: float((g[k] - first[k]).norm()) / max(float(first[k].norm()), 1e-3 * gmax) for k in g}
    w = max(errs, key=errs.get)
    if os.environ.get("VERBOSE") and errs[w] > 1e-2 and not globals().get("_done"):
        _done = True
        for k in g:
            print(f"      {k:44s} {errs[k]:.3e}  |g| {float(first[k].norm()):.3e}")
    print(f"rep {r}: loss {float(loss):.7f}  worst {errs[w]:.3e} ({w})  median {np.median(list(errs.values())):.3e}  tensors > 1e-4: {sum(e > 1e-4 for e in errs.values())}")
