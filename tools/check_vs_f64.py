#!/usr/bin/env python
"""Development aid: one training step (fp32 storage) against the float64 oracle, per tensor.  python tools/check_vs_f64.py B S [nsamples_from]"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "light-3d-unet-front_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import synth, unet_ref
from helpers import oracle_step, per_tensor_errors
from light_unet.models import Lightweight3DUNet, FocalTverskyLoss
B, S = int(sys.argv[1]), int(sys.argv[2])
off = int(sys.argv[3]) if len(sys.argv) > 3 else 0
dev = torch.device("cuda:0")
cfg = unet_ref.UNetCfg(dropout_p=0.0)
sd_np = synth.synth_state_dict(unet_ref.param_shapes(cfg), 1)
x, t = synth.synth_patches(off + B, (S, S, S), 42)
x, t = x[off:], t[off:]
m = Lightweight3DUNet(dropout_p=0.0)
m.load_state_dict(unet_ref.to_torch(sd_np))
m = m.to(dev).set_compute_dtype("f32").train()
loss = FocalTverskyLoss()(m(torch.from_numpy(x).to(dev)), torch.from_numpy(t).to(dev))
loss.backward()
g = {k: p.grad.detach().cpu().numpy() for k, p in m.named_parameters()}
_, l32, g32 = oracle_step(cfg, sd_np, x, t, None)
_, l64, g64 = oracle_step(cfg, sd_np, x, t, None, dtype=torch.float64)
eo, er = per_tensor_errors(g, g64), per_tensor_errors(g32, g64)
wo = max(eo, key=eo.get)
print(f"B={B} S={S} off={off}: loss {float(loss):.7f} (oracle {l64:.7f}); CUDA vs f64 worst {eo[wo]:.3e} ({wo}) median {np.median(list(eo.values())):.3e}; fp32 oracle worst {max(er.values()):.3e} median {np.median(list(er.values())):.3e}")
