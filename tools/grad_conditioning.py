"""Development aid (CPU only, oracle only): how ill-conditioned the per-tensor parameter gradients of the training step are.

  1. the oracle with its STORED tensors rounded to b significand bits (straight-through gradient) against the plain fp32
     oracle, b = 11 (fp16 storage), 13, 16: probabilities move by 1e-3 / 2.6e-4 / 3e-5, individual gradient tensors by
     10 % / 5 % / 1.6 % (median) -- so no 16-bit storage format can give per-tensor gradient parity;
  2. the plain fp32 oracle against the same algorithm in float64: the reference's own fp32 round-off moves individual
     gradient tensors by up to ~8e-3 (2e-3 median) -- the floor of any per-tensor tolerance.

Output kept in profiles/r02_grad_conditioning.txt.      python tools/grad_conditioning.py [batch] [size]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import loss_ref, synth, unet_ref   # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
S = int(sys.argv[2]) if len(sys.argv) > 2 else 48
cfg = unet_ref.UNetCfg(dropout_p=0.1)
sd_np = synth.synth_state_dict(unet_ref.param_shapes(cfg), 1)
x, t = synth.synth_patches(B, (S, S, S), 42)
torch.manual_seed(1234)
masks = unet_ref.draw_dropout_masks(cfg, B)


def step(quant=None, dtype=torch.float32):
    sd = {k: v.to(dtype).requires_grad_(True) for k, v in unet_ref.to_torch(sd_np).items()}
    m = [None if k is None else k.to(dtype) for k in masks]
    ref = unet_ref.forward(sd, torch.from_numpy(x).to(dtype), cfg, m, quant=quant)
    loss_ref.focal_tversky(ref, torch.from_numpy(t).to(dtype)).backward()
    return ref.detach().numpy(), {k: v.grad.numpy().astype(np.float64) for k, v in sd.items()}


def rounded(bits):
    def f(v):
        d = v.detach()
        m, e = torch.frexp(d)
        return v + (torch.ldexp(torch.round(m * 2 ** bits) / 2 ** bits, e) - d)
    return f


def report(tag, p1, g1, p0, g0):
    gmax = max(np.linalg.norm(v) for v in g0.values())
    errs = {k: np.linalg.norm(g1[k] - g0[k]) / (np.linalg.norm(g0[k]) + 1e-3 * gmax) for k in g0}
    w = max(errs, key=errs.get)
    print(f"{tag}: prob rel-L2 {np.linalg.norm(p1 - p0) / np.linalg.norm(p0):.3e}; per-tensor gradient rel-L2 worst {errs[w]:.3e} ({w}), "
          f"median {np.median(list(errs.values())):.3e}; head {errs['out_conv.weight']:.3e}, first conv {errs['init_conv.conv1.depthwise.weight']:.3e}")


print(f"batch {B} x {S}^3, 217K depthwise-separable model, dropout 0.1, Focal Tversky .7/.3/.75")
p0, g0 = step()
for bits in (11, 13, 16):
    report(f"stored tensors rounded to {bits} significand bits vs fp32 oracle", *step(rounded(bits)), p0, g0)
p64, g64 = step(dtype=torch.float64)
report("fp32 oracle vs float64 oracle", p0, g0, p64, g64)
