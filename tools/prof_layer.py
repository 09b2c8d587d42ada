#!/usr/bin/env python
"""One depthwise-separable conv layer launch for ncu captures: python tools/prof_layer.py Cin Cout sc size B [normed]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "light-3d-unet-front_b200"))
from light_unet import _native as nv
Cin, Cout, sc, S, B = (int(v) for v in sys.argv[1:6])
normed = len(sys.argv) > 6 and sys.argv[6] == "1"
DEV = torch.device("cuda:0")
torch.manual_seed(0)
x = torch.randn(B, S, S, S, Cin, device=DEV).to(torch.float16)
vox = S ** 3
xf = x.float()
stats = torch.stack([xf.sum(dim=(1, 2, 3)), (xf * xf).sum(dim=(1, 2, 3))]).double().contiguous()
gamma, beta = torch.ones(Cin, device=DEV), torch.zeros(Cin, device=DEV)
xn = nv.norm(stats, gamma, beta, None, 1e-5, 0.01, vox) if normed else nv.norm()
dw = torch.randn(Cin, 27, device=DEV) / 5
pw = torch.randn(Cout, Cin, device=DEV) / Cin ** 0.5
scw = torch.randn(Cout, Cin, device=DEV) / Cin ** 0.5 if sc else None
t = torch.empty(B, S, S, S, Cout, dtype=torch.float16, device=DEV)
r = torch.empty_like(t) if sc else None
ts = torch.zeros(2 * B * Cout, dtype=torch.float64, device=DEV)
rs = torch.zeros_like(ts)
st = nv.stream_ptr(DEV)
for _ in range(3):
    nv.call("l3d_dwpw_fwd", nv.act(x), xn, B, S, S, S, nv.ptr(dw), nv.ptr(pw), nv.ptr(scw), nv.act(t), nv.ptr(ts),
            nv.act(r), nv.ptr(rs) if sc else None, nv.act(None), st)
torch.cuda.synchronize()
print("done")
