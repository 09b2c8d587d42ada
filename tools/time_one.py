#!/usr/bin/env python
"""Time one depthwise-separable layer shape through l3d_dwpw_fwd (development tool).
    python tools/time_one.py N Cin Cout sc(0|1) S normed(0|1)"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "light-3d-unet-front_b200"))
from light_unet import _native as nv
N, Cin, Cout, sc, S, normed = (int(v) for v in sys.argv[1:7])
DEV = torch.device("cuda:0")
torch.manual_seed(0)
x = torch.randn(N, S, S, S, Cin, device=DEV).to(torch.float16)
vox = S ** 3
xf = x.float()
stats = torch.stack([xf.sum(dim=(1, 2, 3)), (xf * xf).sum(dim=(1, 2, 3))]).double().contiguous()
gamma, beta = torch.ones(Cin, device=DEV), torch.zeros(Cin, device=DEV)
xn = nv.norm(stats, gamma, beta, None, 1e-5, 0.01, vox) if normed else nv.norm()
dw = torch.randn(Cin, 27, device=DEV) / 5
pw = torch.randn(Cout, Cin, device=DEV) / Cin ** 0.5
scw = torch.randn(Cout, Cin, device=DEV) / Cin ** 0.5 if sc else None
t = torch.empty(N, S, S, S, Cout, dtype=torch.float16, device=DEV)
r = torch.empty_like(t) if sc else None
ts = torch.zeros(2 * N * Cout, dtype=torch.float64, device=DEV)
rs = torch.zeros_like(ts)
st = nv.stream_ptr(DEV)
def go():
    nv.call("l3d_dwpw_fwd", nv.act(x), xn, N, S, S, S, nv.ptr(dw), nv.ptr(pw), nv.ptr(scw), nv.act(t), nv.ptr(ts),
            nv.act(r), nv.ptr(rs) if sc else None, nv.act(None), st)
for _ in range(3):
    go()
torch.cuda.synchronize()
c0 = nv.launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    go()
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 10 * 1e3
gb = 2 * N * vox * (Cin + Cout * (2 if sc else 1)) / 1e9
env = {k: v for k, v in os.environ.items() if k.startswith("L3D_")}
print(f"{Cin:3d}->{Cout:3d}{'+sc' if sc else '   '} @{S}^3 N={N} {us:8.1f} us  {gb / us * 1e6:7.0f} GB/s  launches/call {(nv.launch_count() - c0) // 10} [{nv.lib().l3d_last_kernel().decode()}] {env}")
