#!/usr/bin/env python
"""Per-tile timeline of CTA 0 of the rank-1-input implicit-GEMM conv (development tool): python tools/timeline_r1.py [N]"""
import os, sys, ctypes
os.environ["L3D_C3_DEBUG_SKIP"] = str(8 | int(os.environ.get("SKIP", "0")))
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "light-3d-unet-front_b200"))
from light_unet import _native as nv
N = int(sys.argv[1]) if len(sys.argv) > 1 else 325
S, C = 48, 16
DEV = torch.device("cuda:0")
torch.manual_seed(0)
u = torch.randn(N, S, S, S, device=DEV)
pw1 = torch.randn(C, device=DEV)
t1 = u[..., None] * pw1
stats = torch.stack([t1.sum(dim=(1, 2, 3)), (t1 * t1).sum(dim=(1, 2, 3))]).double().contiguous()
del t1
gamma, beta = torch.ones(C, device=DEV), torch.zeros(C, device=DEV)
xn = nv.norm(stats, gamma, beta, None, 1e-5, 0.01, S ** 3)
dw = torch.randn(C, 27, device=DEV) / 5
pw = torch.randn(C, C, device=DEV) / 4
t = torch.empty(N, S, S, S, C, dtype=torch.float16, device=DEV)
ts = torch.zeros(2 * N * C, dtype=torch.float64, device=DEV)
st = nv.stream_ptr(DEV)
for _ in range(3):
    nv.call("l3d_dwpw_fwd_rank1", nv.ptr(u), nv.ptr(pw1), C, xn, N, S, S, S, nv.ptr(dw), nv.ptr(pw), nv.act(t), nv.ptr(ts), st)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
nv.call("l3d_dwpw_fwd_rank1", nv.ptr(u), nv.ptr(pw1), C, xn, N, S, S, S, nv.ptr(dw), nv.ptr(pw), nv.act(t), nv.ptr(ts), st)
e1.record()
torch.cuda.synchronize()
print(f"{e0.elapsed_time(e1) * 1e3:.1f} us")
n = 40
buf = (ctypes.c_longlong * (n * 8))()
nv.lib().l3d_conv3_debug_read(buf, n * 8)
t0 = buf[0]
names = ["w:tma", "w:Afree", "w:act", "w:acc", "w:epi", "i:Aok", "i:accfree", "i:mma"]
print("item " + " ".join(f"{x:>9s}" for x in names) + "   (clocks since first box landed)")
for i in range(n):
    row = [buf[i * 8 + k] for k in range(8)]
    print(f"{i:4d} " + " ".join(f"{(v - t0) if v else 0:9d}" for v in row))
