#!/usr/bin/env python
"""Per-layer timing of the small-volume depthwise-separable layers (12^3 / 6^3 levels of a 48^3 window): slab kernel
(csrc/l3d_fwd_slab.cu) against the 4x8x8-tile stencil kernel (L3D_NO_SLAB=1 must be set at process start to compare).
    python tools/time_small.py [N]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "light-3d-unet-front_b200"))
from light_unet import _native as nv

N = int(sys.argv[1]) if len(sys.argv) > 1 else 325
DEV = torch.device("cuda:0")
LAYERS = [("down2.c1", 32, 64, True, 12, False), ("down2.c2", 64, 64, False, 12, True), ("down3.c1", 64, 128, True, 6, False),
          ("down3.c2", 128, 128, False, 6, True), ("bott.c1", 128, 128, False, 6, False), ("bott.c2", 128, 128, False, 6, True),
          ("up1.c1", 128, 64, True, 12, False), ("up1.c2", 64, 64, False, 12, True)]
tot = 0.0
for name, Cin, Cout, sc, S, normed in LAYERS:
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
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        go()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 10 * 1e3
    tot += us
    gb = 2 * N * vox * (Cin + Cout * (2 if sc else 1)) / 1e9
    print(f"{name:9s} {Cin:3d}->{Cout:3d}{'+sc' if sc else '   '} @{S}^3  {us:8.1f} us  {gb / us * 1e6:7.0f} GB/s  [{nv.lib().l3d_last_kernel().decode()}]")
print(f"total {tot:.1f} us")
