#!/usr/bin/env python
"""Per-layer timing of the depthwise-separable conv kernels (development tool): the composed implicit-GEMM kernel
(every tile height) against the stencil + pointwise-GEMM kernel, on the layer shapes of the 217K model.
    python tools/time_layers.py [B]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "light-3d-unet-front_b200"))
from light_unet import _native as nv

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
DEV = torch.device("cuda:0")
LAYERS = [  # name, Cin, Cout, shortcut, size, normed input
    ("init.c2", 16, 16, False, 48, True), ("up3.c1", 32, 16, True, 48, False), ("up3.c2", 16, 16, False, 48, True),
    ("down1.c1", 16, 32, True, 24, False), ("down1.c2", 32, 32, False, 24, True), ("up2.c1", 64, 32, True, 24, False),
    ("up2.c2", 32, 32, False, 24, True), ("down2.c1", 32, 64, True, 12, False), ("down2.c2", 64, 64, False, 12, True),
    ("up1.c2", 64, 64, False, 12, True),
]


def run(layer, env, iters=10):
    name, Cin, Cout, sc, S, normed = layer
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
    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        def go():
            nv.call("l3d_dwpw_fwd", nv.act(x), xn, B, S, S, S, nv.ptr(dw), nv.ptr(pw), nv.ptr(scw), nv.act(t), nv.ptr(ts),
                    nv.act(r), nv.ptr(rs) if sc else None, nv.act(None), st)
        for _ in range(3):
            go()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            go()
        e1.record()
        torch.cuda.synchronize()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    us = e0.elapsed_time(e1) / iters * 1e3
    gb = 2 * B * vox * (Cin + Cout * (2 if sc else 1)) / 1e9
    return us, gb / (us * 1e-6)


print(f"B={B}")
for layer in LAYERS:
    cols = []
    us, gbs = run(layer, {"L3D_DWS_IGEMM_MAX": 0})
    cols.append(f"stencil {us:7.1f} us {gbs:6.0f} GB/s")
    for tz in (2, 4, 6, 8):
        nacc = 2 if layer[3] else 1
        if tz * layer[2] * nacc > 512 or tz > layer[4]:
            continue
        try:
            us, gbs = run(layer, {"L3D_DWS_IGEMM_MAX": 1 << 20, "L3D_C3_TZ": tz})
            cols.append(f"TZ{tz} {us:7.1f} us {gbs:6.0f}")
        except Exception as e:
            cols.append(f"TZ{tz} n/a")
    print(f"{layer[0]:9s} {layer[1]:3d}->{layer[2]:3d}{'+sc' if layer[3] else '   '} @{layer[4]}^3 | " + " | ".join(cols))
