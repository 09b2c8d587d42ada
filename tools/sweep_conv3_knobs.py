#!/usr/bin/env python
"""Development aid: the seven implicit-GEMM conv launches of an inference step (325 windows, fp16 storage) under the tuning
knobs of l3d_conv3_tc.cu -- re-checked after the switch from bf16 to fp16 storage (identity-norm inputs are now copied
into the operand tile without conversion).      python tools/sweep_conv3_knobs.py [B]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "light-3d-unet-front_b200"))
from light_unet import _native as nv

import json
B = int(sys.argv[1]) if len(sys.argv) > 1 else 325
DEV = torch.device("cuda:0")
LAYERS = [("init.c2*", 16, 16, False, 48, True), ("up3.c1", 32, 16, True, 48, False), ("up3.c2", 16, 16, False, 48, True),
          ("down1.c1", 16, 32, True, 24, False), ("down1.c2", 32, 32, False, 24, True), ("up2.c1", 64, 32, True, 24, False),
          ("up2.c2", 32, 32, False, 24, True)]
KNOBS = [{}, {"L3D_C3_WARPS": 8}, {"L3D_C3_LOADER": 0}, {"L3D_C3_LOADER": 2, "L3D_C3_WARPS": 12}, {"L3D_C3_TZ": 4}, {"L3D_C3_TZ": 6},
         {"L3D_C3_TZ": 4, "L3D_C3_WARPS": 8}, {"L3D_C3_SETS": 1}, {"L3D_C3_NRAW": 1}, {"L3D_C3_NRAW": 2}]
if len(sys.argv) > 2:                      # python tools/sweep_conv3_knobs.py 325 '[{"L3D_C3_PROD": 0}, {}]'
    KNOBS = json.loads(sys.argv[2])


def run(layer, env, iters=5):
    name, Cin, Cout, sc, S, normed = layer
    torch.manual_seed(0)
    x = torch.randn(B, S, S, S, Cin, device=DEV).to(torch.float16)
    vox = S ** 3
    stats = torch.zeros(2, B, Cin, dtype=torch.float64, device=DEV)
    stats[1] = vox
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
    nv.refresh_env()
    try:
        def go():
            nv.call("l3d_dwpw_fwd", nv.act(x), xn, B, S, S, S, nv.ptr(dw), nv.ptr(pw), nv.ptr(scw), nv.act(t), nv.ptr(ts),
                    nv.act(r), nv.ptr(rs) if sc else None, nv.act(None), st)
        for _ in range(2):
            go()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            go()
        e1.record()
        torch.cuda.synchronize()
        kern = nv.lib().l3d_last_kernel().decode()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        nv.refresh_env()
    return e0.elapsed_time(e1) / iters * 1e3, kern


print(f"B={B}; microseconds per launch (* init.c2 runs through the rank-1 entry point in the real step: timed here with a stored input)")
print(f"{'knobs':44s}" + "".join(f"{l[0]:>10s}" for l in LAYERS))
for env in KNOBS:
    row = []
    for layer in LAYERS:
        try:
            us, kern = run(layer, env)
            row.append(f"{us:9.0f}{'' if kern == 'conv3_tc_kernel' else '!'}")
        except Exception as e:
            row.append("      n/a")
    print(f"{str(env):44s}" + " ".join(row))
