#!/usr/bin/env python
"""Development aid: which kernel call of the backward pass first produces run-to-run differences?  The same training step is
run `reps` times on identical inputs; after every libl3d call of the backward pass the gradient buffers of the workspace are
check-summed (double sum and sum of squares), and the per-call checksums of each repetition are compared with repetition 0.
    python tools/diag_bwd_determinism.py [B] [S] [reps]"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "light-3d-unet-front_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import synth, unet_ref
from light_unet import _native as nv, engine
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
m = Lightweight3DUNet(dropout_p=0.0)
m.load_state_dict(unet_ref.to_torch(sd_np))
m = m.to(dev).set_compute_dtype("f32").train()

state = {"ws": None, "log": None, "on": False}
orig_backward = engine.UNetPlan.backward
orig_call = nv.call

def bufs(ws):
    out = {}
    for name in ("g_cat", "g_pooled", "g_out", "gz", "gy", "gu"):
        d = getattr(ws, name, None)
        if isinstance(d, dict):
            for k, v in d.items():
                if torch.is_tensor(v): out[f"{name}[{k}]"] = v
    out["red"] = ws.red
    return out

def call(name, *a, **kw):
    r = orig_call(name, *a, **kw)
    if state["on"]:
        ws = state["ws"]
        snap = {}
        for k, v in bufs(ws).items():
            vv = torch.nan_to_num(v.double(), nan=0.0, posinf=0.0, neginf=0.0)
            snap[k] = (float(vv.sum()), float((vv * vv).sum()))
        state["log"].append((name, nv.TIMER.tag, snap))
    return r

fwd_logs = []
def fwd_bufs(ws):
    out = {}
    for bn, d in ws.blocks.items():
        for k, v in d.items():
            if torch.is_tensor(v): out[f"{bn}.{k}"] = v
    for name in ("cat", "pooled"):
        d = getattr(ws, name, None)
        if isinstance(d, dict):
            for k, v in d.items():
                if torch.is_tensor(v): out[f"{name}[{k}]"] = v
    out["stats"] = ws.stats
    return out

def backward(self, P, ws, g_prob):
    snap = {}
    for k, v in fwd_bufs(ws).items():
        vv = torch.nan_to_num(v.double(), nan=0.0, posinf=0.0, neginf=0.0)
        snap[k] = (float(vv.sum()), float((vv * vv).sum()), vv.clone() if os.environ.get("KEEP") else None)
    fwd_logs.append(snap)
    fill = os.environ.get("FILL", "zero")
    if fill != "none":
        for v in bufs(ws).values():
            if v.dtype != torch.float64: v.fill_(0.0 if fill == "zero" else 1.0)
    state["ws"], state["on"] = ws, True
    try:
        return orig_backward(self, P, ws, g_prob)
    finally:
        state["on"] = False

nv.call = call
engine.nv.call = call
engine.UNetPlan.backward = backward
logs = []
for r in range(reps):
    state["log"] = []
    m.zero_grad(set_to_none=True)
    loss = loss_fn(m(xs.float()), ts)
    loss.backward()
    torch.cuda.synchronize()
    logs.append(state["log"])
print(f"{len(logs[0])} backward calls per step")
for r in range(1, reps):
    d = {}
    for k in fwd_logs[0]:
        a, b = fwd_logs[0][k], fwd_logs[r][k]
        if a[2] is not None:
            d[k] = float((a[2] - b[2]).abs().max() / (a[2].abs().max() + 1e-300))
        else:
            d[k] = abs(a[1] - b[1]) / max(abs(a[1]), 1e-300)
    top = sorted(d.items(), key=lambda kv: -kv[1])[:6]
    print(f"rep {r}: forward tensors vs rep 0 ({'max-abs / max' if os.environ.get('KEEP') else 'sum of squares'}): " + "  ".join(f"{k}:{v:.1e}" for k, v in top))
for r in range(1, reps):
    first = None
    for i, ((n0, t0, s0), (n1, t1, s1)) in enumerate(zip(logs[0], logs[r])):
        worst = 0.0; wk = None
        for k in s0:
            a, b = s0[k], s1[k]
            e = abs(a[1] - b[1]) / max(abs(a[1]), 1e-300)
            if e > worst: worst, wk = e, k
        if worst > 1e-5 and first is None:
            first = (i, n0, t0, wk, worst)
    print(f"rep {r}: first call whose buffers differ by > 1e-5 (relative, sum of squares): {first}")
    if first:
        i = first[0]
        for j in range(max(0, i - 2), min(len(logs[0]), i + 3)):
            (n0, t0, s0), (_, _, s1) = logs[0][j], logs[r][j]
            d = {k: abs(s0[k][1] - s1[k][1]) / max(abs(s0[k][1]), 1e-300) for k in s0}
            top = sorted(d.items(), key=lambda kv: -kv[1])[:3]
            print(f"     call {j:3d} {n0:28s} {t0:12s} " + "  ".join(f"{k}:{v:.1e}" for k, v in top))
