#!/usr/bin/env python
"""Development aid: the residual-merge kernels of the 48^3 level alone at the inference batch (325 windows), GB/s of
algorithmic bytes.    python tools/time_merge.py [B]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "light-3d-unet-front_b200"))
from light_unet import _native as nv

B = int(sys.argv[1]) if len(sys.argv) > 1 else 325
DEV = torch.device("cuda:0")
S, C = 48, 16
vox = S ** 3
torch.manual_seed(0)
t2 = torch.randn(B, S, S, S, C, device=DEV).to(torch.float16)
r = torch.randn(B, S, S, S, C, device=DEV).to(torch.float16)
x1 = torch.randn(B, S, S, S, 1, device=DEV).to(torch.float16)
cat = torch.zeros(B, S, S, S, 2 * C, dtype=torch.float16, device=DEV)
pooled = torch.empty(B, S // 2, S // 2, S // 2, C, dtype=torch.float16, device=DEV)
prob = torch.empty(B, 1, S, S, S, dtype=torch.float32, device=DEV)


def stats_of(t):
    f = t.float()
    return torch.stack([f.sum(dim=(1, 2, 3)), (f * f).sum(dim=(1, 2, 3))]).double().contiguous()


g, b = torch.ones(C, device=DEV), torch.zeros(C, device=DEV)
n2 = nv.norm(stats_of(t2), g, b, None, 1e-5, 1.0, vox)
nr = nv.norm(stats_of(r), g, b, None, 1e-5, 1.0, vox)
scw = torch.randn(C, device=DEV)
hw, hb = torch.randn(1, C, device=DEV), torch.zeros(1, device=DEV)
st = nv.stream_ptr(DEV)
mb = 2 * B * vox * C


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


cases = {
    "head16 (t2 + r -> prob)": (lambda: nv.call("l3d_merge_fwd", nv.act(t2), n2, nv.act(r), nr, B, S, S, S, 0.01, nv.act(None), nv.act(None),
                                                nv.ptr(hw), nv.ptr(hb), 1, nv.ptr(prob), None, st), 2 * mb + 4 * B * vox),
    "merge (t2 + r -> concat half + pooled)": (lambda: nv.call("l3d_merge_fwd", nv.act(t2), n2, nv.act(r), nr, B, S, S, S, 0.01, nv.act(cat, C, C),
                                                               nv.act(pooled), None, None, 0, None, None, st), 3 * mb + mb // 8),
    "merge rank-1 (t2 + x -> concat half + pooled)": (lambda: nv.call("l3d_merge_fwd_rank1", nv.act(t2), n2, nv.act(x1), nv.ptr(scw), nr, B, S, S, S, 0.01,
                                                                      nv.act(cat, C, C), nv.act(pooled), st), 2 * mb + mb // 8 + mb // C),
}
xin = torch.randn(B, S // 2, S // 2, S // 2, 2 * C, device=DEV).to(torch.float16)
ctw = (torch.randn(2 * C, C, 2, 2, 2, device=DEV) / 6).contiguous()
ctb = torch.zeros(C, device=DEV)
cases["convT 32 -> 16 into the concat half (24^3 -> 48^3)"] = (
    lambda: nv.call("l3d_convt_fwd", nv.act(xin), B, S // 2, S // 2, S // 2, nv.ptr(ctw), nv.ptr(ctb), nv.act(cat, 0, C), S, S, S, 0, 0, 0, st),
    2 * B * (vox // 8) * 2 * C + mb)
dense = torch.zeros(B, S, S, S, C, dtype=torch.float16, device=DEV)
cases["merge rank-1 -> DENSE 16-channel tensor + pooled"] = (
    lambda: nv.call("l3d_merge_fwd_rank1", nv.act(t2), n2, nv.act(x1), nv.ptr(scw), nr, B, S, S, S, 0.01, nv.act(dense), nv.act(pooled), st),
    2 * mb + mb // 8 + mb // C)
cases["merge -> DENSE 16-channel tensor + pooled"] = (
    lambda: nv.call("l3d_merge_fwd", nv.act(t2), n2, nv.act(r), nr, B, S, S, S, 0.01, nv.act(dense), nv.act(pooled), None, None, 0, None, None, st),
    3 * mb + mb // 8)
cases["convT 32 -> 16 into a DENSE 16-channel tensor"] = (
    lambda: nv.call("l3d_convt_fwd", nv.act(xin), B, S // 2, S // 2, S // 2, nv.ptr(ctw), nv.ptr(ctb), nv.act(dense), S, S, S, 0, 0, 0, st),
    2 * B * (vox // 8) * 2 * C + mb)
print("knobs:", {k: v for k, v in os.environ.items() if k.startswith("L3D_")})
for name, (fn, nbytes) in cases.items():
    us = timeit(fn)
    print(f"{name:48s} {us:8.1f} us  {nbytes / us / 1e3:7.0f} GB/s")
