#!/usr/bin/env python
"""Per-kernel timing of one training step (development tool): python tools/time_train.py [B] [dtype]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "light-3d-unet-front_b200"))
from light_unet import _native as nv
from light_unet.models import Lightweight3DUNet, FocalTverskyLoss
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dtype = sys.argv[2] if len(sys.argv) > 2 else "f32"
torch.manual_seed(0)
m = Lightweight3DUNet(dropout_p=0.1).cuda().set_compute_dtype(dtype).train()
opt = torch.optim.AdamW(m.parameters(), lr=1e-4, weight_decay=1e-5, fused=True)
loss_fn = FocalTverskyLoss()
x = torch.rand(B, 1, 48, 48, 48, device="cuda")
t = (torch.rand_like(x) > 0.98).float()
def step():
    opt.zero_grad()
    loss = loss_fn(m(x), t)
    loss.backward()
    opt.step()
    return loss
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"train step B={B} {dtype}: {ms:.3f} ms -> {B / ms * 1e3:.0f} patches/s")
nv.TIMER.start()
for _ in range(3): step()
rec = nv.TIMER.stop()
tot = sum(v[1] for v in rec.values())
byname = {}
for (name, tag), (n, tt, _b) in rec.items():
    a = byname.setdefault(name, [0, 0.0]); a[0] += n; a[1] += tt
for name, (n, tt) in sorted(byname.items(), key=lambda kv: -kv[1][1]):
    print(f" == {name:22s} n={n//3:3d}/step {tt/3:8.3f} ms/step {100*tt/tot:5.1f}%")
for (name, tag), (n, tt, _b) in sorted(rec.items(), key=lambda kv: -kv[1][1])[:28]:
    print(f"   {name:20s} {tag:14s} n={n:3d} avg {tt / n * 1e3:9.1f} us  {100 * tt / tot:5.1f}%")
print(f"   sum of libl3d kernels {tot / 3:.3f} ms per step")
