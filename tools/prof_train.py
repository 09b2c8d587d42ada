#!/usr/bin/env python
"""Minimal training-step driver for ncu captures: python tools/prof_train.py [B] [dtype] [iters]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "light-3d-unet-front_b200"))
from light_unet.models import Lightweight3DUNet, FocalTverskyLoss
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dtype = sys.argv[2] if len(sys.argv) > 2 else "f16"
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 2
torch.manual_seed(0)
m = Lightweight3DUNet(dropout_p=0.1).cuda().set_compute_dtype(dtype).train()
x = torch.rand(B, 1, 48, 48, 48, device="cuda")
t = (torch.rand_like(x) > 0.98).float()
for _ in range(iters):
    m.zero_grad()
    FocalTverskyLoss()(m(x), t).backward()
torch.cuda.synchronize()
print("done")
