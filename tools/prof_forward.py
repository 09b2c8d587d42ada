#!/usr/bin/env python
"""Minimal forward driver for ncu captures: python tools/prof_forward.py [B] [dtype] [iters]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "light-3d-unet-front_b200"))
from light_unet.models import Lightweight3DUNet
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dtype = sys.argv[2] if len(sys.argv) > 2 else "f16"
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 2
torch.manual_seed(0)
m = Lightweight3DUNet(dropout_p=0.0).cuda().set_compute_dtype(dtype).eval()
x = torch.rand(B, 1, 48, 48, 48, device="cuda")
with torch.no_grad():
    for _ in range(iters):
        m(x)
torch.cuda.synchronize()
print("done")
