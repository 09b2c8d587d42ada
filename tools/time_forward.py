#!/usr/bin/env python
"""Ad-hoc per-kernel timing of the forward pass / sliding window (development tool)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "light-3d-unet-front_b200"))
from light_unet import _native as nv                      # noqa: E402
from light_unet.models import Lightweight3DUNet            # noqa: E402
from light_unet.utils import sliding_window_device         # noqa: E402
import light_unet.utils as lu                              # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    dtype = sys.argv[2] if len(sys.argv) > 2 else "f16"
    kw = {}
    if len(sys.argv) > 3 and sys.argv[3] == "dense":
        kw = dict(use_depthwise_separable=False, use_grouped=False)
    if len(sys.argv) > 3 and sys.argv[3] == "grouped":
        kw = dict(use_depthwise_separable=False, use_grouped=True)
    torch.manual_seed(0)
    m = Lightweight3DUNet(dropout_p=0.0, **kw).cuda().set_compute_dtype(dtype).eval()
    x = torch.rand(B, 1, 48, 48, 48, device="cuda")
    with torch.no_grad():
        for _ in range(3):
            m(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            m(x)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"forward B={B} {dtype} {kw}: {ms:.3f} ms  -> {B / ms * 1e3:.0f} patches/s, {ms / B * 1e3:.1f} us/patch")
        nv.TIMER.start()
        for _ in range(5):
            m(x)
        rec = nv.TIMER.stop()
        tot = sum(v[1] for v in rec.values())
        for (name, tag), (n, t, _b) in sorted(rec.items(), key=lambda kv: -kv[1][1]):
            print(f"   {name:18s} {tag:10s} n={n:3d} avg {t / n * 1e3:9.1f} us  {100 * t / tot:5.1f}%")
        print(f"   sum of kernels {tot / 5:.3f} ms per forward")
        if len(sys.argv) > 3:
            return
        # sliding window on the C3 volume
        vol = torch.rand(128, 128, 320, device="cuda")
        for wb in (32, 65, 109, 163, 325):
            lu.WINDOW_BATCH = wb
            for _ in range(2):
                sliding_window_device(vol, m, (48, 48, 48), 0.5, True, threshold=0.3)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(3):
                sliding_window_device(vol, m, (48, 48, 48), 0.5, True, threshold=0.3)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            print(f"sliding window 128x128x320 window_batch={wb}: {ms:.2f} ms -> {128 * 128 * 320 / ms * 1e3 / 1e6:.1f} Mvoxel/s")


if __name__ == "__main__":
    main()
