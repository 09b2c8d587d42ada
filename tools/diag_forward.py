#!/usr/bin/env python
"""Per-block forward diagnostics on the GPU: compares every raw tensor (t1, t2, block output) of the CUDA path with
the oracle's taps, so a failing parity test can be localised to one kernel in a single gpurun call.
(Test tooling: uses oracle/ as the checker.)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "light-3d-unet-front_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import load_unet_case          # noqa: E402
from oracle import unet_ref                 # noqa: E402
from light_unet.models import Lightweight3DUNet   # noqa: E402


def main():
    cases = sys.argv[1:] or ["dws_16", "dws_20_pad", "grouped_16", "dense_16"]
    for name in cases:
        for dtype in ("f32", "f16"):
            z, meta, cfg, sd_np, x, t = load_unet_case(name)
            m = Lightweight3DUNet(in_channels=cfg.in_channels, out_channels=cfg.out_channels,
                                  encoder_channels=list(cfg.encoder_channels),
                                  use_depthwise_separable=cfg.use_depthwise_separable, use_grouped=cfg.use_grouped,
                                  groups=cfg.groups, dropout_p=cfg.dropout_p)
            m.load_state_dict(unet_ref.to_torch(sd_np))
            m = m.cuda().set_compute_dtype(dtype).eval()
            with torch.no_grad():
                y = m(torch.from_numpy(x).cuda())
            torch.cuda.synchronize()
            taps = {}
            with torch.no_grad():
                ref = unet_ref.forward(unet_ref.to_torch(sd_np), torch.from_numpy(x), cfg, taps=taps)
            ws = list(m._plan._ws.values())[-1]
            print(f"== {name} / {dtype}: final max|dprob| = {(y.cpu() - ref).abs().max().item():.3e}")
            for b in m._plan.blocks:
                buf = ws.blocks[b.name]
                for key in ("t1", "t2", "out"):
                    if key == "out":
                        if b.name in ("init_conv", "down1", "down2"):
                            got = ws.cat[b.level][..., b.cout:]
                        elif "out" in buf:
                            got = buf["out"]
                        else:
                            continue
                    else:
                        got = buf[key]
                    got = got.float().permute(0, 4, 1, 2, 3).cpu()
                    want = taps[f"{b.prefix}.{key}"]
                    err = (got - want).abs().max().item()
                    scale = want.abs().max().item()
                    flag = "" if err <= (2e-4 if dtype == "f32" else 5e-2) * max(scale, 1.0) else "   <-- MISMATCH"
                    print(f"   {b.name:10s} {key:3s} max|d|={err:.3e} (max|ref|={scale:.3e}){flag}")


if __name__ == "__main__":
    main()
