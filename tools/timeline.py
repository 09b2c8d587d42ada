#!/usr/bin/env python
"""Per-tile timeline of CTA 0 of the implicit-GEMM conv (development tool): python tools/timeline.py Cin Cout sc size B normed"""
import os, sys, ctypes
os.environ["L3D_C3_DEBUG_SKIP"] = str(8 | int(os.environ.get("SKIP", "0")))
os.environ.setdefault("L3D_DWS_IGEMM_MAX", "1000000")
sys.argv = [sys.argv[0]] + sys.argv[1:]
import runpy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
runpy.run_path(os.path.join(ROOT, "tools", "prof_layer.py"), run_name="__main__")
sys.path.insert(0, os.path.join(ROOT, "light-3d-unet-front_b200"))
from light_unet import _native as nv
n = 40
buf = (ctypes.c_longlong * (n * 8))()
nv.lib().l3d_conv3_debug_read(buf, n * 8)
t0 = buf[0]
names = ["w:tma", "w:Afree", "w:act", "w:acc", "w:epi", "i:Aok", "i:accfree", "i:mma"]
print("item " + " ".join(f"{x:>9s}" for x in names) + "   (clocks since first box landed)")
for i in range(n):
    row = [buf[i * 8 + k] for k in range(8)]
    print(f"{i:4d} " + " ".join(f"{(v - t0) if v else 0:9d}" for v in row))
