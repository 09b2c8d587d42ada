"""Development aid: one eval forward of the default model through the rank-1 first-block path (run under compute-sanitizer)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "light-3d-unet-front_b200"))
from light_unet.models import Lightweight3DUNet
S = int(sys.argv[1]) if len(sys.argv) > 1 else 16
m = Lightweight3DUNet(dropout_p=0.0).to("cuda:0").set_compute_dtype("f16").eval()
x = torch.rand(2, 1, S, S, S, device="cuda:0")
with torch.no_grad():
    y = m(x)
torch.cuda.synchronize()
print("ok", float(y.mean()))
