"""Development aid: fp16-storage forward parity against the oracle at other patch sizes (64^3, 96^3, 50^3, 40^3)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "light-3d-unet-front_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import synth, unet_ref
from light_unet.models import Lightweight3DUNet
from light_unet import _native as nv
cfg = unet_ref.UNetCfg(dropout_p=0.0)
sd_np = synth.synth_state_dict(unet_ref.param_shapes(cfg), 3)
m = Lightweight3DUNet(dropout_p=0.0); m.load_state_dict(unet_ref.to_torch(sd_np)); m = m.to("cuda:0").set_compute_dtype("f16").eval()
for S, B in ((64, 2), (96, 1), (50, 2), (40, 3)):
    x, _ = synth.synth_patches(B, (S, S, S), 7)
    with torch.no_grad():
        y = m(torch.from_numpy(x).cuda()).cpu().numpy()
        t0 = time.time(); ref = unet_ref.forward(unet_ref.to_torch(sd_np), torch.from_numpy(x), cfg).numpy(); dt = time.time() - t0
    e = float(np.sqrt(((y - ref) ** 2).sum() / (ref ** 2).sum()))
    print(f"{S}^3 x{B}: rel-L2 {e:.3e} (oracle {dt:.1f} s) last kernel {nv.lib().l3d_last_kernel().decode()}")
    bad = e >= 1e-2 or locals().get("bad", False)
print("ok" if not bad else "ABOVE 1e-2")
