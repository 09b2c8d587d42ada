"""the module-by-module overlay of this package on a reference checkout (l3d_overlay.py).  Needs
the reference tree (present in the build container only; skipped on the GPU box), runs in a subprocess because the
overlay must be installed before the first `import light_unet`."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"

SCRIPT = r'''
import sys, types
sys.dont_write_bytecode = True
sys.modules.setdefault("nibabel", types.ModuleType("nibabel"))       # the reference's datasets import nibabel at module level
try:
    import torch.utils.tensorboard                                    # trainer.py:12
except Exception:
    tb = types.ModuleType("torch.utils.tensorboard"); tb.SummaryWriter = object
    sys.modules["torch.utils.tensorboard"] = tb
sys.path.insert(0, sys.argv[1])
import l3d_overlay
l3d_overlay.install(sys.argv[2])
from light_unet.core.trainer import Trainer                          # the REFERENCE's trainer
import light_unet.core.trainer as tr, light_unet.models.unet3d as un, light_unet.models.metrics as me, light_unet.utils as ut
import light_unet.datasets.loader as lo, light_unet.core.inferencer as inf, light_unet.models as models
ours, ref = sys.argv[1], sys.argv[2]
assert tr.__file__.startswith(ref) and lo.__file__.startswith(ref), (tr.__file__, lo.__file__)
for m in (un, me, ut, inf):
    assert m.__file__.startswith(ours), m.__file__
# the names the reference trainer bound at import time are the native ones
assert tr.Lightweight3DUNet is un.Lightweight3DUNet and tr.calculate_metrics is me.calculate_metrics
assert tr.sliding_window_inference_3d is ut.sliding_window_inference_3d
# the reference's package-level re-exports (models/__init__.py:6-41) still work, datasets included
assert models.PatchDataset.__module__ == "light_unet.datasets.patch_dataset" and models.Lightweight3DUNet is un.Lightweight3DUNet
from light_unet.datasets.device_patches import DevicePatchSampler   # our addition sits beside the reference's datasets
m = un.Lightweight3DUNet()
assert m.count_parameters()["total"] == 217228
print("overlay ok")
'''


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "light_unet")), reason="needs the reference checkout")
def test_reference_trainer_imports_through_the_overlay():
    r = subprocess.run([sys.executable, "-B", "-c", SCRIPT, os.path.join(ROOT, "light-3d-unet-front_b200"), REF],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "overlay ok" in r.stdout, r.stdout + r.stderr
