"""Development aid: repeats the worst fp32 training-gradient parity case in one process and prints the distribution of the
worst per-tensor error (see the tolerance note in tests/test_gpu_train.py)."""
import sys, os, io, contextlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "light-3d-unet-front_b200"))
import test_gpu_train as T
vals = []
for i in range(60):
    buf = io.StringIO()
    try:
        with contextlib.redirect_stdout(buf):
            T.test_train_step_gradients("dws_20_pad", "f32", 1.0, 1e-4)
    except AssertionError as e:
        print("assert", str(e)[:200])
    out = buf.getvalue().strip().splitlines()[-1] if buf.getvalue().strip() else ""
    vals.append(out.split("worst grad rel-L2")[-1].strip())
from collections import Counter
print(Counter(v.split("(")[-1] for v in vals))
print(sorted(vals, key=lambda s: float(s.split()[0]))[-12:])
print(sorted(vals, key=lambda s: float(s.split()[0]))[:3])
