#!/usr/bin/env python
"""Development aid: per-object SASS opcode histogram of the Blackwell-specific instructions (tcgen05 MMA / TMEM / TMA /
mbarrier), from the objects build.py leaves under light-3d-unet-front_b200/build/.  Output kept in profiles/.

    python tools/sass_histogram.py > profiles/r02_sass_histogram.txt
"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["UTCHMMA", "UTCQMMA", "UTCMMA", "UTCCP", "LDTM", "STTM", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "LDGSTS", "HMMA", "HFMA2", "FFMA", "DFMA", "DMUL", "DADD",
        "ATOM", "RED", "ATOMS", "LDG", "STG", "LDS", "STS", "SHFL", "BAR", "ELECT", "R2UR"]

print("SASS opcode counts per object (cuobjdump -sass, sm_100a); UTCHMMA = tcgen05.mma kind::f16, LDTM = tcgen05.ld, UTMALDG = TMA tensor load, "
      "SYNCS = mbarrier ops")
for obj in sorted(glob.glob(os.path.join(ROOT, "light-3d-unet-front_b200", "build", "*.o"))):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    cnt = collections.Counter()
    kernels = re.findall(r"Function : (\S+)", out)
    for line in out.splitlines():
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        for k in KEYS:
            if op == k or op.startswith(k + "."):
                cnt[k] += 1
                break
    print(f"\n{os.path.basename(obj)}: {len(kernels)} kernels")
    print("  " + "  ".join(f"{k}={cnt[k]}" for k in KEYS if cnt[k]))
