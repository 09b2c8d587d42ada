#!/usr/bin/env python
"""Two whole-volume inference steps (the bench workload) for ncu captures: python tools/prof_infer.py [n_steps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "light-3d-unet-front_b200"))
import torch
import bench
inf = bench.make_inferencer("dws", "f16", torch.device("cuda:0"))
from oracle import synth
vol = torch.from_numpy(synth.synth_volume(bench.VOLUME, seed=42, n_blobs=6)).cuda()
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    prob, boxes = inf.infer_volume(vol, threshold=0.3, return_device=True)
torch.cuda.synchronize()
print("done", len(boxes))
