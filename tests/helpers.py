"""Shared helpers for the parity tests (fixture loading, seeded synthetic
weights/inputs).  Uses oracle/ as the checker only."""
import json
import os

import numpy as np
import torch

from oracle import synth, unet_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

UNET_CASES = ["dws_16", "dws_24", "dws_20_pad", "dws_small_enc", "grouped_16", "dense_16", "dws_48_c1"]


def load_unet_case(name):
    z = np.load(os.path.join(GOLDEN, f"unet_{name}.npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    cfg = unet_ref.UNetCfg(in_channels=meta["in_channels"], out_channels=meta["out_channels"],
                           encoder_channels=tuple(meta["encoder_channels"]),
                           use_depthwise_separable=meta["dws"], use_grouped=meta["grouped"],
                           groups=meta["groups"], dropout_p=meta["dropout_p"])
    sd_np = synth.synth_state_dict(unet_ref.param_shapes(cfg), meta["wseed"])
    x, t = synth.synth_patches(meta["batch"], tuple(meta["size"]), meta["xseed"])
    return z, meta, cfg, sd_np, x, t


def sub(a, s):
    return a if not s else a[:, :, ::s, ::s, ::s]
