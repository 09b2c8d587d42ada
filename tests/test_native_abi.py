"""CPU tests: the C-ABI library loads and exports every symbol include/l3d.h declares, and the host-side
mirror of the reference interface (module tree, state_dict, window grid, Gaussian map, error behaviour)
matches the fixtures generated from the reference.  No kernel is launched here."""
import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch

from helpers import GOLDEN, ROOT
from oracle import unet_ref


def _declared_symbols():
    with open(os.path.join(ROOT, "include", "l3d.h")) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(l3d_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from light_unet import _native
    lib = ctypes.CDLL(_native.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"libl3d.so does not export {name}"
    assert sorted(_native.EXPORTS) == declared
    assert _native.lib().l3d_abi_version() == _native.ABI_VERSION


def test_state_dict_matches_reference_layout():
    from light_unet.models import Lightweight3DUNet
    for kw, cfg in [({}, unet_ref.UNetCfg()),
                    (dict(use_depthwise_separable=False), unet_ref.UNetCfg(use_depthwise_separable=False)),
                    (dict(use_depthwise_separable=False, use_grouped=False),
                     unet_ref.UNetCfg(use_depthwise_separable=False, use_grouped=False)),
                    (dict(encoder_channels=[8, 16, 32, 64]), unet_ref.UNetCfg(encoder_channels=(8, 16, 32, 64)))]:
        m = Lightweight3DUNet(**kw)
        shapes = unet_ref.param_shapes(cfg)
        sd = m.state_dict()
        assert list(sd.keys()) == list(shapes.keys())
        for k, v in sd.items():
            assert tuple(v.shape) == tuple(shapes[k]), k
        assert len(list(m.buffers())) == 0
    m = Lightweight3DUNet()
    assert m.count_parameters() == {"total": 217228, "trainable": 217228}
    assert m.in_channels == 1 and m.out_channels == 1 and m.encoder_channels == [16, 32, 64, 128]


def test_seeded_default_init_equals_reference():
    from light_unet.models import Lightweight3DUNet
    with open(os.path.join(GOLDEN, "default_init.json")) as f:
        fx = json.load(f)
    for tag, kw in [("dws", {}), ("grouped", dict(use_depthwise_separable=False)),
                    ("dense", dict(use_depthwise_separable=False, use_grouped=False))]:
        torch.manual_seed(1234)
        m = Lightweight3DUNet(**kw)
        for k, v in m.state_dict().items():
            s, a = fx[tag][k]
            assert abs(float(v.double().sum()) - s) < 1e-9 and abs(float(v.double().abs().sum()) - a) < 1e-9, (tag, k)


def test_no_cpu_fallback():
    from light_unet.models import Lightweight3DUNet, FocalTverskyLoss
    from light_unet.utils import sliding_window_inference_3d
    from light_unet._native import NativeError
    m = Lightweight3DUNet()
    with pytest.raises(NativeError):
        m(torch.zeros(1, 1, 16, 16, 16))
    with pytest.raises(NativeError):
        FocalTverskyLoss()(torch.rand(8), torch.rand(8))
    with pytest.raises(NativeError):
        sliding_window_inference_3d(np.zeros((16, 16, 16), np.float32), m, (16, 16, 16), device=torch.device("cpu"))
    with pytest.raises(ValueError):
        sliding_window_inference_3d(np.zeros((16, 16), np.float32), m, (16, 16, 16), device=torch.device("cpu"))
    with pytest.raises(ValueError):
        m(torch.zeros(1, 2, 16, 16, 16))


def test_loss_factory_and_asserts():
    from light_unet.models import FocalTverskyLoss, CombinedLoss, DiceLoss, get_loss_function
    assert isinstance(get_loss_function({}), FocalTverskyLoss)
    assert isinstance(get_loss_function({"name": "DiceLoss"}), DiceLoss)
    c = get_loss_function({"use_combined_loss": True})
    assert isinstance(c, CombinedLoss) and c.ftl_weight == 0.8
    f = get_loss_function({"name": "FocalTverskyLoss", "alpha": 0.6, "beta": 0.4, "gamma": 1.5})
    assert (f.alpha, f.beta, f.gamma, f.smooth) == (0.6, 0.4, 1.5, 1e-6)
    with pytest.raises(ValueError):
        get_loss_function({"name": "nope"})
    with pytest.raises(AssertionError):
        FocalTverskyLoss(alpha=0.7, beta=0.7)


def test_window_grid_and_gaussian_match_reference_fixtures():
    from light_unet.utils import window_positions, _get_gaussian_importance_map
    with open(os.path.join(GOLDEN, "window_grid.json")) as f:
        grids = json.load(f)
    for key, exp in grids.items():
        shape, patch, ov = key.split("|")
        assert [list(p) for p in window_positions(eval(shape), eval(patch), float(ov))] == exp, key
    z = np.load(os.path.join(GOLDEN, "gaussian.npz"))
    for patch in [(48, 48, 48), (16, 16, 16), (32, 48, 64), (7, 9, 11)]:
        g = _get_gaussian_importance_map(patch)
        tag = "x".join(map(str, patch))
        assert g.dtype == np.float32
        assert np.array_equal(g[:, patch[1] // 2, patch[2] // 2], z[f"gz_{tag}"])
        assert np.array_equal(g[patch[0] // 2, patch[1] // 2, :], z[f"gx_{tag}"])
        assert g.min() == z[f"gmin_{tag}"]


def test_inferencer_type_error():
    from light_unet.core.inferencer import Inferencer
    with pytest.raises(TypeError):
        Inferencer(42, "x.pth")


def test_split_concat_plan_choice(monkeypatch):
    """The top-level [up | skip] concat is kept as two dense tensors only where l3d_dwpw_fwd2 applies: inference, fp16 storage,
    depthwise-separable up3.conv1 with 16 + 16 input channels (engine.UNetPlan.split_cat0) -- host logic, no GPU needed."""
    import torch
    from light_unet.models.unet3d import Lightweight3DUNet
    mk = lambda enc, dws: Lightweight3DUNet(in_channels=1, out_channels=1, start_channels=enc[0], encoder_channels=enc,
                                            use_depthwise_separable=dws, use_grouped=True, groups=8, dropout_p=0.0)._plan
    monkeypatch.delenv("L3D_SPLIT_CAT", raising=False)
    monkeypatch.delenv("L3D_NO_IGEMM", raising=False)
    p = mk([16, 32, 64, 128], True)
    assert p.split_cat0(torch.float16, False)
    assert not p.split_cat0(torch.float16, True)          # training reads the interleaved buffer in its backward kernels
    assert not p.split_cat0(torch.float32, False)
    assert not mk([16, 32, 64, 128], False).split_cat0(torch.float16, False)      # grouped / dense variants
    assert not mk([32, 64, 128, 256], True).split_cat0(torch.float16, False)      # 32 + 32 channels: not two 16-channel chunks
    monkeypatch.setenv("L3D_SPLIT_CAT", "0")
    assert not p.split_cat0(torch.float16, False)
    monkeypatch.delenv("L3D_SPLIT_CAT")
    monkeypatch.setenv("L3D_NO_IGEMM", "1")
    assert not p.split_cat0(torch.float16, False)
