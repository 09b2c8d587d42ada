"""Residual merge (+ MaxPool3d(2)) kernels (unet3d.py:95-100,110-116): the column-per-thread and cell-per-thread mappings
against a plain PyTorch fp32 reference of the same op, and against each other bit for bit, on even and odd extents."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _run(nv, t2, r, n2, nr, dims, C, pool, rank1_w=None):
    B = t2.shape[0]
    D, H, W = dims
    out = torch.zeros(B, D, H, W, 2 * C, dtype=t2.dtype, device=DEV)          # written as the upper half of a concat buffer
    pooled = torch.zeros(B, D // 2, H // 2, W // 2, C, dtype=t2.dtype, device=DEV) if pool else None
    st = nv.stream_ptr(torch.device(DEV))
    if rank1_w is None:
        nv.call("l3d_merge_fwd", nv.act(t2), n2, nv.act(r), nr, B, D, H, W, 0.01, nv.act(out, C, C), nv.act(pooled), None, None, 0, None, None, st)
    else:
        nv.call("l3d_merge_fwd_rank1", nv.act(t2), n2, nv.act(r), nv.ptr(rank1_w), nr, B, D, H, W, 0.01, nv.act(out, C, C), nv.act(pooled), st)
    torch.cuda.synchronize()
    return out[..., C:].clone(), pooled


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
@pytest.mark.parametrize("dims,C", [((8, 8, 8), 16), ((7, 9, 11), 16), ((6, 5, 10), 32), ((5, 6, 7), 64), ((4, 4, 6), 128)])
def test_merge_mappings_agree_and_match_torch(dims, C, dtype):
    from light_unet import _native as nv
    torch.manual_seed(C + dims[2])
    B = 3
    D, H, W = dims
    vox = D * H * W
    t2 = torch.randn(B, D, H, W, C, device=DEV).to(dtype)
    r = torch.randn(B, D, H, W, C, device=DEV).to(dtype)
    g2, b2 = torch.rand(C, device=DEV) + 0.5, torch.randn(C, device=DEV) * 0.1
    gr, br = torch.rand(C, device=DEV) + 0.5, torch.randn(C, device=DEV) * 0.1

    def stats_of(t):
        f = t.double()
        return torch.stack([f.sum(dim=(1, 2, 3)), (f * f).sum(dim=(1, 2, 3))]).contiguous()
    s2, sr = stats_of(t2), stats_of(r)
    n2 = nv.norm(s2, g2, b2, None, 1e-5, 1.0, vox)
    nr = nv.norm(sr, gr, br, None, 1e-5, 1.0, vox)
    res = {}
    for mode in ("0", "1", "-1"):          # 0: column per thread (32-byte vectors in fp16); 1: cell per thread; -1: column, 16-byte vectors
        os.environ["L3D_MERGE_CELL"] = mode
        nv.lib().l3d_env_refresh()
        try:
            res[mode] = _run(nv, t2, r, n2, nr, dims, C, True)
        finally:
            os.environ.pop("L3D_MERGE_CELL", None)
            nv.lib().l3d_env_refresh()
    for mode in ("1", "-1"):
        assert torch.equal(res["0"][0], res[mode][0]) and torch.equal(res["0"][1], res[mode][1])
    # PyTorch fp32 reference: InstanceNorm3d(affine) of both tensors, add, LeakyReLU(0.01), MaxPool3d(2)
    a = F.instance_norm(t2.float().permute(0, 4, 1, 2, 3), weight=g2, bias=b2, eps=1e-5)
    b = F.instance_norm(r.float().permute(0, 4, 1, 2, 3), weight=gr, bias=br, eps=1e-5)
    want = F.leaky_relu(a + b, 0.01)
    tol = 2e-5 if dtype == torch.float32 else 4e-3
    got = res["0"][0].float().permute(0, 4, 1, 2, 3)
    assert (got - want).abs().max() <= tol * max(1.0, want.abs().max().item())
    wantp = F.max_pool3d(got, 2)                                   # the pool sees the stored values
    assert torch.equal(res["0"][1].float().permute(0, 4, 1, 2, 3), wantp)


@pytest.mark.parametrize("dims", [(8, 8, 8), (7, 9, 11)])
def test_rank1_merge_mappings_agree(dims):
    """First block of a single-channel image: the shortcut r = sc (x) x is evaluated on the fly from x."""
    from light_unet import _native as nv
    torch.manual_seed(5)
    B, C = 2, 16
    D, H, W = dims
    vox = D * H * W
    t2 = torch.randn(B, D, H, W, C, device=DEV).to(torch.float16)
    x = torch.randn(B, D, H, W, 1, device=DEV).to(torch.float16)
    scw = torch.randn(C, device=DEV)
    g, b = torch.ones(C, device=DEV), torch.zeros(C, device=DEV)
    f2 = t2.double()
    s2 = torch.stack([f2.sum(dim=(1, 2, 3)), (f2 * f2).sum(dim=(1, 2, 3))]).contiguous()
    rr = x.double() * scw.double()
    sr = torch.stack([rr.sum(dim=(1, 2, 3)), (rr * rr).sum(dim=(1, 2, 3))]).contiguous()
    n2 = nv.norm(s2, g, b, None, 1e-5, 1.0, vox)
    nr = nv.norm(sr, g, b, None, 1e-5, 1.0, vox)
    res = {}
    for mode in ("0", "-1", "1"):          # 0: column per thread, 32-byte accesses; -1: column per thread, 16-byte vectors; 1: cell per thread
        os.environ["L3D_MERGE_CELL"] = mode
        nv.lib().l3d_env_refresh()
        try:
            res[mode] = _run(nv, t2, x, n2, nr, dims, C, True, rank1_w=scw)
        finally:
            os.environ.pop("L3D_MERGE_CELL", None)
            nv.lib().l3d_env_refresh()
    for mode in ("-1", "1"):
        assert torch.equal(res["0"][0], res[mode][0]) and torch.equal(res["0"][1], res[mode][1])
    a = F.instance_norm(t2.float().permute(0, 4, 1, 2, 3), eps=1e-5)
    bb = F.instance_norm(rr.float().permute(0, 4, 1, 2, 3), eps=1e-5)
    want = F.leaky_relu(a + bb, 0.01)
    got = res["0"][0].float().permute(0, 4, 1, 2, 3)
    assert (got - want).abs().max() <= 4e-3 * max(1.0, want.abs().max().item())


@pytest.mark.parametrize("nvox_dims", [(8, 8, 8), (5, 7, 9)])
def test_head16_voxel_per_thread_matches_half_voxel_kernel(nvox_dims):
    """Residual merge + 1x1x1 head + sigmoid for C = 16 (unet3d.py:201-202,220-221): the 32-byte-access kernel against the
    half-voxel kernel bit for bit (same summation order), and against PyTorch."""
    from light_unet import _native as nv
    torch.manual_seed(11)
    B, C = 3, 16
    D, H, W = nvox_dims
    vox = D * H * W
    t2 = torch.randn(B, D, H, W, C, device=DEV).to(torch.float16)
    r = torch.randn(B, D, H, W, C, device=DEV).to(torch.float16)
    g, b = torch.rand(C, device=DEV) + 0.5, torch.randn(C, device=DEV) * 0.1

    def stats_of(t):
        f = t.double()
        return torch.stack([f.sum(dim=(1, 2, 3)), (f * f).sum(dim=(1, 2, 3))]).contiguous()
    s2, sr = stats_of(t2), stats_of(r)
    n2 = nv.norm(s2, g, b, None, 1e-5, 1.0, vox)
    nr = nv.norm(sr, g, b, None, 1e-5, 1.0, vox)
    hw, hb = torch.randn(1, C, device=DEV) / 4, torch.randn(1, device=DEV)
    st = nv.stream_ptr(torch.device(DEV))
    res = {}
    for mode in ("1", "0"):
        os.environ["L3D_MERGE_256"] = mode
        nv.lib().l3d_env_refresh()
        try:
            out = torch.zeros(B, D, H, W, C, dtype=torch.float16, device=DEV)
            prob = torch.zeros(B, 1, D, H, W, device=DEV)
            logits = torch.zeros(B, 1, D, H, W, device=DEV)
            nv.call("l3d_merge_fwd", nv.act(t2), n2, nv.act(r), nr, B, D, H, W, 0.01, nv.act(out), nv.act(None), nv.ptr(hw), nv.ptr(hb), 1,
                    nv.ptr(prob), nv.ptr(logits), st)
            torch.cuda.synchronize()
            res[mode] = (out, prob, logits)
        finally:
            os.environ.pop("L3D_MERGE_256", None)
            nv.lib().l3d_env_refresh()
    for a, bb in zip(res["1"], res["0"]):
        assert torch.equal(a, bb)
    x1 = F.instance_norm(t2.float().permute(0, 4, 1, 2, 3), weight=g, bias=b, eps=1e-5)
    x2 = F.instance_norm(r.float().permute(0, 4, 1, 2, 3), weight=g, bias=b, eps=1e-5)
    act = F.leaky_relu(x1 + x2, 0.01)
    want = F.conv3d(act, hw.view(1, C, 1, 1, 1), hb)
    assert (res["1"][2] - want).abs().max() < 2e-2          # fp16 storage of the activations the head reads
    assert torch.equal(res["1"][1], torch.sigmoid(res["1"][2])) or (res["1"][1] - torch.sigmoid(res["1"][2])).abs().max() < 1e-6
