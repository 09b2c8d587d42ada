"""Layer-level GPU parity of the tensor-core pointwise backward (csrc/l3d_bwd_tc.cu) through the C-ABI entry point
l3d_pw_bwd: against a float64 torch restatement of the same math (InstanceNorm backward on load, dgrad, wgrad) and
against the CUDA-core kernel it replaces (L3D_NO_TC_BWD=1) on identical f16-stored inputs."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
EPS, SLOPE = 1e-5, 0.01


def _case(N, dims, Cg, Cu, has_nt, u_norm, seed, store=torch.float16):
    g = torch.Generator().manual_seed(seed)
    D, H, W = dims
    vox = D * H * W
    gz = (torch.randn(N, D, H, W, Cg, generator=g) * 1e-3 + 2e-3).float()          # gradient with a common offset
    t = torch.randn(N, D, H, W, Cg, generator=g).to(store)
    u = torch.randn(N, D, H, W, Cu, generator=g).to(store)
    w = (torch.randn(Cg, Cu, generator=g) / np.sqrt(Cu)).float()
    gam_t, bet_t = torch.rand(Cg, generator=g) + 0.5, torch.randn(Cg, generator=g) * 0.2
    gam_u, bet_u = torch.rand(Cu, generator=g) + 0.5, torch.randn(Cu, generator=g) * 0.2

    def stats_of(x):
        xf = x.double()
        return torch.stack([xf.sum(dim=(1, 2, 3)), (xf * xf).sum(dim=(1, 2, 3))])

    st_t, st_u = stats_of(t), stats_of(u)
    # ---- float64 reference
    gzd, td, ud = gz.double(), t.double(), u.double()
    if has_nt:
        mean = st_t[0] / vox
        rstd = 1.0 / torch.sqrt(st_t[1] / vox - mean * mean + EPS)
        xhat = (td - mean[:, None, None, None, :]) * rstd[:, None, None, None, :]
        red = torch.stack([gzd.sum(dim=(1, 2, 3)), (gzd * xhat).sum(dim=(1, 2, 3))])
        k1, k2 = red[0] / vox, red[1] / vox
        gr = gam_t.double() * rstd
        g_t = gr[:, None, None, None, :] * (gzd - k1[:, None, None, None, :] - xhat * k2[:, None, None, None, :])
    else:
        red = torch.zeros(2, N, Cg, dtype=torch.float64)
        g_t = gzd
    if u_norm:
        mean_u = st_u[0] / vox
        rstd_u = 1.0 / torch.sqrt(st_u[1] / vox - mean_u * mean_u + EPS)
        sc = gam_u.double() * rstd_u
        sh = bet_u.double() - mean_u * sc
        ua = F.leaky_relu(ud * sc[:, None, None, None, :] + sh[:, None, None, None, :], SLOPE)
    else:
        ua = ud
    ref_gu = g_t @ w.double()
    ref_gw = torch.einsum("ndhwc,ndhwk->ck", g_t, ua)
    return dict(gz=gz, t=t, u=u, w=w, st_t=st_t, st_u=st_u, red=red, gam_t=gam_t, bet_t=bet_t, gam_u=gam_u, bet_u=bet_u,
                ref_gu=ref_gu, ref_gw=ref_gw, vox=vox)


def _run(c, N, dims, has_nt, u_norm, accumulate, env):
    from light_unet import _native as nv
    D, H, W = dims
    dev = {k: (v.to(DEV).contiguous() if torch.is_tensor(v) else v) for k, v in c.items()}
    nt = nv.norm(dev["st_t"], dev["gam_t"], dev["bet_t"], None, EPS, 1.0, c["vox"]) if has_nt else nv.norm()
    un = nv.norm(dev["st_u"], dev["gam_u"], dev["bet_u"], None, EPS, SLOPE, c["vox"]) if u_norm else nv.norm()
    torch.manual_seed(3)
    gw0 = torch.randn_like(dev["w"]) * 1e-2
    gu0 = torch.randn(N, D, H, W, c["u"].shape[-1], device=DEV) * 1e-3
    g_w, g_u = gw0.clone(), gu0.clone()
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    nv.refresh_env()
    try:
        nv.call("l3d_pw_bwd", nv.act(dev["gz"]), nv.act(dev["t"]) if has_nt else nv.act(None), nt, nv.ptr(dev["red"]) if has_nt else None,
                nv.act(dev["u"]), un, N, D, H, W, nv.ptr(dev["w"]), nv.ptr(g_w), nv.act(g_u), 1 if accumulate else 0,
                nv.stream_ptr(torch.device(DEV)))
        torch.cuda.synchronize()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        nv.refresh_env()
    return (g_u - (gu0 if accumulate else 0)).double().cpu(), (g_w - gw0).double().cpu()


def _rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-300))


@pytest.mark.parametrize("store", [torch.float16, torch.float32], ids=["f16", "f32"])
@pytest.mark.parametrize("Cg,Cu", [(16, 16), (16, 32), (32, 16), (64, 32), (64, 128), (128, 128), (32, 64)])
@pytest.mark.parametrize("has_nt,u_norm,accumulate", [(True, False, False), (True, True, True), (False, False, False)])
def test_pw_bwd_tensor_core(Cg, Cu, has_nt, u_norm, accumulate, store):
    """store: the storage type of the activations t and u (fp16: split exactly into bf16 hi + lo; fp32: 16 significand
    bits); the gradient gz is fp32 in both modes."""
    N, dims = 2, (10, 9, 13)          # 1170 voxels per sample: ragged last tile
    c = _case(N, dims, Cg, Cu, has_nt, u_norm, Cg * 131 + Cu, store)
    gu_tc, gw_tc = _run(c, N, dims, has_nt, u_norm, accumulate, {})
    gu_cc, gw_cc = _run(c, N, dims, has_nt, u_norm, accumulate, {"L3D_NO_TC_BWD": "1"})
    e = dict(gu_tc=_rel(gu_tc, c["ref_gu"]), gw_tc=_rel(gw_tc, c["ref_gw"]), gu_cc=_rel(gu_cc, c["ref_gu"]), gw_cc=_rel(gw_cc, c["ref_gw"]))
    print(Cg, Cu, has_nt, u_norm, accumulate, {k: f"{v:.2e}" for k, v in e.items()})
    # hi/lo f16 operand pairs: ~2^-16 per product; the fp32 CUDA-core kernel is the yardstick
    assert e["gu_tc"] < 2e-4 and e["gw_tc"] < 2e-4, e
    assert e["gu_cc"] < 2e-4 and e["gw_cc"] < 2e-4, e


@pytest.mark.parametrize("store", [torch.float16, torch.float32], ids=["f16", "f32"])
@pytest.mark.parametrize("Cin,Cout", [(32, 16), (64, 32), (128, 64), (16, 16)])
@pytest.mark.parametrize("lo,out_dims,accumulate", [((5, 6, 7), (10, 12, 14), False), ((3, 4, 5), (7, 9, 10), True)])
def test_convt_bwd_tensor_core(Cin, Cout, lo, out_dims, accumulate, store):
    """ConvTranspose3d(k=2, s=2) backward: input, weight and bias gradients against torch autograd in float64, for the
    tensor-core kernel and the CUDA-core kernel it replaces; g_out is the lower channel half of a concat-buffer gradient
    and the output volume may be larger than 2x the input (centre-pad path, unet3d.py:130-138)."""
    from light_unet import _native as nv
    N = 2
    g = torch.Generator().manual_seed(Cin * 7 + Cout + lo[0])
    x = torch.randn(N, *lo, Cin, generator=g).to(store)
    w = (torch.randn(Cin, Cout, 2, 2, 2, generator=g) / np.sqrt(Cin)).float()
    gcat = (torch.randn(N, *out_dims, 2 * Cout, generator=g) * 1e-3).float()
    off = [(out_dims[k] - 2 * lo[k]) // 2 for k in range(3)]
    # float64 reference through autograd on the un-padded up-sampled region
    xd = x.double().permute(0, 4, 1, 2, 3).requires_grad_(True)
    wd = w.double().requires_grad_(True)
    bd = torch.zeros(Cout, dtype=torch.float64, requires_grad=True)
    y = F.conv_transpose3d(xd, wd, bd, stride=2)
    gsub = gcat[:, off[0]:off[0] + 2 * lo[0], off[1]:off[1] + 2 * lo[1], off[2]:off[2] + 2 * lo[2], :Cout].double().permute(0, 4, 1, 2, 3)
    y.backward(gsub)
    ref_gx = xd.grad.permute(0, 2, 3, 4, 1)
    res = {}
    for name, env in (("tc", {}), ("cc", {"L3D_NO_TC_BWD": "1"})):
        xdv, wdv, gd = x.to(DEV), w.to(DEV), gcat.to(DEV)
        torch.manual_seed(5)
        gw0 = torch.randn_like(wdv) * 1e-2
        gb0 = torch.randn(Cout, device=DEV) * 1e-2
        gx0 = torch.randn(N, *lo, Cin, device=DEV) * 1e-3
        g_w, g_b, g_x = gw0.clone(), gb0.clone(), gx0.clone()
        old = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        nv.refresh_env()
        try:
            nv.call("l3d_convt_bwd", nv.act(gd, 0, Cout), out_dims[0], out_dims[1], out_dims[2], off[0], off[1], off[2], nv.act(xdv), N,
                    lo[0], lo[1], lo[2], nv.ptr(wdv), nv.ptr(g_w), nv.ptr(g_b), nv.act(g_x), 1 if accumulate else 0,
                    nv.stream_ptr(torch.device(DEV)))
            torch.cuda.synchronize()
        finally:
            for k, v in old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
            nv.refresh_env()
        res[name] = (_rel((g_x - (gx0 if accumulate else 0)).double().cpu(), ref_gx), _rel((g_w - gw0).double().cpu(), wd.grad),
                     _rel((g_b - gb0).double().cpu(), bd.grad))
    print(Cin, Cout, lo, out_dims, {k: tuple(f"{e:.1e}" for e in v) for k, v in res.items()})
    for v in res.values():
        assert max(v) < 2e-4, res


@pytest.mark.parametrize("C,dims", [(16, (9, 11, 13)), (32, (8, 16, 8)), (64, (5, 6, 7))])
@pytest.mark.parametrize("has_norm,accumulate", [(True, False), (True, True), (False, False)])
def test_dw_bwd_layer(C, dims, has_norm, accumulate):
    """Depthwise 3x3x3 backward through the C-ABI (input gradient through Dropout3d / LeakyReLU / the producer's
    InstanceNorm prologue, depthwise weight gradient, norm reductions) against a float64 torch restatement."""
    from light_unet import _native as nv
    N = 2
    D, H, W = dims
    vox = D * H * W
    g = torch.Generator().manual_seed(C + D)
    gu = (torch.randn(N, D, H, W, C, generator=g) * 1e-3).float()
    x = torch.randn(N, D, H, W, C, generator=g).to(torch.float16)
    dw = (torch.randn(C, 27, generator=g) / 5).float()
    gam, bet = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.2
    drop = (torch.rand(N, C, generator=g) > 0.2).float() / 0.8
    xd = x.double()
    stats = torch.stack([xd.sum(dim=(1, 2, 3)), (xd * xd).sum(dim=(1, 2, 3))])
    mean = stats[0] / vox
    rstd = 1.0 / torch.sqrt(stats[1] / vox - mean * mean + EPS)
    xhat = (xd - mean[:, None, None, None, :]) * rstd[:, None, None, None, :]
    if has_norm:
        pre = xhat * gam.double() + bet.double()
        a = F.leaky_relu(pre, SLOPE) * drop.double()[:, None, None, None, :]
        dact = torch.where(pre > 0, 1.0, SLOPE) * drop.double()[:, None, None, None, :]
    else:
        a, dact = xd, torch.ones_like(xd)
    a_n = a.permute(0, 4, 1, 2, 3).clone().requires_grad_(True)
    wd = dw.double().view(C, 1, 3, 3, 3).clone().requires_grad_(True)
    u = F.conv3d(a_n, wd, padding=1, groups=C)
    u.backward(gu.double().permute(0, 4, 1, 2, 3))
    ref_gy = a_n.grad.permute(0, 2, 3, 4, 1) * dact
    ref_gdw = wd.grad.view(C, 27)
    ref_red = torch.stack([ref_gy.sum(dim=(1, 2, 3)), (ref_gy * xhat).sum(dim=(1, 2, 3))])
    dev = lambda t: t.to(DEV).contiguous()
    gud, xdv, dwd, st, gd, bd, dd = dev(gu), dev(x), dev(dw), dev(stats), dev(gam), dev(bet), dev(drop)
    xn = nv.norm(st, gd, bd, dd, EPS, SLOPE, vox) if has_norm else nv.norm()
    torch.manual_seed(9)
    gy0 = torch.randn(N, D, H, W, C, device=DEV) * 1e-3
    gdw0 = torch.randn(C, 27, device=DEV) * 1e-2
    gy, gdw = gy0.clone(), gdw0.clone()
    red = torch.zeros(2, N, C, dtype=torch.float64, device=DEV)
    nv.call("l3d_dw_bwd", nv.act(gud), nv.act(xdv), xn, N, D, H, W, nv.ptr(dwd), nv.ptr(gdw), nv.act(gy), 1 if accumulate else 0,
            nv.ptr(red) if has_norm else None, nv.stream_ptr(torch.device(DEV)))
    torch.cuda.synchronize()
    got_gy = (gy - (gy0 if accumulate else 0)).double().cpu()
    assert _rel(got_gy, ref_gy) < 1e-5 and _rel((gdw - gdw0).double().cpu(), ref_gdw) < 1e-5
    if has_norm:
        assert _rel(red.cpu(), ref_red) < 1e-5
