"""Layer-level GPU parity of the tcgen05 implicit-GEMM 3x3x3 conv (csrc/l3d_conv3_tc.cu) through the C-ABI:
every tile height / accumulator-set / TMA-buffer configuration, ragged volumes, dense, grouped and composed
depthwise-separable (+ shortcut) weights, against torch conv3d in fp32 on the same f16-stored inputs."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
EPS, SLOPE = 1e-5, 0.01


def _rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _inputs(N, dims, Cin, seed):
    g = torch.Generator().manual_seed(seed)
    D, H, W = dims
    x = torch.randn(N, D, H, W, Cin, generator=g).to(torch.float16)
    vox = D * H * W
    xf = x.float()
    # statistics of the stored tensor, as the producer's epilogue would have accumulated them: {sum, sumsq}[N][C]
    stats = torch.stack([xf.sum(dim=(1, 2, 3)), (xf * xf).sum(dim=(1, 2, 3))]).double()
    gamma = torch.rand(Cin, generator=g) + 0.5
    beta = torch.randn(Cin, generator=g) * 0.3
    mean = stats[0] / vox
    var = stats[1] / vox - mean * mean
    rstd = 1.0 / torch.sqrt(var + EPS)
    scale = (gamma.double() * rstd).float()                       # [N][C]
    shift = (beta.double() - mean * gamma.double() * rstd).float()
    a = F.leaky_relu(xf * scale[:, None, None, None, :] + shift[:, None, None, None, :], SLOPE)
    return x, stats, gamma, beta, a, vox


def _expected_launches(case):
    """Dense / grouped layers whose fp16 weight tiles exceed 60 KB run as output-channel slices (l3d_conv3_fwd)."""
    kind, _, _, Cin, Cout = case[:5]
    if kind != "dense" or 27 * Cin * Cout * 2 <= 60 * 1024:
        return 1
    for cs in (32, 16):
        if Cout % cs == 0 and Cout > cs and (cs == 16 or 27 * Cin * cs * 2 <= 60 * 1024):
            return Cout // cs
    return 1


def _run(case, env, launches=None):
    if launches is None:
        launches = _expected_launches(case)
    from light_unet import _native as nv
    kind, N, dims, Cin, Cout, groups, use_norm = case
    D, H, W = dims
    x, stats, gamma, beta, a, vox = _inputs(N, dims, Cin, 7)
    if not use_norm:
        a = x.float()
    g = torch.Generator().manual_seed(11)
    st = nv.stream_ptr(torch.device(DEV))
    xd = x.to(DEV)
    keep = [xd]
    xn = nv.norm()
    if use_norm:
        sd, gd, bd = stats.to(DEV).contiguous(), gamma.to(DEV), beta.to(DEV)
        keep += [sd, gd, bd]
        xn = nv.norm(sd, gd, bd, None, EPS, SLOPE, vox)
    t = torch.zeros(N, D, H, W, Cout, dtype=torch.float16, device=DEV)
    t_stats = torch.zeros(2 * N * Cout, dtype=torch.float64, device=DEV)
    a_ncdhw = a.permute(0, 4, 1, 2, 3).contiguous()
    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    nv.refresh_env()
    try:
        before = nv.launch_count()
        if kind == "dense":
            w = torch.randn(Cout, Cin // groups, 3, 3, 3, generator=g) / np.sqrt(27 * Cin / groups)
            wd = w.to(DEV)
            nv.call("l3d_conv3_fwd", nv.act(xd), xn, N, D, H, W, nv.ptr(wd), groups, nv.act(t), nv.ptr(t_stats), None, nv.act(None), None, st)
            ref_t = F.conv3d(a_ncdhw, w, padding=1, groups=groups)
            ref_r, r = None, None
        else:
            dw = torch.randn(Cin, 1, 3, 3, 3, generator=g) / np.sqrt(27.0)
            pw = torch.randn(Cout, Cin, 1, 1, 1, generator=g) / np.sqrt(Cin)
            sc = torch.randn(Cout, Cin, 1, 1, 1, generator=g) / np.sqrt(Cin)
            dwd, pwd, scd = dw.to(DEV), pw.to(DEV), sc.to(DEV)
            r = torch.zeros(N, D, H, W, Cout, dtype=torch.float16, device=DEV)
            r_stats = torch.zeros(2 * N * Cout, dtype=torch.float64, device=DEV)
            nv.call("l3d_dwpw_fwd", nv.act(xd), xn, N, D, H, W, nv.ptr(dwd), nv.ptr(pwd), nv.ptr(scd), nv.act(t),
                    nv.ptr(t_stats), nv.act(r), nv.ptr(r_stats), nv.act(None), st)
            ref_t = F.conv3d(F.conv3d(a_ncdhw, dw, padding=1, groups=Cin), pw)
            ref_r = F.conv3d(a_ncdhw, sc)
        torch.cuda.synchronize()
        assert nv.launch_count() - before == launches
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        nv.refresh_env()
    outs = [(t, t_stats, ref_t)] + ([(r, r_stats, ref_r)] if r is not None else [])
    for got, gstats, ref in outs:
        got_f = got.float().cpu()
        e = _rel(got_f.permute(0, 4, 1, 2, 3), ref)
        assert e < 4e-3, (case, env, e)
        # the statistics are sums over exactly the stored (f16-rounded) values
        want = torch.stack([got_f.double().sum(dim=(1, 2, 3)), (got_f.double() ** 2).sum(dim=(1, 2, 3))]).reshape(-1)
        gs = gstats.cpu()
        assert float((gs - want).abs().max() / (want.abs().max() + 1e-30)) < 2e-4, (case, env)


CASES = [
    ("dense", 2, (20, 19, 21), 16, 16, 1, True),
    ("dense", 1, (9, 33, 8), 32, 32, 1, True),
    ("dense", 2, (8, 16, 24), 16, 64, 1, False),
    ("dense", 1, (7, 18, 10), 16, 128, 1, True),       # 3*Cout > 256: per-plane (unmerged) MMAs
    ("dense", 1, (12, 16, 16), 32, 16, 8, True),       # grouped weights
    ("dws", 2, (20, 19, 21), 16, 16, 1, True),
    ("dws", 1, (24, 24, 24), 32, 16, 1, False),
    ("dws", 1, (13, 16, 9), 16, 32, 1, True),
    ("dws", 1, (12, 12, 12), 64, 32, 1, True),
]


@pytest.mark.parametrize("tz", [0, 2, 4, 6, 8])
@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c[0]}-{c[3]}to{c[4]}-{'x'.join(map(str, c[2]))}")
def test_conv3_tc_tile_heights(case, tz):
    env = {"L3D_DWS_IGEMM_MAX": 1 << 20}
    if tz:
        nacc = 2 if case[0] == "dws" else 1
        if tz * case[4] * nacc > 512:
            pytest.skip("accumulators of this tile height do not fit TMEM")
        env["L3D_C3_TZ"] = tz
    _run(case, env)


@pytest.mark.parametrize("knobs", [{"L3D_C3_SETS": 1}, {"L3D_C3_NRAW": 1}, {"L3D_C3_NOMERGE": 1},
                                   {"L3D_C3_TZ": 4, "L3D_C3_SETS": 1, "L3D_C3_NRAW": 1},
                                   {"L3D_C3_LOADER": 2, "L3D_C3_WARPS": 12}, {"L3D_C3_LOADER": 2, "L3D_C3_WARPS": 12, "L3D_C3_TZ": 4},
                                   {"L3D_C3_LOADER": 0}, {"L3D_C3_ROT": 1, "L3D_C3_NRAW": 1}, {"L3D_C3_ROT": 1, "L3D_C3_NRAW": 1, "L3D_C3_TZ": 2, "L3D_C3_LOADER": 0}, {"L3D_C3_TMASPLIT": 2}, {"L3D_C3_TMASPLIT": 5, "L3D_C3_TZ": 8},
                                   # requests by the MMA issuer / the producer warp without the half-box split / with it at every tile height / two CTAs per SM
                                   {"L3D_C3_PROD": 0}, {"L3D_C3_PROD": 1, "L3D_C3_SPLIT2": 0}, {"L3D_C3_PROD": 1, "L3D_C3_TZ": 2, "L3D_C3_NRAW": 1}, {"L3D_C3_PROD": 1, "L3D_C3_TZ": 4, "L3D_C3_NRAW": 1},
                                   {"L3D_C3_PROD": 1, "L3D_C3_TZ": 6, "L3D_C3_NRAW": 1}, {"L3D_C3_PROD": 1, "L3D_C3_TZ": 6, "L3D_C3_NRAW": 2}, {"L3D_C3_OCC2": 1}, {"L3D_C3_OCC2": 1, "L3D_C3_TZ": 2}, {"L3D_C3_WARPS": 8}])
def test_conv3_tc_pipeline_variants(knobs):
    for case in (CASES[0], CASES[5], CASES[6]):
        _run(case, dict(knobs, L3D_DWS_IGEMM_MAX=1 << 20))


def test_conv3_tc_output_channel_halves():
    """64 -> 32 (+ shortcut) at >= 16^3: one implicit-GEMM launch with a shorter tile (default dispatch), or -- with the
    weight budget of the narrow layers (L3D_DWS_IGEMM_MAX=1024) -- two launches over output-channel halves, each writing
    its channel slice of t / r and of the statistics rows."""
    _run(("dws", 2, (16, 17, 24), 64, 32, 1, True), {}, launches=1)
    _run(("dws", 1, (24, 24, 24), 64, 32, 1, False), {}, launches=1)
    _run(("dws", 2, (16, 17, 24), 64, 32, 1, True), {"L3D_DWS_IGEMM_MAX": 1024}, launches=2)
    _run(("dws", 1, (24, 24, 24), 64, 32, 1, False), {"L3D_DWS_IGEMM_MAX": 1024}, launches=2)


@pytest.mark.parametrize("case,launches", [(("dense", 1, (8, 16, 16), 64, 64, 1, True), 4), (("dense", 1, (6, 16, 8), 32, 64, 1, False), 2),
                                           (("dense", 1, (5, 12, 9), 128, 128, 1, True), 8), (("dense", 1, (8, 16, 8), 64, 64, 8, True), 4)])
def test_conv3_tc_wide_layers_in_output_slices(case, launches):
    """Dense / grouped layers whose 27 weight tiles do not fit shared memory run as several implicit-GEMM launches over
    output-channel slices (grouped: the slice keeps its absolute channel index for the group lookup)."""
    _run(case, {}, launches=launches)


@pytest.mark.parametrize("dims", [(16, 16, 16), (20, 19, 24), (9, 33, 8), (12, 10, 20), (48, 48, 48)])
@pytest.mark.parametrize("tz", [0, 2, 4, 6, 8])
def test_rank1_first_block_layers(dims, tz):
    """l3d_dw_c1_fwd (u = dw * x, analytic statistics of pw (x) u and sc (x) x) and l3d_dwpw_fwd_rank1 (the second
    depthwise-separable conv reading lrelu(IN(pw (x) u)) evaluated on the fly) against torch in fp32."""
    from light_unet import _native as nv
    N, C = 2, 16
    D, H, W = dims
    vox = D * H * W
    g = torch.Generator().manual_seed(3)
    x = torch.rand(N, D, H, W, 1, generator=g).to(torch.float16)
    dw1 = torch.randn(1, 1, 3, 3, 3, generator=g) / np.sqrt(27.0)
    pw1 = torch.randn(C, 1, 1, 1, 1, generator=g)
    sc = torch.randn(C, 1, 1, 1, 1, generator=g)
    gamma, beta = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.3
    dw2 = torch.randn(C, 1, 3, 3, 3, generator=g) / np.sqrt(27.0)
    pw2 = torch.randn(C, C, 1, 1, 1, generator=g) / np.sqrt(C)
    st = nv.stream_ptr(torch.device(DEV))
    xd, dw1d, pw1d, scd, gd, bd, dw2d, pw2d = (v.to(DEV) for v in (x, dw1, pw1, sc, gamma, beta, dw2, pw2))
    u = torch.zeros(N, D, H, W, dtype=torch.float32, device=DEV)
    s1 = torch.zeros(2 * N * C, dtype=torch.float64, device=DEV)
    sr = torch.zeros(2 * N * C, dtype=torch.float64, device=DEV)
    nv.call("l3d_dw_c1_fwd", nv.act(xd), nv.norm(), N, D, H, W, nv.ptr(dw1d), nv.ptr(pw1d), nv.ptr(scd), C, nv.ptr(u), nv.ptr(s1), nv.ptr(sr), st)
    torch.cuda.synchronize()
    x_ncdhw = x.float().permute(0, 4, 1, 2, 3).contiguous()
    u_ref = F.conv3d(x_ncdhw, dw1, padding=1)
    assert _rel(u.cpu()[:, None], u_ref) < 1e-6
    t1_ref = F.conv3d(u_ref, pw1).double()
    r_ref = F.conv3d(x_ncdhw, sc).double()
    for got, ref in ((s1, t1_ref), (sr, r_ref)):
        want = torch.stack([ref.sum(dim=(2, 3, 4)), (ref * ref).sum(dim=(2, 3, 4))]).reshape(-1)
        assert float((got.cpu() - want).abs().max() / want.abs().max()) < 1e-5
    a1 = F.leaky_relu(F.instance_norm(t1_ref.float(), weight=gamma, bias=beta, eps=EPS), SLOPE)
    t2_ref = F.conv3d(F.conv3d(a1, dw2, padding=1, groups=C), pw2)
    t2 = torch.zeros(N, D, H, W, C, dtype=torch.float16, device=DEV)
    s2 = torch.zeros(2 * N * C, dtype=torch.float64, device=DEV)
    n1 = nv.norm(s1, gd, bd, None, EPS, SLOPE, vox)
    old = os.environ.get("L3D_C3_TZ")
    if tz:
        os.environ["L3D_C3_TZ"] = str(tz)
    nv.refresh_env()
    try:
        nv.call("l3d_dwpw_fwd_rank1", nv.ptr(u), nv.ptr(pw1d), C, n1, N, D, H, W, nv.ptr(dw2d), nv.ptr(pw2d), nv.act(t2), nv.ptr(s2), st)
        torch.cuda.synchronize()
    finally:
        if old is None:
            os.environ.pop("L3D_C3_TZ", None)
        else:
            os.environ["L3D_C3_TZ"] = old
        nv.refresh_env()
    got_f = t2.float().cpu()
    e = _rel(got_f.permute(0, 4, 1, 2, 3), t2_ref)
    assert e < 4e-3, (dims, tz, e)
    want = torch.stack([got_f.double().sum(dim=(1, 2, 3)), (got_f.double() ** 2).sum(dim=(1, 2, 3))]).reshape(-1)
    assert float((s2.cpu() - want).abs().max() / (want.abs().max() + 1e-30)) < 2e-4


SLAB_CASES = [  # N, dims, Cin, Cout, shortcut, normed input
    (3, (12, 12, 12), 32, 64, True, False), (2, (12, 12, 12), 64, 64, False, True), (2, (12, 12, 12), 128, 64, True, False),
    (3, (6, 6, 6), 64, 128, True, False), (5, (6, 6, 6), 128, 128, False, True),
    (2, (10, 12, 12), 64, 64, False, True), (1, (7, 8, 4), 64, 64, True, True), (2, (5, 6, 12), 64, 128, False, True),
    (2, (8, 8, 8), 64, 128, True, False), (1, (3, 16, 16), 48, 80, True, True),
]


@pytest.mark.parametrize("case", SLAB_CASES, ids=lambda c: f"{c[2]}to{c[3]}{'+sc' if c[4] else ''}-{'x'.join(map(str, c[1]))}")
def test_dwpw_slab_kernel(case):
    """Small-volume depthwise-separable conv (csrc/l3d_fwd_slab.cu: whole-plane slabs, the 12^3 / 6^3 levels of a 48^3
    window) through l3d_dwpw_fwd against torch in fp32 on the same f16-stored inputs."""
    from light_unet import _native as nv
    N, dims, Cin, Cout, has_sc, use_norm = case
    D, H, W = dims
    x, stats, gamma, beta, a, vox = _inputs(N, dims, Cin, 13)
    if not use_norm:
        a = x.float()
    g = torch.Generator().manual_seed(17)
    st = nv.stream_ptr(torch.device(DEV))
    xd = x.to(DEV)
    xn = nv.norm()
    if use_norm:
        sd, gd, bd = stats.to(DEV).contiguous(), gamma.to(DEV), beta.to(DEV)
        xn = nv.norm(sd, gd, bd, None, EPS, SLOPE, vox)
    dw = torch.randn(Cin, 1, 3, 3, 3, generator=g) / np.sqrt(27.0)
    pw = torch.randn(Cout, Cin, 1, 1, 1, generator=g) / np.sqrt(Cin)
    sc = torch.randn(Cout, Cin, 1, 1, 1, generator=g) / np.sqrt(Cin)
    dwd, pwd, scd = dw.to(DEV), pw.to(DEV), sc.to(DEV)
    # outputs are views into wider buffers (ldc > C), as the engine's concat buffers are
    tb = torch.zeros(N, D, H, W, 2 * Cout, dtype=torch.float16, device=DEV)
    t = tb[..., Cout:]
    r = torch.zeros(N, D, H, W, Cout, dtype=torch.float16, device=DEV)
    t_stats = torch.zeros(2 * N * Cout, dtype=torch.float64, device=DEV)
    r_stats = torch.zeros(2 * N * Cout, dtype=torch.float64, device=DEV)
    nv.call("l3d_dwpw_fwd", nv.act(xd), xn, N, D, H, W, nv.ptr(dwd), nv.ptr(pwd), nv.ptr(scd) if has_sc else None, nv.act(tb, Cout, Cout),
            nv.ptr(t_stats), nv.act(r) if has_sc else nv.act(None), nv.ptr(r_stats) if has_sc else None, nv.act(None), st)
    torch.cuda.synchronize()
    assert nv.lib().l3d_last_kernel() == b"dwpw_slab_kernel"
    assert float(tb[..., :Cout].float().abs().max()) == 0.0          # the other half of the buffer is untouched
    a_ncdhw = a.permute(0, 4, 1, 2, 3).contiguous()
    outs = [(t, t_stats, F.conv3d(F.conv3d(a_ncdhw, dw, padding=1, groups=Cin), pw))]
    if has_sc:
        outs.append((r, r_stats, F.conv3d(a_ncdhw, sc)))
    for got, gstats, ref in outs:
        got_f = got.float().cpu()
        e = _rel(got_f.permute(0, 4, 1, 2, 3), ref)
        assert e < 4e-3, (case, e)
        want = torch.stack([got_f.double().sum(dim=(1, 2, 3)), (got_f.double() ** 2).sum(dim=(1, 2, 3))]).reshape(-1)
        assert float((gstats.cpu() - want).abs().max() / (want.abs().max() + 1e-30)) < 2e-4, case


TC_CASES = [(2, (9, 17, 13), 16, 16, True, True), (1, (8, 8, 16), 32, 64, True, True), (2, (6, 6, 6), 128, 128, False, True),
            (1, (12, 12, 12), 128, 64, True, False), (1, (13, 10, 9), 64, 32, True, True), (1, (24, 24, 24), 16, 32, True, True)]


@pytest.mark.parametrize("store,tol", [(torch.float16, 2e-3), (torch.float32, 2e-6)], ids=["f16", "f32"])
@pytest.mark.parametrize("case", TC_CASES, ids=lambda c: f"{c[2]}to{c[3]}{'+sc' if c[4] else ''}-{'x'.join(map(str, c[1]))}")
def test_dwpw_tc_training_forward(case, store, tol):
    """The training forward (csrc/l3d_fwd_tc.cu: depthwise stencil on the CUDA cores, pointwise (+ shortcut) GEMM on
    tcgen05, the depthwise output u saved for the weight gradient) in both storage modes.  fp32 storage carries every MMA
    operand as an fp16 hi + lo pair: outputs within 2e-6 relative L2 of torch fp32, i.e. tensor cores at fp32-class accuracy."""
    from light_unet import _native as nv
    N, dims, Cin, Cout, has_sc, use_norm = case
    D, H, W = dims
    x, stats, gamma, beta, a, vox = _inputs(N, dims, Cin, 21)
    x = x.float().to(store)
    if store == torch.float32:          # statistics / activated reference of the fp32-stored tensor
        xf = x.double()
        stats = torch.stack([xf.sum(dim=(1, 2, 3)), (xf * xf).sum(dim=(1, 2, 3))])
        a = F.leaky_relu(F.instance_norm(x.permute(0, 4, 1, 2, 3), weight=gamma, bias=beta, eps=EPS), SLOPE).permute(0, 2, 3, 4, 1)
    if not use_norm:
        a = x.float()
    g = torch.Generator().manual_seed(29)
    st = nv.stream_ptr(torch.device(DEV))
    xd = x.to(DEV)
    xn = nv.norm()
    if use_norm:
        sd, gd, bd = stats.to(DEV).contiguous(), gamma.to(DEV), beta.to(DEV)
        xn = nv.norm(sd, gd, bd, None, EPS, SLOPE, vox)
    dw = torch.randn(Cin, 1, 3, 3, 3, generator=g) / np.sqrt(27.0)
    pw = torch.randn(Cout, Cin, 1, 1, 1, generator=g) / np.sqrt(Cin)
    sc = torch.randn(Cout, Cin, 1, 1, 1, generator=g) / np.sqrt(Cin)
    dwd, pwd, scd = dw.to(DEV), pw.to(DEV), sc.to(DEV)
    t = torch.zeros(N, D, H, W, Cout, dtype=store, device=DEV)
    r = torch.zeros(N, D, H, W, Cout, dtype=store, device=DEV)
    u = torch.zeros(N, D, H, W, Cin, dtype=store, device=DEV)
    t_stats = torch.zeros(2 * N * Cout, dtype=torch.float64, device=DEV)
    r_stats = torch.zeros(2 * N * Cout, dtype=torch.float64, device=DEV)
    nv.call("l3d_dwpw_fwd", nv.act(xd), xn, N, D, H, W, nv.ptr(dwd), nv.ptr(pwd), nv.ptr(scd) if has_sc else None, nv.act(t),
            nv.ptr(t_stats), nv.act(r) if has_sc else nv.act(None), nv.ptr(r_stats) if has_sc else None, nv.act(u), st)
    torch.cuda.synchronize()
    assert nv.lib().l3d_last_kernel() == b"dwpw_tc_kernel"
    a_ncdhw = a.double().permute(0, 4, 1, 2, 3).contiguous()
    u_ref = F.conv3d(a_ncdhw, dw.double(), padding=1, groups=Cin)
    outs = [("u", u, None, u_ref), ("t", t, t_stats, F.conv3d(u_ref, pw.double()))]
    if has_sc:
        outs.append(("r", r, r_stats, F.conv3d(a_ncdhw, sc.double())))
    for name, got, gstats, ref in outs:
        got_f = got.double().cpu()
        e = _rel(got_f.permute(0, 4, 1, 2, 3), ref)
        assert e < tol, (case, name, e)
        if gstats is not None:
            want = torch.stack([got_f.sum(dim=(1, 2, 3)), (got_f ** 2).sum(dim=(1, 2, 3))]).reshape(-1)
            assert float((gstats.cpu() - want).abs().max() / (want.abs().max() + 1e-30)) < 2e-4, (case, name)


@pytest.mark.parametrize("store,tol", [(torch.float16, 2e-3), (torch.float32, 2e-6)], ids=["f16", "f32"])
@pytest.mark.parametrize("Cin,Cout,lo,out_dims", [(32, 16, (5, 6, 7), (10, 12, 14)), (64, 32, (3, 4, 5), (7, 9, 10)), (128, 64, (3, 3, 3), (6, 6, 6))])
def test_convt_tc_forward(Cin, Cout, lo, out_dims, store, tol):
    """ConvTranspose3d(k=2, s=2) + bias into the lower half of a concat buffer at the centre-pad offset, both storage
    modes (fp32 storage, 128 -> 64: the operand tiles do not fit and the generic kernel runs -- same result)."""
    from light_unet import _native as nv
    N = 2
    g = torch.Generator().manual_seed(Cin + Cout)
    x = torch.randn(N, *lo, Cin, generator=g).to(store)
    w = (torch.randn(Cin, Cout, 2, 2, 2, generator=g) / np.sqrt(Cin)).float()
    b = torch.randn(Cout, generator=g).float()
    off = [(out_dims[k] - 2 * lo[k]) // 2 for k in range(3)]
    cat = torch.zeros(N, *out_dims, 2 * Cout, dtype=store, device=DEV)
    xd, wd, bd = x.to(DEV), w.to(DEV), b.to(DEV)
    nv.call("l3d_convt_fwd", nv.act(xd), N, lo[0], lo[1], lo[2], nv.ptr(wd), nv.ptr(bd), nv.act(cat, 0, Cout), out_dims[0], out_dims[1], out_dims[2],
            off[0], off[1], off[2], nv.stream_ptr(torch.device(DEV)))
    torch.cuda.synchronize()
    ref = F.conv_transpose3d(x.double().permute(0, 4, 1, 2, 3), w.double(), b.double(), stride=2).permute(0, 2, 3, 4, 1)
    got = cat[:, off[0]:off[0] + 2 * lo[0], off[1]:off[1] + 2 * lo[1], off[2]:off[2] + 2 * lo[2], :Cout].double().cpu()
    assert _rel(got, ref) < tol
    assert float(cat[..., Cout:].float().abs().max()) == 0.0


@pytest.mark.parametrize("dims", [(16, 16, 16), (9, 17, 13), (48, 48, 48)])
def test_split_concat_input_matches_interleaved(dims):
    """l3d_dwpw_fwd2: the decoder block of the top level reads [ConvTranspose output | skip] as two dense 16-channel tensors.
    Same MMAs in the same order as l3d_dwpw_fwd on the interleaved 32-channel buffer: raw outputs bit-identical,
    statistics equal up to the order of the atomics."""
    from light_unet import _native as nv
    N, (D, H, W) = 3, dims
    g = torch.Generator().manual_seed(3)
    lo = torch.randn(N, D, H, W, 16, generator=g).to(torch.float16).to(DEV)
    hi = torch.randn(N, D, H, W, 16, generator=g).to(torch.float16).to(DEV)
    cat = torch.cat([lo, hi], dim=-1).contiguous()
    dw = (torch.randn(32, 1, 3, 3, 3, generator=g) / 5).to(DEV)
    pw = (torch.randn(16, 32, generator=g) / 32 ** 0.5).to(DEV)
    sc = (torch.randn(16, 32, generator=g) / 32 ** 0.5).to(DEV)
    st = nv.stream_ptr(torch.device(DEV))
    outs = []
    for split in (False, True):
        t = torch.zeros(N, D, H, W, 16, dtype=torch.float16, device=DEV)
        r = torch.zeros_like(t)
        ts = torch.zeros(2 * N * 16, dtype=torch.float64, device=DEV)
        rs = torch.zeros_like(ts)
        if split:
            nv.call("l3d_dwpw_fwd2", nv.act(lo), nv.act(hi), nv.norm(), N, D, H, W, nv.ptr(dw), nv.ptr(pw), nv.ptr(sc), nv.act(t), nv.ptr(ts),
                    nv.act(r), nv.ptr(rs), st)
        else:
            nv.call("l3d_dwpw_fwd", nv.act(cat), nv.norm(), N, D, H, W, nv.ptr(dw), nv.ptr(pw), nv.ptr(sc), nv.act(t), nv.ptr(ts),
                    nv.act(r), nv.ptr(rs), nv.act(None), st)
        torch.cuda.synchronize()
        outs.append((t, r, ts, rs))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    for k in (2, 3):
        assert torch.allclose(outs[0][k], outs[1][k], rtol=1e-5, atol=1e-4)      # per-CTA partial sums are fp32
    # against torch: depthwise 3x3x3 then pointwise, and the 1x1x1 shortcut
    x = cat.float().permute(0, 4, 1, 2, 3)
    want_t = F.conv3d(F.conv3d(x, dw, padding=1, groups=32), pw.view(16, 32, 1, 1, 1))
    want_r = F.conv3d(x, sc.view(16, 32, 1, 1, 1))
    assert _rel(outs[1][0].float().permute(0, 4, 1, 2, 3), want_t) < 2e-3
    assert _rel(outs[1][1].float().permute(0, 4, 1, 2, 3), want_r) < 2e-3
    with pytest.raises(nv.NativeError):            # interleaved halves are not dense tensors
        nv.call("l3d_dwpw_fwd2", nv.act(cat, 0, 16), nv.act(cat, 16, 16), nv.norm(), N, D, H, W, nv.ptr(dw), nv.ptr(pw), nv.ptr(sc),
                nv.act(outs[0][0]), nv.ptr(outs[0][2]), nv.act(outs[0][1]), nv.ptr(outs[0][3]), st)
