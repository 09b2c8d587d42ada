"""tcgen05 building blocks (descriptor layout, TMEM accumulate / load) against a torch fp32 matmul."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("MT,K,N", [(1, 16, 16), (1, 32, 16), (2, 16, 32), (2, 64, 64), (2, 128, 128), (1, 128, 64), (4, 32, 128)])
def test_tcgen05_gemm_selftest(MT, K, N):
    from light_unet import _native as nv
    torch.manual_seed(MT * 1000 + K + N)
    A = torch.randn(MT * 128, K, device="cuda")
    W = torch.randn(N, K, device="cuda")
    D = torch.full((MT * 128, N), float("nan"), device="cuda")
    nv.call("l3d_tc_selftest", nv.ptr(A), nv.ptr(W), MT, K, N, nv.ptr(D), nv.stream_ptr(A.device))
    torch.cuda.synchronize()
    ref = A.half().float() @ W.half().float().t()          # fp16 operands, fp32 accumulation
    err = (D - ref).abs().max().item()
    assert err < 1e-3 * max(1.0, ref.abs().max().item()), (MT, K, N, err)


def _tf32(x):
    return (x.view(torch.int32) & ~0x1fff).view(torch.float32)     # truncate to 10 mantissa bits


@pytest.mark.parametrize("K,N", [(8, 16), (16, 16), (32, 64), (128, 128), (64, 256)])
def test_tf32_kmajor_gemm(K, N):
    from light_unet import _native as nv
    torch.manual_seed(K + N)
    G = torch.randn(128, K, device="cuda")
    W = torch.randn(N, K, device="cuda")
    D = torch.full((128, N), float("nan"), device="cuda")
    nv.call("l3d_tc_selftest_tf32", nv.ptr(G), nv.ptr(W), 0, 128, K, N, nv.ptr(D), nv.stream_ptr(G.device))
    torch.cuda.synchronize()
    ref = G.double() @ W.double().t()
    err = (D.double() - ref).abs().max().item()
    assert err < 4e-3 * max(1.0, ref.abs().max().item()), (K, N, err)


@pytest.mark.parametrize("M,N", [(16, 16), (32, 16), (16, 32), (64, 64), (128, 128), (128, 256), (8, 16)])
def test_bf16_mnmajor_voxel_reduction(M, N):
    """The weight-gradient form of the tensor-core backward kernels: both operands MN-major over voxel-planar bf16 tiles
    [channel/8][128 voxels][8], reduction over the 128 voxels (kind::tf32 rejects MN-major operands on this part -- it
    returns zeros -- which is why those kernels use bf16 hi/lo pairs; mixing a bf16 A with an fp16 B in one kind::f16 MMA
    is an illegal instruction, which is why the stored fp16 activations are split into bf16 hi + lo as well)."""
    from light_unet import _native as nv
    torch.manual_seed(M * 7 + N)
    G = torch.randn(128, M, device="cuda")
    U = torch.randn(128, N, device="cuda")
    D = torch.full((128, N), float("nan"), device="cuda")
    nv.call("l3d_tc_selftest_mn16", nv.ptr(G), nv.ptr(U), M, N, 0, nv.ptr(D), nv.stream_ptr(G.device))
    torch.cuda.synchronize()
    ref = G.bfloat16().double().t() @ U.bfloat16().double()
    err = (D[:M].double() - ref).abs().max().item()
    assert err < 1e-4 * max(1.0, ref.abs().max().item()), (M, N, err)
