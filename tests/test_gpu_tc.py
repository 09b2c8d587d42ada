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
