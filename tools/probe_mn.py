import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "light-3d-unet-front_b200"))
from light_unet import _native as nv
M, N = 128, 16
VAR = int(sys.argv[1]) if len(sys.argv) > 1 else 0
def run(G, U):
    D = torch.full((128, N), float("nan"), device="cuda")
    nv.call("l3d_tc_selftest_mn16", nv.ptr(G), nv.ptr(U), M, N, VAR, nv.ptr(D), nv.stream_ptr(G.device))
    torch.cuda.synchronize()
    return D
U = torch.ones(128, N, device="cuda")
for (v0, m0) in [(0, 0), (0, 1), (0, 4), (0, 5), (1, 0), (7, 0), (8, 0), (9, 3), (0, 8), (0, 32), (64, 17)]:
    G = torch.zeros(128, M, device="cuda"); G[v0, m0] = 1.0
    D = run(G, U)
    nz = torch.nonzero(D.nan_to_num(99.0) != 0)
    rows = sorted(set(nz[:, 0].tolist()))
    print(f"G one-hot v={v0} m={m0}: nonzero D rows {rows[:12]} vals {D[rows[0]].tolist()[:4] if rows else None}")
# which n lights for U one-hot
G = torch.ones(128, M, device="cuda")
for (v0, n0) in [(0, 0), (0, 1), (0, 4), (1, 0), (8, 0), (9, 5)]:
    U = torch.zeros(128, N, device="cuda"); U[v0, n0] = 1.0
    D = run(G, U)
    nz = torch.nonzero(D.nan_to_num(99.0) != 0)
    cols = sorted(set(nz[:, 1].tolist())); rows = sorted(set(nz[:, 0].tolist()))
    print(f"U one-hot v={v0} n={n0}: nonzero D cols {cols} rows {len(rows)} first {rows[:6]}")

torch.manual_seed(0)
G = torch.randn(128, M, device="cuda"); U = torch.randn(128, N, device="cuda")
D = run(G, U)
ref = G.bfloat16().double().t() @ U.bfloat16().double()
print("random: max err", (D[:M].double() - ref).abs().max().item(), "ref max", ref.abs().max().item())
