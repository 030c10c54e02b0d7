"""Debug probe for twowl_pair_dw (MN-major tcgen05 operands): one-hot inputs -> where does the 1 land?"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "link-prediction-gnn_b200"))
import torch
from twowl_b200 import ops

C = int(sys.argv[1]) if len(sys.argv) > 1 else 64
M = 64
ones = torch.ones(M, device="cuda")
for (m0, co0, ci0) in [(0, 0, 0), (0, 1, 0), (0, 0, 1), (1, 0, 0), (0, 5, 9), (9, 37 % C, 21 % C), (63, C - 1, C - 1), (8, 4, 4)]:
    dOf = torch.zeros(M, C, device="cuda"); dOr = torch.zeros(M, C, device="cuda"); H = torch.zeros(M, C, device="cuda")
    dOf[m0, co0] = 1.0
    dOr[m0, co0] = 2.0
    H[m0, ci0] = 1.0
    gf, gr = ops.pair_dw(dOf, dOr, ones, ones, H)
    nzf = torch.nonzero(gf).tolist()[:6]
    nzr = torch.nonzero(gr).tolist()[:6]
    print(f"probe m={m0} co={co0} ci={ci0}: dWf nonzero {nzf} vals {[round(float(gf[i, j]), 3) for i, j in nzf]} | dWr {nzr} "
          f"vals {[round(float(gr[i, j]), 3) for i, j in nzr]}  sums {float(gf.sum()):.3f} {float(gr.sum()):.3f}")
torch.manual_seed(0)
dOf, dOr, H = (torch.randn(M, C, device="cuda") for _ in range(3))
gf, gr = ops.pair_dw(dOf, dOr, ones, ones, H)
ref = dOf.double().t() @ H.double()
print("random: max|got|", float(gf.abs().max()), "max|ref|", float(ref.abs().max()), "max err", float((gf.double() - ref).abs().max()))
print("got[0,:4]", gf[0, :4].tolist(), "ref[0,:4]", ref[0, :4].tolist())
print("got^T err", float((gf.double().t() - ref).abs().max()))
