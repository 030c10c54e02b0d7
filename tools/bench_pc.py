"""Micro-benchmark of twowl_pair_conv alone: forward shape (1 source, 1 gather, statistics) and backward shape
(2 sources, 2 gathers) at M rows x C columns.   python tools/bench_pc.py [M] [C] [N_table]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "link-prediction-gnn_b200"))
import torch
from twowl_b200 import ops

M = int(sys.argv[1]) if len(sys.argv) > 1 else 60_000_000
C = int(sys.argv[2]) if len(sys.argv) > 2 else 64
NT = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 20
dev = torch.device("cuda")
torch.manual_seed(0)
A = [torch.randn(M, C, device=dev) for _ in range(2)]
W = [torch.randn(C, C, device=dev) / C ** 0.5 for _ in range(2)]
T = [torch.randn(NT, C, device=dev) for _ in range(2)]
idx = [torch.randint(0, NT, (M,), device=dev, dtype=torch.int32) for _ in range(2)]
coef = [torch.rand(M, device=dev) for _ in range(2)]
rs = [torch.rand(M, device=dev) for _ in range(2)]
bias = torch.randn(C, device=dev)
ms = torch.rand(C, device=dev) + 0.5
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(name, fn, nbytes, reps=5):
    fn()
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    t = sorted(ts)[len(ts) // 2]
    print(f"{name:28s} M={M} C={C}: {t:8.3f} ms  {nbytes / t / 1e6:8.1f} GB/s algorithmic", flush=True)


timeit("fwd (1 src, 1 gather, stats)", lambda: ops.pair_conv([A[0]], [W[0]], [0], row_scale=[rs[0]], gathers=[(T[0], idx[0], coef[0])],
                                                             bias=bias, stats_mean_scale=ms), M * (4 * C + 4 * C + 4 * C + 16))
timeit("bwd (2 src, 2 gathers)", lambda: ops.pair_conv(A, W, [1, 1], row_scale=rs, gathers=[(T[g], idx[g], coef[g]) for g in range(2)]),
       M * (8 * C + 4 * C + 2 * (4 * C + 8) + 8))
timeit("plain linear", lambda: ops.pair_conv([A[0]], [W[0]], [0]), M * 8 * C)
