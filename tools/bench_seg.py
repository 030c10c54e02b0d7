"""Micro-benchmark of the segmented gather-reduce flavours the TwoWL step issues (same graph as bench.py)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "link-prediction-gnn_b200"))
import torch
import bench
from twowl_b200 import ops, graph as G
import TwoWL.utils as U

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 20
C = int(sys.argv[2]) if len(sys.argv) > 2 else 64
dev = torch.device("cuda")
g = bench.make_graph("rmat", 0, dev, scale, (16_000_000 >> (20 - scale)) if scale != 20 else None)
n, pos, pred, pos1 = g["n"], g["pos"], g["pred"], g["pos1"]
E, P = pos.shape[1], pred.shape[1]
R = E + P
ws = G.build_wedge_struct(n, pos, pred)
nb = g["und"] // 10
idx1 = U.double(torch.randperm(g["und"], device=dev)[:nb], for_index=True)
blocked = torch.zeros(E, dtype=torch.uint8, device=dev)
blocked[idx1] = 1
ws = ws.with_blocked(blocked)
_, centre, dinv, selfw, bnode = ws.prepared()
pt = G.pair_table(pos1, n)
ng = G.node_graph(pos, n)
H = torch.randn(R, C, device=dev)
X = torch.randn(n, C, device=dev)
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
    ms = sorted(ts)[len(ts) // 2]
    print(f"{name:34s} {ms:8.3f} ms  {nbytes / ms / 1e6:8.1f} GB/s (algorithmic {nbytes / 1e9:.2f} GB)")


row = 4 * C
timeit("SH: in-list gather of H (E rows)", lambda: ops.seg_reduce(ws.in_ptr, ws.in_ids, n, H, plan=ws.in_plan, flip=1, src_scale=dinv[0],
                                                                  skip_mask=blocked), E * (row + 9) + n * (row + 8))
timeit("dS: out-list gather of dO (R rows)", lambda: ops.seg_reduce(ws.out_ptr, ws.out_ids, n, H, plan=ws.out_plan, flip=0,
                                                                    src_scale=dinv[0]), R * (row + 8) + n * (row + 8))
timeit("pair_init bwd (R rows * X[dst])", lambda: ops.seg_reduce(pt.ptr_s, pt.ids_s, n, H, plan=pt.plan_s, X2=X, mul_idx=pt.dst),
       R * (2 * row + 8) + n * (row + 8))
timeit("pair_init bwd mated (one pass)", lambda: ops.seg_reduce(pt.ptr_s, pt.ids_s, n, H, plan=pt.plan_s, X2=X, mul_idx=pt.dst, pair_sum=True),
       R * (3 * row + 8) + n * (row + 8))
timeit("node GCN fwd (E rows of X)", lambda: ops.seg_reduce(ng.ptr, ng.col, n, X, plan=ng.plan, src_scale=ng.dinv, dst_scale=ng.dinv,
                                                            skip_self=True, self_mode=1), E * (row + 8) + n * (2 * row + 8))
timeit("dense copy H -> H' (roofline ref)", lambda: H.clone(), 2 * R * row)
