"""Throughput of the integer graph operators of TwoWL/utils.py on a power-law graph whose wedge index fits in HBM:
get_ei2 (wedge join), blockei2 (order-preserving compaction), reverse, sample_block, double, degree, and the CSR build the
explicit pair path makes over the wedges.   python tools/bench_index.py [scale] [edge_samples]
GB/s = algorithmic bytes (int64 at the API) / CUDA-event time, median of 3 after one warm-up."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "link-prediction-gnn_b200")):
    sys.path.insert(0, p)
import torch
import bench
import TwoWL.utils as U
from twowl_b200 import ops

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 18
samples = int(sys.argv[2]) if len(sys.argv) > 2 else 1_300_000
dev = torch.device("cuda", 0)
g = bench.make_graph("collab", 0, dev, scale, samples)
n, pos, pred = g["n"], g["pos"], g["pred"]
E, P = pos.shape[1], pred.shape[1]
print(f"# R-MAT scale {scale}, {samples} edge samples: n = {n}, E = {E} observed edge rows, P = {P} prediction rows, one B200", flush=True)


def timeit(name, fn, nbytes, reps=3):
    out = fn()
    ts = []
    for _ in range(reps):
        del out
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    t = sorted(ts)[len(ts) // 2]
    print(f"{name:46s} {t:9.3f} ms   {nbytes / 1e9:8.2f} GB   {nbytes / t / 1e6:8.1f} GB/s", flush=True)
    return out


ei2 = timeit("get_ei2 (count + scan + fill)", lambda: U.get_ei2(n, pos, pred), 0)
T = ei2.shape[1]
print(f"# T = {T} wedges ({16 * T / 1e9:.1f} GB as int64 [2,T])")
timeit("get_ei2 again, bytes = 16 T written + lists", lambda: U.get_ei2(n, pos, pred), 16 * T + 8 * (2 * E + P))
blk = U.double(torch.randperm(E // 2, device=dev)[: E // 20], for_index=True)       # 10 % of the undirected edges, as train.py
out = timeit("blockei2 (mask + count + order-kept fill)", lambda: U.blockei2(ei2, blk), 0)
Tb = out.shape[1]
timeit("blockei2, bytes = 16 (T + T')", lambda: U.blockei2(ei2, blk), 16 * (T + Tb))
timeit("sample_block (ei filter + degree + blockei2)", lambda: U.sample_block(blk, n, pos, ei2), 16 * (T + Tb) + 16 * (2 * E))
wi = U.get_ei2_implicit(n, pos, pred)
timeit("sample_block on the factored index (WedgeIndex)", lambda: U.sample_block(blk, n, pos, wi), 16 * (2 * E) + E)
del out
r = timeit("reverse (two [2,T'] outputs)", lambda: U.reverse(ei2), 16 * T * 3)
del r
timeit("double (edges) + degree", lambda: (U.double(pos[:, ::2].contiguous()), U.degree(pos, n)), 16 * E * 1.5 + 8 * E + 8 * n)
buf = torch.empty(2 * T, dtype=torch.int64, device=dev)
timeit("reference: torch fill of 16 T bytes (write only)", lambda: buf.fill_(7), 16 * T)
buf2 = torch.empty(2 * T, dtype=torch.int64, device=dev)
timeit("reference: torch copy of 16 T bytes (read + write)", lambda: buf2.copy_(buf), 32 * T)
del buf, buf2
keys = ei2[1]
timeit("csr_build over the wedges by target pair", lambda: ops.csr_build(keys, E + P), 8 * T + 4 * T + 8 * (E + P))
