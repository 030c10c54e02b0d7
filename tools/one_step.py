"""ONE fwd+bwd step of a bench workload between cudaProfilerStart / Stop (for `ncu --profile-from-start off`), after 3 warm-up steps.
    python tools/one_step.py [workload] [hidden]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "link-prediction-gnn_b200")):
    sys.path.insert(0, p)
import torch
import bench
from twowl_b200 import functional as F2
import TwoWL.model.model as model
import TwoWL.utils as U

wl = sys.argv[1] if len(sys.argv) > 1 else "rmat"
hidden = int(sys.argv[2]) if len(sys.argv) > 2 else bench.WORKLOADS[wl][3]
dev = torch.device("cuda", 0)
g = bench.make_graph(wl, 0, dev)
n, pos, pred, pos1 = g["n"], g["pos"], g["pred"], g["pos1"]
E, P = pos.shape[1], pred.shape[1]
ei2 = U.get_ei2_implicit(n, pos, pred)
nb = max(2, g["und"] // 10)
torch.manual_seed(0)
mod = model.LocalWLNet(int(U.degree(pos, n).max().item()), False, None, hidden, hidden, 1, 1, 0., 0., 0., 0., 0., 0.).to(dev).train()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def prepare(i):
    i1, i2, y = (t.to(dev) for t in bench.draw_batch(g["und"], P // 2, nb, i))
    idx1 = U.double(i1, for_index=True)
    idx = torch.cat((idx1, U.double(i2, for_index=True) + E))
    return U.sample_block(idx1, n, pos, ei2) + (idx, y)


def step(inp):
    ei_new, x_new, ei2_new, idx, y = inp
    for p_ in mod.parameters():
        p_.grad = None
    out = mod(x_new, ei_new, pos1, idx, ei2_new)
    F2.bce_with_logits(out, y).backward()


for i in range(3):
    step(prepare(i))
inp = prepare(3)
flush.fill_(1)
torch.cuda.synchronize()
torch.cuda.profiler.start()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
step(inp)
b.record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(f"one step: {a.elapsed_time(b):.2f} ms")
