"""Row-sharded step over NCCL against the single-GPU step on the SAME batch (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/check_rowshard_nccl.py [workload] [hidden]

Every rank evaluates the whole step on its own GPU (the reference: LocalWLNet without a RowShard) and its share of the row-sharded
step (pair rows + node blocks cut over the N ranks, pipelined table all-reduces, parameter gradients summed by
dist.allreduce_grads); rank 0 prints the largest differences of the logits, the loss and every parameter gradient, relative to the
tensor's largest magnitude, and exits non-zero above 2e-5 (the summation orders differ, nothing else).
tests/test_gpu_rowshard.py checks the same against the oracle on one GPU over gloo; this is the NCCL / multi-GPU leg."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "link-prediction-gnn_b200")):
    sys.path.insert(0, p)
import torch
import torch.distributed as dist

import bench
import TwoWL.model.model as model
import TwoWL.utils as U
from twowl_b200 import dist as D
from twowl_b200 import functional as F2
from twowl_b200.rowshard import RowShard, comm_summary

wl = sys.argv[1] if len(sys.argv) > 1 else "collab"
hidden = int(sys.argv[2]) if len(sys.argv) > 2 else 64
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()

g = bench.make_graph(wl, 0, dev)
n, pos, pred, pos1 = g["n"], g["pos"], g["pred"], g["pos1"]
E, P = pos.shape[1], pred.shape[1]
ei2 = U.get_ei2_implicit(n, pos, pred)
nb = max(2, g["und"] // 10)
i1, i2, y = (t.to(dev) for t in bench.draw_batch(g["und"], P // 2, nb, 0, replicate=True))
idx1 = U.double(i1, for_index=True)
idx = torch.cat((idx1, U.double(i2, for_index=True) + E))
ei_new, x_new, ei2_new = U.sample_block(idx1, n, pos, ei2)

res = []
for sharded in (False, True):
    torch.manual_seed(0)
    mod = model.LocalWLNet(int(U.degree(pos, n).max().item()), False, None, hidden, hidden, 2, 1, 0., 0., 0., 0., 0., 0.).to(dev).train()
    with torch.no_grad():
        for p_ in mod.parameters():
            if p_.dim() == 1:
                p_.add_(0.2 * torch.randn_like(p_))
    if sharded:
        mod.row_shard = RowShard()
    out = mod(x_new, ei_new, pos1, idx, ei2_new)
    loss = F2.bce_with_logits(out, y)
    loss.backward()
    if sharded:
        D.allreduce_grads(mod.parameters())
    res.append((out.detach().clone(), loss.detach().clone(), {k: p_.grad.detach().clone() for k, p_ in mod.named_parameters()}))
    del mod, out, loss
    torch.cuda.synchronize()
(o0, l0, g0), (o1, l1, g1) = res
worst = 0.0
lines = []
for name, a, b in [("logits", o0, o1), ("loss", l0.reshape(1), l1.reshape(1))] + [("grad " + k, g0[k], g1[k]) for k in sorted(g0)]:
    scale = float(a.abs().max())
    err = float((a.double() - b.double()).abs().max()) / max(scale, 1e-30)
    worst = max(worst, err if scale > 1e-12 else 0.0)
    lines.append(f"  {name:42s} max|a|={scale:.3e}  max|diff|/max|a|={err:.2e}")
if rank == 0:
    print(f"row-sharded vs single GPU, {wl} hidden {hidden}, world {world} over NCCL: worst relative difference {worst:.2e}")
    print("\n".join(lines))
    print(comm_summary())
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if worst <= 2e-5 else 1)
