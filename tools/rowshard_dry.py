"""One rank's COMPUTE of a world-size-W row-sharded step, timed on ONE device with the collectives skipped (RowShard(dry=True)):
where the per-rank time goes when nothing is exchanged - load balance over the ranks and the work that does not shrink with W.
The outputs are partial sums, not results; parity of the sharded path is tests/test_gpu_rowshard.py.
    python tools/rowshard_dry.py [world] [hidden] [ranks, e.g. 0,3,7]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "link-prediction-gnn_b200")):
    sys.path.insert(0, p)
import torch
import bench
from twowl_b200 import ops, graph as G
from twowl_b200.rowshard import RowShard
import TwoWL.model.model as model
import TwoWL.utils as U

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
hidden = int(sys.argv[2]) if len(sys.argv) > 2 else 64
ranks = [int(r) for r in sys.argv[3].split(",")] if len(sys.argv) > 3 else list(range(world))
wl = os.environ.get("WORKLOAD", "rmat")
dev = torch.device("cuda", 0)
g = bench.make_graph(wl, 0, dev)
n, pos, pred, pos1 = g["n"], g["pos"], g["pred"], g["pos1"]
E, P = pos.shape[1], pred.shape[1]
ei2 = U.get_ei2_implicit(n, pos, pred)
nb = max(2, g["und"] // 10)
torch.manual_seed(0)
mod = model.LocalWLNet(int(U.degree(pos, n).max().item()), False, None, hidden, hidden, 1, 1, 0., 0., 0., 0., 0., 0.).to(dev).train()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
batches = [[t.to(dev) for t in bench.draw_batch(g["und"], P // 2, nb, i, replicate=True)] for i in range(6)]


def step(i):
    i1, i2, y = batches[i]
    idx1 = U.double(i1, for_index=True)
    idx = torch.cat((idx1, U.double(i2, for_index=True) + E))
    ei_new, x_new, ei2_new = U.sample_block(idx1, n, pos, ei2)
    for p_ in mod.parameters():
        p_.grad = None
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out = mod(x_new, ei_new, pos1, idx, ei2_new)
    torch.nn.functional.binary_cross_entropy_with_logits(out, y).backward()
    b.record()
    return a, b


if os.environ.get("NCU"):      # under ncu --profile-from-start off: one step of ranks[0] after two warm-up steps
    mod.row_shard = RowShard(rank=ranks[0], world=world, dry=True)
    step(0); step(1)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    step(2)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    sys.exit(0)

if os.environ.get("DRY_KINETO"):      # every kernel of one step (torch glue included) and the idle time between them
    from torch.profiler import profile, ProfilerActivity
    mod.row_shard = RowShard(rank=ranks[0], world=world, dry=True)
    for i in range(3):
        step(i)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        a, b = step(3)
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    per = {}
    for e in evs:
        k = per.setdefault(e.name[:70], [0, 0.0])
        k[0] += 1
        k[1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
    busy = sum(v[1] for v in per.values()) / 1e3
    print(f"rank {ranks[0]}/{world}: step {a.elapsed_time(b):.2f} ms under the profiler, {len(evs)} device activities, {busy:.2f} ms busy")
    for name, (cnt, us) in sorted(per.items(), key=lambda kv: -kv[1][1])[:45]:
        print(f"   {us / 1e3:8.3f} ms {cnt:4d} x  {name}")
    sys.exit(0)

for r in ranks:
    mod.row_shard = RowShard(rank=r, world=world, dry=True)
    for i in range(3):
        step(i)
    torch.cuda.synchronize()
    evs = []
    for i in range(3, 6):
        flush.fill_(i)
        evs.append(step(i))
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in evs)[1]
    flush.fill_(1)
    ops.profile_start()
    step(3)
    per = {}
    order = []
    rec = ops.profile_stop()
    if os.environ.get("DRY_LIST"):
        for name, nbytes, t in rec:
            if t >= 0.03:
                print(f"      {name:28s} {t:7.3f} ms  {nbytes / 1e9:7.2f} GB alg", flush=True)
    for name, nbytes, t in rec:
        if name not in per:
            order.append(name)
        per[name] = per.get(name, 0.0) + t
    tot = sum(per.values())
    print(f"rank {r}/{world}: step {ms:.2f} ms (fwd+bwd, events around the step); sum of op events {tot:.2f} ms", flush=True)
    print("   " + ", ".join(f"{k} {per[k]:.2f}" for k in sorted(per, key=lambda k: -per[k]) if per[k] >= 0.05), flush=True)
    G.clear_cache()
    torch.cuda.empty_cache()
