"""Per-parameter deviation table: row-sharded (world W) vs single GPU vs fp64 oracle on the fb-pages-food train step."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "link-prediction-gnn_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from test_gpu_rowshard import _run_world
from rowshard_worker import CFG, build_step
from oracle import twowl_oracle as O
world, c2, seed = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
depth2 = int(sys.argv[4]) if len(sys.argv) > 4 else 1
got = _run_world(world, c2, seed, "/tmp/rs_out.npz", depth2)
mod, args, y = build_step(c2, seed, depth2)
out = mod(*args)
loss = torch.nn.functional.binary_cross_entropy_with_logits(out, y); loss.backward()
sd = {k: v.detach().cpu() for k, v in mod.state_dict().items()}
x, e1, pos, idx, ei2 = args
ei2_np = ei2.materialize().cpu().numpy() if hasattr(ei2, "materialize") else ei2.cpu().numpy()
r64 = O.fwd_bwd({k: v.double() for k, v in sd.items()}, x.cpu(), e1.cpu(), pos.cpu(), idx.cpu(), ei2_np, y.cpu().double(), CFG["act0"], CFG["act1"])
print("logits: shard-1gpu %.3e  shard-f64 %.3e  1gpu-f64 %.3e" % (np.abs(got["logits"] - out.detach().cpu().numpy()).max(),
      np.abs(got["logits"] - r64[0].numpy()).max(), np.abs(out.detach().cpu().numpy() - r64[0].numpy()).max()))
for k, p in mod.named_parameters():
    g1 = p.grad.cpu().double().numpy(); gs = got["grad/" + k].astype(np.float64); g64 = r64[2][k].numpy()
    print("%-34s max|ref| %.2e  shard-1gpu %.2e  shard-f64 %.2e  1gpu-f64 %.2e" % (k, np.abs(g64).max(), np.abs(gs - g1).max(), np.abs(gs - g64).max(), np.abs(g1 - g64).max()))
