"""configs[0] (fb-pages-food seed-0 split from tests/golden): wall time of the drop-in train() step (train.py:11-47: batch draw,
edge blocking, forward, BCE, backward, Adam, device AUC, one host read) eager vs with the step replayed as a CUDA graph.
    python tools/bench_train_small.py [steps]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "link-prediction-gnn_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch
from helpers import fb_split
import TwoWL.model.model as model
import TwoWL.model.train as T
import TwoWL.utils as U
from TwoWL.operators.datasets import dataset
from twowl_b200.graphed import GraphedTrainStep

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
g = np.load(os.path.join(ROOT, "tests", "golden", "fb_pages_food_seed0.npz"), allow_pickle=False)
fb = {k: g[k] for k in g.files}
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
n = int(fb["num_nodes"][0])
ei, pred, pos1 = fb_split(fb, 0)
dei, dpred, dpos = dev(ei), dev(pred), dev(pos1)
x = U.degree(dei, n)
ds = dataset(x, None, dei, None, dpos, torch.zeros(1), U.get_ei2(n, dei, dpred))
for name, use_graph in (("eager", False), ("cuda graph", True)):
    torch.manual_seed(0)
    mod = model.LocalWLNet(int(x.max().item()), False, None, channels_1wl=64, channels_2wl=32, depth1=2, depth2=1).cuda()   # default dropouts
    opt = torch.optim.Adam(mod.parameters(), lr=0.01)
    step = GraphedTrainStep(mod, n, dei, dpos, ds.ei2, n_block=384, n_links=384) if use_graph else None
    for _ in range(20):
        T.train(mod, opt, ds, 384, 0, step)
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(steps):
        loss, auc, _ = T.train(mod, opt, ds, 384, 0, step)
    torch.cuda.synchronize()
    dt = (time.time() - t0) / steps
    print(f"{name:11s}: {dt * 1e3:7.3f} ms per train() step of 384 target links ({384 / dt:9.0f} links/s), last loss {loss:.4f} auc {auc:.3f}")
print("reference CPU path (SURVEY 6, measured at survey time): 133 ms per step")
