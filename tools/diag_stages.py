"""Where along the chain does the full-size step's error appear? Intermediates of the product's forward (captured by wrapping
the ops it calls) against the float64 evaluation (tests/ref64.py, debug=...), next to the torch-fp32 evaluation's own:
node-level output h, the per-node tables (S W^T), the pre-GraphNorm outputs at the selected rows, GraphNorm's column mean /
variance, the post-GraphNorm rows and the logits. Errors are max |x - fp64| / max |fp64| and the share of elements outside
the 1e-6 + 1e-5 |b| band.   python tools/diag_stages.py [workload] [hidden]"""
import gc, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "link-prediction-gnn_b200"), os.path.join(ROOT, "tests")]
import torch
import bench, ref64
import TwoWL.model.model as model
import TwoWL.utils as U
from twowl_b200 import graph as G, ops

wl = sys.argv[1] if len(sys.argv) > 1 else "rmat"
hidden = int(sys.argv[2]) if len(sys.argv) > 2 else 64
dev = torch.device("cuda", 0)
g = bench.make_graph(wl, 0, dev)
n, pos, pred, pos1 = g["n"], g["pos"], g["pred"], g["pos1"]
E, P = pos.shape[1], pred.shape[1]
nb = g["und"] // 10
i1, i2, y = (t.to(dev) for t in bench.draw_batch(g["und"], P // 2, nb, 0))
idx1 = U.double(i1, for_index=True)
idx = torch.cat((idx1, U.double(i2, for_index=True) + E))
torch.manual_seed(3)
mod = model.LocalWLNet(int(U.degree(pos, n).max().item()), False, None, hidden, hidden, 1, 1, 0., 0., 0., 0., 0., 0.)
with torch.no_grad():
    for p in mod.parameters():
        if p.dim() == 1:
            p.add_(0.2 * torch.randn_like(p))
mod = mod.to(dev).train()
mod.pair_locality = False          # original row / node ids, so that the tables line up with the reference's
sd = {k: v.detach() for k, v in mod.state_dict().items()}

ei2 = U.get_ei2_implicit(n, pos, pred)
ei_new, x_new, ei2_new = U.sample_block(idx1, n, pos, ei2)
blocked = torch.zeros(E, dtype=torch.bool, device=dev)
blocked[idx1] = True

# ---- the product's forward with its intermediates captured
cap = {"pc": [], "lin": [], "pinit": []}
_pc, _lin, _pi = ops.pair_conv, ops.linear_fwd, ops.pair_init_fwd


def pc(*a, **k):
    r = _pc(*a, **k)
    cap["pc"].append(r)
    return r


def lin(*a, **k):
    r = _lin(*a, **k)
    cap["lin"].append(r)
    return r


def pinit(x, *a, **k):
    cap["h"] = x
    return _pi(x, *a, **k)


ops.pair_conv, ops.linear_fwd, ops.pair_init_fwd = pc, lin, pinit
with torch.no_grad():
    out = mod(x_new, ei_new, pos1, idx, ei2_new)
ops.pair_conv, ops.linear_fwd, ops.pair_init_fwd = _pc, _lin, _pi
ours = {"h": cap["h"].double().cpu(), "logits": out.double().cpu()}
big = [r for r in cap["pc"] if isinstance(r, tuple) and r[0].shape[0] == pos1.shape[0]]
assert len(big) == 2, [tuple(r[0].shape) if isinstance(r, tuple) else tuple(r.shape) for r in cap["pc"]]
ours["sel"] = [big[k][0][idx].double().cpu() for k in range(2)]
ours["stats"] = [big[k][1].double().cpu() for k in range(2)]
tabs = [r for r in cap["lin"] if r.shape[0] == n and r.shape[1] == hidden]
ours["S"] = [t.double().cpu() for t in tabs[-2:]]
del cap, big, tabs, out, ei2, ei2_new
ei_plain = ei_new.clone()
del ei_new
G.clear_cache(); gc.collect(); torch.cuda.empty_cache()

d64 = {}
lg64, _, _ = ref64.step(sd, x_new, ei_plain, pos1, idx, E, blocked, y, debug=d64)
d64 = {k: ([t.double().cpu() for t in v] if isinstance(v, list) else v.double().cpu()) for k, v in d64.items()}
d64["logits"] = lg64.double().cpu()
gc.collect(); torch.cuda.empty_cache()
d32 = {}
lg32, _, _ = ref64.step(sd, x_new, ei_plain, pos1, idx, E, blocked, y, dtype=torch.float32, debug=d32)
d32 = {k: ([t.double().cpu() for t in v] if isinstance(v, list) else v.double().cpu()) for k, v in d32.items()}
d32["logits"] = lg32.double().cpu()


def err(a, b):
    e = (a - b).abs()
    return f"max/max {float(e.max() / b.abs().max()):.2e}, rms/rms {float(e.pow(2).mean().sqrt() / b.pow(2).mean().sqrt()):.2e}, outside band {100 * float((e > 1e-6 + 1e-5 * b.abs()).double().mean()):.2f} %"


print(f"{wl} hidden {hidden}: stage | product vs fp64 | torch fp32 vs fp64")
print("h (node level)      |", err(ours["h"], d64["h"]), "|", err(d32["h"], d64["h"]))
for k, nm in enumerate(("fwd", "rev")):
    print(f"S W^T {nm}           |", err(ours["S"][k], d64["S"][k]), "|", err(d32["S"][k], d64["S"][k]))
for k, nm in enumerate(("fwd", "rev")):
    print(f"O {nm} at idx        |", err(ours["sel"][k], d64["sel"][k]), "|", err(d32["sel"][k], d64["sel"][k]))
C = hidden
for k, nm in enumerate(("fwd", "rev")):
    st = ours["stats"][k]
    print(f"stats {nm} (product's [2C] vector: first C / last C) vs fp64 mean, var, 1/sqrt(var+eps):")
    for name, ref in (("mean", d64["mean"][k]), ("var", d64["var"][k]), ("rstd", (d64["var"][k] + 1e-5).rsqrt())):
        for part, v in (("first", st[:C]), ("last", st[C:2 * C])):
            print(f"     {part} half vs {name:5s}: max rel {float(((v - ref).abs() / ref.abs().clamp_min(1e-30)).max()):.2e}")
    print(f"     torch fp32 mean: max rel {float(((d32['mean'][k] - d64['mean'][k]).abs() / d64['mean'][k].abs()).max()):.2e}, "
          f"var: {float(((d32['var'][k] - d64['var'][k]).abs() / d64['var'][k]).max()):.2e}; "
          f"|mean|/std max {float((d64['mean'][k].abs() / d64['var'][k].sqrt()).max()):.2f}")
print("logits              |", err(ours["logits"], d64["logits"]), "|", err(d32["logits"], d64["logits"]))
