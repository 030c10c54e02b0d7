"""Node-level part of the full-size step (model.py:71-73: Embedding -> GraphNorm -> GCNConv -> GraphNorm -> ReLU on N = 1 M
nodes), stage by stage: the product's kernels (tensor-core split-tf32 or exact-fp32 SIMT linear layers) and plain torch fp32
against float64.   python tools/diag_node.py [workload] [hidden]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "link-prediction-gnn_b200"), os.path.join(ROOT, "tests")]
import torch
import bench, ref64
import TwoWL.model.model as model
import TwoWL.utils as U
from twowl_b200 import ops

wl = sys.argv[1] if len(sys.argv) > 1 else "rmat"
hidden = int(sys.argv[2]) if len(sys.argv) > 2 else 64
dev = torch.device("cuda", 0)
g = bench.make_graph(wl, 0, dev)
n, pos, pred, pos1 = g["n"], g["pos"], g["pred"], g["pos1"]
E, P = pos.shape[1], pred.shape[1]
nb = g["und"] // 10
i1, i2, y = (t.to(dev) for t in bench.draw_batch(g["und"], P // 2, nb, 0))
idx1 = U.double(i1, for_index=True)
torch.manual_seed(3)
mod = model.LocalWLNet(int(U.degree(pos, n).max().item()), False, None, hidden, hidden, 1, 1, 0., 0., 0., 0., 0., 0.)
with torch.no_grad():
    for p in mod.parameters():
        if p.dim() == 1:
            p.add_(0.2 * torch.randn_like(p))
mod = mod.to(dev).train()
sd = {k: v.detach() for k, v in mod.state_dict().items()}
ei2 = U.get_ei2_implicit(n, pos, pred)
ei_new, x_new, ei2_new = U.sample_block(idx1, n, pos, ei2)
ei_plain = ei_new.clone()


def ref_stages(dt):
    p = {k: v.to(dt) for k, v in sd.items()}
    e0 = p["emb.0.weight"].index_select(0, x_new)
    x1 = ref64._gn(e0, p["emb.1.weight"], p["emb.1.bias"], p["emb.1.mean_scale"])
    src, dst = ei_plain[0], ei_plain[1]
    keep = src != dst
    src, dst = src[keep], dst[keep]
    d = torch.ones(n, dtype=dt, device=dev).index_add_(0, dst, torch.ones(dst.numel(), dtype=dt, device=dev))
    dis = d.pow(-0.5)
    z = x1 @ p["conv1s.0.modlist.0.lin.weight"].t()
    adj = torch.sparse_coo_tensor(torch.stack((dst, src)), dis[src] * dis[dst], (n, n)).coalesce()
    agg = (dis * dis).unsqueeze(1) * z + torch.sparse.mm(adj, z) + p["conv1s.0.modlist.0.bias"]
    h = torch.relu(ref64._gn(agg, p["conv1s.0.modlist.1.weight"], p["conv1s.0.modlist.1.bias"], p["conv1s.0.modlist.1.mean_scale"]))
    return dict(x1=x1.double(), z=z.double(), agg=agg.double(), h=h.double())


class _Stop(Exception):
    pass


def our_stages(impl):
    ops.LINEAR_IMPL = impl
    cap = {}
    _lin, _seg, _gna, _pi = ops.linear_fwd, ops.seg_reduce, ops.graphnorm_apply, ops.pair_init_fwd

    def lin(*a, **k):
        r = _lin(*a, **k)
        cap.setdefault("z", r)
        return r

    def seg(*a, **k):
        r = _seg(*a, **k)
        cap.setdefault("agg", r)
        return r

    def gna(*a, **k):
        r = _gna(*a, **k)
        cap["x1" if "x1" not in cap else "h"] = r
        return r

    def pi(x, *a, **k):
        raise _Stop

    ops.linear_fwd, ops.seg_reduce, ops.graphnorm_apply, ops.pair_init_fwd = lin, seg, gna, pi
    try:
        with torch.no_grad():
            mod(x_new, ei_new, pos1, torch.cat((idx1, idx1)), ei2_new)
    except _Stop:
        pass
    finally:
        ops.linear_fwd, ops.seg_reduce, ops.graphnorm_apply, ops.pair_init_fwd = _lin, _seg, _gna, _pi
    return {k: (v[0] if isinstance(v, tuple) else v).double() for k, v in cap.items()}


def err(a, b):
    e = (a - b).abs()
    return f"rms/rms {float(e.pow(2).mean().sqrt() / b.pow(2).mean().sqrt()):.2e} max/max {float(e.max() / b.abs().max()):.2e} out-of-band {100 * float((e > 1e-6 + 1e-5 * b.abs()).double().mean()):.2f} %"


r64, r32 = ref_stages(torch.float64), ref_stages(torch.float32)
o2, o0 = our_stages(2), our_stages(0)
print(f"{wl} hidden {hidden}, N = {n}: stage | torch fp32 | product, tensor-core linear | product, SIMT fp32 linear")
for k in ("x1", "z", "agg", "h"):
    print(f"{k:4s} | {err(r32[k], r64[k])} | {err(o2[k], r64[k]) if k in o2 else 'n/a'} | {err(o0[k], r64[k]) if k in o0 else 'n/a'}")
# the same with the product's stage k fed by the float64 result of stage k-1 rounded to fp32: each kernel's own error
p32 = {k: v.float().contiguous() for k, v in sd.items()}
x1_in = r64["x1"].float().contiguous()
for impl in (2, 0):
    ops.LINEAR_IMPL = impl
    z = ops.linear_fwd(x1_in, p32["conv1s.0.modlist.0.lin.weight"].contiguous())
    zr = x1_in.double() @ sd["conv1s.0.modlist.0.lin.weight"].double().t()
    print(f"linear alone (impl {impl}) on exact inputs: {err(z.double(), zr)}")
agg_in = r64["agg"].float().contiguous()
gn = mod.conv1s[0].modlist[1]
st = ops.graphnorm_stats(agg_in, gn.mean_scale.detach(), gn.eps)
hh = ops.graphnorm_apply(agg_in, st, gn.weight.detach(), gn.bias.detach(), gn.mean_scale.detach(), 0.0, 0, True)
hh = hh[0] if isinstance(hh, tuple) else hh
hr = torch.relu(ref64._gn(agg_in.double(), sd["conv1s.0.modlist.1.weight"].double(), sd["conv1s.0.modlist.1.bias"].double(), sd["conv1s.0.modlist.1.mean_scale"].double()))
h32 = torch.relu(ref64._gn(agg_in, p32["conv1s.0.modlist.1.weight"], p32["conv1s.0.modlist.1.bias"], p32["conv1s.0.modlist.1.mean_scale"]))
print(f"GraphNorm+ReLU alone on exact inputs: product {err(hh.double(), hr)} | torch fp32 {err(h32.double(), hr)}")
