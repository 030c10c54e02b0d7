"""Where does the end-to-end error of the FULL-SIZE step (tests/test_gpu_fullsize.py, rmat / hidden 64) come from? The same
step under the product's code-path toggles against the float64 evaluation (tests/ref64.py), next to two runs of the torch-fp32
evaluation of the same program (its atomics make its own error move from run to run).
    python tools/diag_fullsize.py [workload] [hidden]"""
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
sd = {k: v.detach() for k, v in mod.state_dict().items()}


def ours():
    ei2 = U.get_ei2_implicit(n, pos, pred)
    ei_new, x_new, ei2_new = U.sample_block(idx1, n, pos, ei2)
    for p in mod.parameters():
        p.grad = None
    out = mod(x_new, ei_new, pos1, idx, ei2_new)
    torch.nn.functional.binary_cross_entropy_with_logits(out, y).backward()
    res = out.detach().cpu().double(), {k: p.grad.detach().cpu().double() for k, p in mod.named_parameters()}
    del out, ei2, ei_new, ei2_new
    G.clear_cache(); gc.collect(); torch.cuda.empty_cache()
    return res


ei2 = U.get_ei2_implicit(n, pos, pred)
ei_new, x_new, _ = U.sample_block(idx1, n, pos, ei2)
ei_plain = ei_new.clone()
blocked = torch.zeros(E, dtype=torch.bool, device=dev)
blocked[idx1] = True
del ei2, ei_new, _
G.clear_cache(); gc.collect(); torch.cuda.empty_cache()
lg64, l64, g64 = ref64.step(sd, x_new, ei_plain, pos1, idx, E, blocked, y)
lg64 = lg64.cpu().double(); g64 = {k: v.cpu().double() for k, v in g64.items()}


def report(name, lg, gr):
    le = float((lg - lg64).abs().max())
    rel = {k: float((gr[k] - g64[k]).abs().max() / g64[k].abs().max()) for k in g64 if float(g64[k].abs().max()) > 1e-12}
    top = sorted(rel, key=lambda k: -rel[k])[:4]
    band = lambda a, b: float(((a - b).abs() > 1e-6 + 1e-5 * b.abs()).double().mean())
    out_l = band(lg, lg64)
    out_g = {k: band(gr[k], g64[k]) for k in g64}
    wk = max(out_g, key=lambda k: out_g[k])
    print(f"   outside the float64 band: logits {100 * out_l:.2f} %, all gradient elements "
          f"{100 * sum(out_g[k] * g64[k].numel() for k in g64) / sum(v.numel() for v in g64.values()):.3f} %, worst tensor {wk} {100 * out_g[wk]:.1f} %")
    print(f"{name:34s} logits {le:.2e} (max |logit| {float(lg64.abs().max()):.1f}) | grads rel-to-max: " +
          ", ".join(f"{k.replace('.modlist', '')} {rel[k]:.1e}" for k in top), flush=True)


for t in range(2):
    lg32, l32, g32 = ref64.step(sd, x_new, ei_plain, pos1, idx, E, blocked, y, dtype=torch.float32)
    report(f"torch fp32 evaluation, run {t}", lg32.cpu().double(), {k: v.cpu().double() for k, v in g32.items()})
    del lg32, g32
gc.collect(); torch.cuda.empty_cache()

for name, impl, fused, fro, loc in [("product (tc, fused, regrouped)", 2, True, True, True), ("no regrouping by hub", 2, True, True, False),
                                    ("no fused readout", 2, True, False, True), ("op-by-op tc", 2, False, False, True),
                                    ("op-by-op simt fp32", 0, False, False, True), ("op-by-op simt, no regrouping", 0, False, False, False)][:int(os.environ.get("DIAG_PATHS", "6"))]:
    ops.LINEAR_IMPL = impl
    mod.fused_pair_layer, mod.fused_readout, mod.pair_locality = fused, fro, loc
    try:
        report(name, *ours())
    except Exception as e:      # e.g. out of memory on the un-fused paths at this size
        print(f"{name:34s} FAILED: {type(e).__name__}: {str(e)[:200]}", flush=True)
        G.clear_cache(); gc.collect(); torch.cuda.empty_cache()
