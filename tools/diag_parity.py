"""Where does the end-to-end error come from? One small train step under different code-path toggles, max |x - fp64| of the
logits and of the gradients (relative to the tensor's largest magnitude), next to the fp32 oracle's own error."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "link-prediction-gnn_b200"), os.path.join(ROOT, "tests")]
from oracle import twowl_oracle as O
import TwoWL.model.model as model
import TwoWL.utils as U
from twowl_b200 import ops

def case(n, m, hidden, seed, skew):
    rng = np.random.default_rng(seed)
    und = np.minimum((rng.pareto(1.5, size=(2, m)) * n / 30).astype(np.int64), n - 1) if skew else rng.integers(0, n, size=(2, m))
    pos, pred = O.synthetic_split(n, und, seed)
    E = pos.shape[1]; nb = max(2, (E // 2) // 10)
    idx1 = O.double(rng.permutation(E // 2)[:nb], for_index=True)
    idx2 = O.double(rng.permutation(pred.shape[1] // 2)[:nb], for_index=True) + E
    pos1 = np.concatenate([pos.T, pred.T]); y = torch.cat((torch.ones(nb), torch.zeros(nb))).unsqueeze(-1)
    return pos, pred, pos1, idx1, idx2, y

d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
for (n, m, hidden, skew) in [(400, 1200, 32, False), (800, 2500, 128, True), (3000, 9000, 64, True), (20000, 80000, 64, True)]:
    pos, pred, pos1, idx1, idx2, y = case(n, m, hidden, 0, skew)
    ei2 = O.get_ei2(n, pos, pred)
    o_ei, o_x, o_ei2 = O.sample_block(idx1, n, pos, ei2)
    sd = O.init_state_dict(int(O.degree(pos, n).max()), hidden, hidden, 1, 1, seed=0)
    pos2 = np.concatenate([idx1, idx2])
    args = (torch.from_numpy(o_x), torch.from_numpy(o_ei), torch.from_numpy(pos1), torch.from_numpy(pos2), o_ei2, y)
    p32, l32, g32 = O.fwd_bwd(sd, *args)
    p64, l64, g64 = O.fwd_bwd({k: v.double() for k, v in sd.items()}, *args[:-1], y.double())
    def gerr(g):
        return max(float((g[k].double() - g64[k]).abs().max() / g64[k].abs().max()) for k in g64 if float(g64[k].abs().max()) > 1e-12)
    print(f"--- n={n} m={m} hidden={hidden} skew={skew}: fp32 oracle logits err {float((p32.double()-p64).abs().max()):.2e}, grads rel {gerr(g32):.2e}")
    dei2 = U.get_ei2(n, d(pos), d(pred))
    ei_new, x_new, ei2_new = U.sample_block(d(idx1), n, d(pos), dei2)
    for name, impl, fused, fro, path in [("fused tc", 2, True, True, "structured"), ("fused tc, no fused readout", 2, True, False, "structured"),
                                         ("op-by-op tc", 2, False, False, "structured"), ("op-by-op simt", 0, False, False, "structured"),
                                         ("explicit tc", 2, False, False, "explicit"), ("explicit simt", 0, False, False, "explicit")]:
        ops.LINEAR_IMPL = impl
        mod = model.LocalWLNet(sd["emb.0.weight"].shape[0] - 1, False, None, hidden, hidden, 1, 1, 0., 0., 0., 0., 0., 0.)
        mod.load_state_dict(sd); mod = mod.cuda().train()
        mod.pair_path, mod.fused_pair_layer, mod.fused_readout = path, fused, fro
        out = mod(x_new, ei_new, d(pos1), d(pos2), ei2_new)
        loss = torch.nn.functional.binary_cross_entropy_with_logits(out, y.cuda()); loss.backward()
        g = {k: p.grad.cpu() for k, p in mod.named_parameters()}
        worst = max((k for k in g64 if float(g64[k].abs().max()) > 1e-12), key=lambda k: float((g[k].double() - g64[k]).abs().max() / g64[k].abs().max()))
        print(f"   {name:28s} logits err {float((out.cpu().double()-p64).abs().max()):.2e}  grads rel {gerr(g):.2e} ({worst})")
