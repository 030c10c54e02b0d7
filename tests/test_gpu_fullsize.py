"""BASELINE configs[3] at FULL size (R-MAT 1 M nodes / 16 M edge samples -> E = 30 M observed edge rows, R = 60 M pair rows,
3 M target links per step): the CUDA path against an independent fp64 evaluation and size-independent properties.

The CPU oracle cannot run here (the explicit wedge index has 6.5e10 columns), so the reference is a float64 evaluation of
LocalWLNet.forward (model.py:68-84) written below with plain torch ops on the device - index_add_ / index_select over the
edge lists, nothing from the package - using the per-node form of the pair-level GCNConv (sum over a centre's live in-edges),
which tests/test_gpu_model.py pins against the explicit-index oracle at the sizes the oracle can run. Checked:
  * logits and loss vs fp64 (north_star tolerance with an absolute floor of 1e-5 x the largest logit),
  * bit-identical logits and gradients on a second run (no atomics on data anywhere),
  * every parameter gradient through central differences of the fp64 loss, one parameter group at a time, along our gradient
    with randomly rescaled elements: dL/d(eps) of the independent forward must equal <our gradient, direction> to 2e-3 relative.
"""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HIDDEN = 32          # fp64 [R, C] reference tensors are 15.4 GB each at C = 32


def _gn(x, w, b, a, eps=1e-5):
    mean = x.mean(0, keepdim=True)
    out = x - mean * a
    return w * out / (out.pow(2).mean(0, keepdim=True) + eps).sqrt() + b


def _ref_forward(sd, deg, ei, pos1, idx, E, blocked):
    """float64 LocalWLNet.forward, depth1 = depth2 = 1, dropout 0. blocked: bool [E], the sampled edge ids."""
    n = deg.numel()
    h = _gn(sd["emb.0.weight"].index_select(0, deg), sd["emb.1.weight"], sd["emb.1.bias"], sd["emb.1.mean_scale"])
    # node-level GCNConv (model.py:73): add self loops, symmetric normalisation by in-degree
    src, dst = ei[0], ei[1]
    d = torch.ones(n, dtype=torch.float64, device=h.device).index_add_(0, dst, torch.ones_like(dst, dtype=torch.float64))
    dis = d.pow(-0.5)
    z = h @ sd["conv1s.0.modlist.0.lin.weight"].t()
    out = (dis * dis).unsqueeze(1) * z
    out.index_add_(0, dst, (dis[src] * dis[dst]).unsqueeze(1) * z.index_select(0, src))
    h = torch.relu(_gn(out + sd["conv1s.0.modlist.0.bias"], sd["conv1s.0.modlist.1.weight"], sd["conv1s.0.modlist.1.bias"],
                       sd["conv1s.0.modlist.1.mean_scale"]))
    del z, out
    H = h.index_select(0, pos1[:, 0]) * h.index_select(0, pos1[:, 1])                       # model.py:75
    R = H.shape[0]
    rows = torch.arange(R, device=H.device)
    mate = rows ^ 1
    psrc = pos1[:, 0]
    live = ~blocked                                                                         # [E] observed edges still in the graph
    dst_e = pos1[:E, 1]
    cnt = torch.zeros(n, dtype=torch.float64, device=H.device).index_add_(0, dst_e[live], torch.ones(int(live.sum()), dtype=torch.float64,
                                                                                                    device=H.device))
    live_all = torch.zeros(R, dtype=torch.bool, device=H.device)
    live_all[:E] = live
    sel = []
    for pre, fwd in (("conv2s.0.", True), ("conv2s_r.0.", False)):
        # edge2 = [a^1; b]: target row t takes the rows a^1 of the live in-edges a of src[t]; its id-self-loop is a = t^1.
        # edge2_r = [a; b^1]: target row t takes the rows a of the live in-edges a of src[t^1]; its id-self-loop is a = t.
        centre = psrc if fwd else psrc[mate]
        loop_edge = mate if fwd else rows                 # the edge id whose wedge into t is the self loop (t, t)
        has_loop = live_all[loop_edge]                    # it exists iff that id is a live observed edge (its target is the centre)
        degt = cnt[centre] - has_loop.double() + 1.0       # PyG add_remaining_self_loops: loops out, one weight-1 loop in
        dinv = degt.pow(-0.5)
        Z = H @ sd[pre + "modlist.0.lin.weight"].t()
        a_ids = torch.nonzero(live).reshape(-1)           # live in-edges
        srow = (a_ids ^ 1) if fwd else a_ids              # the row each one contributes
        S = torch.zeros(n, Z.shape[1], dtype=torch.float64, device=H.device)
        S.index_add_(0, dst_e[a_ids], dinv[srow].unsqueeze(1) * Z.index_select(0, srow))
        out = dinv.unsqueeze(1) * S.index_select(0, centre)
        # the self-loop wedge is inside S with weight dinv^2: take it out, add the weight-1 (normalised dinv^2) loop -> net zero
        out += ((~has_loop).double() * dinv * dinv).unsqueeze(1) * Z
        out += sd[pre + "modlist.0.bias"]
        del Z, S
        mean = out.mean(0, keepdim=True)
        a = sd[pre + "modlist.1.mean_scale"]
        var = (out - mean * a).pow(2).mean(0, keepdim=True)
        o = out.index_select(0, idx)
        del out
        sel.append(torch.relu(sd[pre + "modlist.1.weight"] * (o - mean * a) / (var + 1e-5).sqrt() + sd[pre + "modlist.1.bias"]))
    hsel = sel[0] + sel[1]
    return (hsel[0::2] * hsel[1::2]) @ sd["pred.weight"].t() + sd["pred.bias"]


@pytest.mark.parametrize("workload,hidden", [("rmat", HIDDEN), ("collab", 64), ("collab", 128), ("collab", 256)])
def test_rmat_full_size_step_vs_fp64_and_properties(workload, hidden):
    """("rmat", 32): BASELINE configs[3] at full size. ("collab", 64 / 128 / 256): configs[2]'s graph (R = 5 M pair rows) at the
    widths of the configs[4] sweep - the column-window / multi-launch paths of the wide layers against the same fp64 reference."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if torch.cuda.get_device_properties(0).total_memory < 150 * 2 ** 30:
        pytest.skip("needs a 180 GB device")
    sys.path.insert(0, ROOT)
    import bench
    import TwoWL.model.model as model
    import TwoWL.utils as U
    from twowl_b200 import graph as G
    G.clear_cache()
    dev = torch.device("cuda", 0)
    g = bench.make_graph(workload, 0, dev)
    n, pos, pred, pos1 = g["n"], g["pos"], g["pred"], g["pos1"]
    E, P = pos.shape[1], pred.shape[1]
    if workload == "rmat":
        assert n == 1 << 20 and E > 29_000_000 and E + P > 59_000_000       # the BASELINE configs[3] sizes
    else:
        assert n == 1 << 18 and E + P > 4_900_000                           # configs[2]: ogbl-collab scale
    ei2 = U.get_ei2_implicit(n, pos, pred)
    nb = g["und"] // 10
    i1, i2, y = (t.to(dev) for t in bench.draw_batch(g["und"], P // 2, nb, 0))
    idx1 = U.double(i1, for_index=True)
    idx = torch.cat((idx1, U.double(i2, for_index=True) + E))
    ei_new, x_new, ei2_new = U.sample_block(idx1, n, pos, ei2)

    torch.manual_seed(3)
    mod = model.LocalWLNet(int(U.degree(pos, n).max().item()), False, None, hidden, hidden, 1, 1, 0., 0., 0., 0., 0., 0.)
    with torch.no_grad():
        for p in mod.parameters():
            if p.dim() == 1:
                p.add_(0.2 * torch.randn_like(p))
    mod = mod.to(dev).train()

    def ours():
        for p in mod.parameters():
            p.grad = None
        out = mod(x_new, ei_new, pos1, idx, ei2_new)
        loss = torch.nn.functional.binary_cross_entropy_with_logits(out, y)
        loss.backward()
        return out.detach().clone(), loss.detach().clone(), {k: p.grad.detach().clone() for k, p in mod.named_parameters()}

    out, loss, grads = ours()
    out2, loss2, grads2 = ours()
    assert torch.equal(out, out2) and torch.equal(loss, loss2), "forward is not deterministic"
    for k in grads:
        assert torch.equal(grads[k], grads2[k]), f"gradient {k} is not deterministic"
    del out2, grads2

    blocked = torch.zeros(E, dtype=torch.bool, device=dev)
    blocked[idx1] = True
    assert ei_new.shape[1] == E - idx1.numel()
    sd = {k: v.detach().double() for k, v in mod.state_dict().items()}

    def ref_loss(sdx):
        with torch.no_grad():
            lg = _ref_forward(sdx, x_new, ei_new, pos1, idx, E, blocked)
            return lg, torch.nn.functional.binary_cross_entropy_with_logits(lg, y.double())

    lg64, l64 = ref_loss(sd)
    err = (out.double() - lg64).abs()
    bound = 1e-6 + 1e-5 * float(lg64.abs().max()) + 1e-5 * lg64.abs()
    assert bool((err <= bound).all()), f"logits: max err {float(err.max()):.3e}, max |ref| {float(lg64.abs().max()):.3e}"
    assert abs(float(loss) - float(l64)) <= 1e-6 + 1e-5 * abs(float(l64))
    del lg64

    # gradients: central differences of the independent fp64 loss along a direction inside one parameter group. The direction is
    # our own gradient with every element rescaled by a random factor in [0.5, 1.5): it overlaps the gradient (so the directional
    # derivative is large against rounding), and a wrong scale, a missing term or a misplaced element all change <error, v>.
    groups = {"emb": ["emb."], "conv1s": ["conv1s."], "conv2s": ["conv2s.0."], "conv2s_r": ["conv2s_r.0."], "pred": ["pred."]}
    gen = torch.Generator(device="cpu").manual_seed(11)
    for name, prefixes in groups.items():
        keys = [k for k in grads if any(k.startswith(p) for p in prefixes)]
        assert keys
        direction = {k: grads[k].double() * (0.5 + torch.rand(sd[k].shape, generator=gen, dtype=torch.float64)).to(dev) for k in keys}
        scale = sum(float((v * v).sum()) for v in direction.values()) ** 0.5
        assert scale > 0, f"{name}: zero gradient"
        direction = {k: v / scale for k, v in direction.items()}
        eps = 1e-4      # small against the ReLU kinks the step crosses (their effect on a central difference is linear in eps)
        plus = {k: (v + eps * direction[k] if k in direction else v) for k, v in sd.items()}
        minus = {k: (v - eps * direction[k] if k in direction else v) for k, v in sd.items()}
        fd = (float(ref_loss(plus)[1]) - float(ref_loss(minus)[1])) / (2 * eps)
        an = sum(float((grads[k].double() * direction[k]).sum()) for k in keys)
        assert abs(fd - an) <= 2e-3 * abs(an) + 1e-9, f"{name}: finite difference {fd:.6e} vs <grad, v> {an:.6e}"
