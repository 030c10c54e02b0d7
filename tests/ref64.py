"""TEST INFRASTRUCTURE ONLY - a float64 evaluation of one TwoWL train step (LocalWLNet.forward of model.py:68-84 with
depth1 = depth2 = 1 and dropout 0, binary_cross_entropy_with_logits, backward: train.py:36-38) that runs at sizes where neither
the reference nor the CPU oracle can: the explicit wedge index of R-MAT 1M/16M has 6.5e10 columns.

Written with plain torch ops only (index_select / index_add_ / matmul / autograd; nothing from the package under test) on
whatever device the inputs live on. The pair-level GCNConv is evaluated in its per-node form

    out[t] = dinv[t] * S[centre(t)] + [t has no live id-self-loop] * dinv[t]^2 * Z[t] + bias,
    S[i]   = sum over the live in-edges a of i of dinv[row(a)] * Z[row(a)],     row(a) = a^1 (edge2) or a (edge2_r)

and everything of size [R, C] is produced, used and dropped one row chunk at a time. The three kinds of global coupling - the
per-node sums S, GraphNorm's column moments, and the rows the readout selects - are explicit variables, so the backward is
torch autograd applied chunk by chunk:

    A  node-level part under autograd                    -> h [N, C]
    B  (no grad) S_f, S_r over all chunks
    C  (no grad) column moments s1 = sum out, s2 = sum out^2 and the selected rows of out_f / out_r
    D  autograd on the small head (GraphNorm at the selected rows, ReLU, branch sum, readout, BCE)
                                                          -> d sel, d s1, d s2, GraphNorm / pred parameter gradients
    E  per chunk, autograd of out_c(h, W, bias, S) seeded with  d out_c = scatter(d sel) + d s1 + 2 out_c d s2
                                                          -> accumulates d h, d W, d bias, d S
    F  per chunk, autograd of the chunk's share of S(h, W) seeded with d S -> accumulates d h, d W
    G  backward of the node-level part seeded with d h

tests/test_ref64_cpu.py pins this file to the CPU oracle (explicit [2,T] index, plain autograd) on small graphs with chunks much
smaller than the graph; tests/test_gpu_fullsize.py uses it as the reference at BASELINE configs[3]'s full size.
"""
from __future__ import annotations

import torch

EPS = 1e-5


def _gn(x, w, b, a):
    mean = x.mean(0, keepdim=True)
    out = x - mean * a
    return w * out / (out.pow(2).mean(0, keepdim=True) + EPS).sqrt() + b


def _node_part(p, deg, ei, n, dt=torch.float64):
    """model.py:71-73 with depth1 = 1: Embedding -> GraphNorm -> GCNConv -> GraphNorm -> ReLU."""
    h = _gn(p["emb.0.weight"].index_select(0, deg), p["emb.1.weight"], p["emb.1.bias"], p["emb.1.mean_scale"])
    src, dst = ei[0], ei[1]
    keep = src != dst                                   # PyG add_remaining_self_loops: existing loops out, one weight-1 loop per node in
    src, dst = src[keep], dst[keep]
    one = torch.ones(dst.numel(), dtype=dt, device=h.device)
    d = torch.ones(n, dtype=dt, device=h.device).index_add_(0, dst, one)
    dis = d.pow(-0.5)
    z = h @ p["conv1s.0.modlist.0.lin.weight"].t()
    adj = torch.sparse_coo_tensor(torch.stack((dst, src)), dis[src] * dis[dst], (n, n)).coalesce()
    out = (dis * dis).unsqueeze(1) * z + torch.sparse.mm(adj, z) + p["conv1s.0.modlist.0.bias"]
    return torch.relu(_gn(out, p["conv1s.0.modlist.1.weight"], p["conv1s.0.modlist.1.bias"], p["conv1s.0.modlist.1.mean_scale"]))


class _Dir:
    """Per-row constants of one direction of the pair layer (edge2 = [a^1; b] or edge2_r = [a; b^1], utils.py:71-78)."""

    def __init__(self, fwd, pos1, E, live_all, cnt, dt=torch.float64):
        R = pos1.shape[0]
        rows = torch.arange(R, device=pos1.device)
        psrc = pos1[:, 0]
        self.fwd = fwd
        self.centre = psrc if fwd else psrc[rows ^ 1]
        has_loop = live_all[rows ^ 1] if fwd else live_all          # the wedge (t, t) exists iff that id is a live observed edge
        self.dinv = (cnt[self.centre] - has_loop.to(dt) + 1.0).pow(-0.5)
        self.coef = (~has_loop).to(dt) * self.dinv * self.dinv
        self.pre = "conv2s.0." if fwd else "conv2s_r.0."


def step(sd, deg, ei, pos1, idx, E, blocked, y, chunk_rows=1 << 22, dtype=torch.float64, debug=None):
    """-> (logits [L,1], loss, {state_dict key: gradient}), all float64 (`dtype=torch.float32` evaluates the same program in
    fp32: what plain torch fp32 arithmetic - the reference's - makes of these formulas at this size). sd: state_dict of a depth-1/1 LocalWLNet;
    deg / ei: the node-level inputs AFTER sample_block (x_new, ei_new); pos1 int64 [R,2]; idx int64 [2L]; E = number of observed
    edge rows; blocked bool [E] = the sampled edge ids (their wedges are gone, utils.py:48-50); y [L,1]."""
    dev = pos1.device
    dt = dtype
    p = {k: v.detach().to(dev, dt).clone().requires_grad_(True) for k, v in sd.items()}
    n = deg.numel()
    R = pos1.shape[0]
    assert R % 2 == 0 and E % 2 == 0 and chunk_rows % 2 == 0
    C = p["pred.weight"].shape[1]
    live = ~blocked
    dst_e = pos1[:E, 1]
    cnt = torch.zeros(n, dtype=dt, device=dev).index_add_(0, dst_e[live], torch.ones(int(live.sum()), dtype=dt, device=dev))
    live_all = torch.zeros(R, dtype=torch.bool, device=dev)
    live_all[:E] = live
    dirs = [_Dir(True, pos1, E, live_all, cnt, dt), _Dir(False, pos1, E, live_all, cnt, dt)]
    chunks = [(lo, min(lo + chunk_rows, R)) for lo in range(0, R, chunk_rows)]
    idx = idx.reshape(-1)
    order = torch.argsort(idx)
    idx_sorted = idx[order]

    # ---- A
    h = _node_part(p, deg, ei, n, dt)
    hd = h.detach().requires_grad_(True)

    def H_of(hh, lo, hi):
        return hh.index_select(0, pos1[lo:hi, 0]) * hh.index_select(0, pos1[lo:hi, 1])                 # model.py:75

    def S_share(d, Z, lo, hi):
        """This chunk's rows' contribution to S_d [n, C]: live in-edges a in [lo, min(hi, E))."""
        hiE = min(hi, E)
        out = torch.zeros(n, C, dtype=dt, device=dev)
        if hiE <= lo:
            return out
        a = lo + torch.nonzero(live[lo:hiE]).reshape(-1)
        srow = (a ^ 1) if d.fwd else a
        return out.index_add(0, dst_e[a], d.dinv[srow].unsqueeze(1) * Z.index_select(0, srow - lo))

    def out_of(d, Z, S, lo, hi, par):
        return (d.dinv[lo:hi].unsqueeze(1) * S.index_select(0, d.centre[lo:hi]) + d.coef[lo:hi].unsqueeze(1) * Z
                + par[d.pre + "modlist.0.bias"])

    # ---- B
    S = [torch.zeros(n, C, dtype=dt, device=dev) for _ in dirs]
    with torch.no_grad():
        for lo, hi in chunks:
            Hc = H_of(hd, lo, hi)
            for k, d in enumerate(dirs):
                S[k] += S_share(d, Hc @ p[d.pre + "modlist.0.lin.weight"].t(), lo, hi)
    # ---- C
    s1 = [torch.zeros(C, dtype=dt, device=dev) for _ in dirs]
    s2 = [torch.zeros(C, dtype=dt, device=dev) for _ in dirs]
    sel = [torch.zeros(idx.numel(), C, dtype=dt, device=dev) for _ in dirs]
    with torch.no_grad():
        for lo, hi in chunks:
            Hc = H_of(hd, lo, hi)
            j0, j1 = (int(v) for v in torch.searchsorted(idx_sorted, torch.tensor([lo, hi], device=dev)))
            for k, d in enumerate(dirs):
                o = out_of(d, Hc @ p[d.pre + "modlist.0.lin.weight"].t(), S[k], lo, hi, p)
                s1[k] += o.sum(0)
                s2[k] += (o * o).sum(0)
                if j1 > j0:
                    sel[k][order[j0:j1]] = o.index_select(0, idx_sorted[j0:j1] - lo)
    # ---- C2 (no grad): the variance the way GraphNorm computes it - the mean first, then the mean square of the shifted
    # values - so that the fp32 evaluation does not suffer the cancellation of the one-pass form (in float64 both agree)
    v2 = [torch.zeros(C, dtype=dt, device=dev) for _ in dirs]
    with torch.no_grad():
        for lo, hi in chunks:
            Hc = H_of(hd, lo, hi)
            for k, d in enumerate(dirs):
                o = out_of(d, Hc @ p[d.pre + "modlist.0.lin.weight"].t(), S[k], lo, hi, p)
                v2[k] += (o - p[d.pre + "modlist.1.mean_scale"] * (s1[k] / R)).pow(2).sum(0)
    # ---- D
    for t in s1 + s2 + sel + S:
        t.requires_grad_(True)
    branches = []
    for k, d in enumerate(dirs):
        a = p[d.pre + "modlist.1.mean_scale"]
        mean = s1[k] / R
        var = s2[k] / R - (2 * a - a * a) * mean * mean                     # E[(x - a m)^2] = E[x^2] - (2a - a^2) m^2
        var = var + (v2[k] / R - var).detach()                              # the two-pass VALUE, the one-pass form's gradient
        branches.append(torch.relu(p[d.pre + "modlist.1.weight"] * (sel[k] - a * mean) / (var + EPS).sqrt() + p[d.pre + "modlist.1.bias"]))
    hsel = branches[0] + branches[1]
    if debug is not None:       # intermediates for tools/diag_stages.py: where along the chain an implementation's error appears
        debug.update(h=h.detach(), S=[t.detach() for t in S], mean=[(s1[k] / R).detach() for k in range(2)],
                     var=[(v2[k] / R).detach() for k in range(2)], sel=[t.detach() for t in sel], hsel=hsel.detach())
    logits = (hsel[0::2] * hsel[1::2]) @ p["pred.weight"].t() + p["pred.bias"]            # model.py:78-83
    loss = torch.nn.functional.binary_cross_entropy_with_logits(logits, y.to(dev, dt))
    loss.backward()
    # ---- E
    for lo, hi in chunks:
        Hc = H_of(hd, lo, hi)
        j0, j1 = (int(v) for v in torch.searchsorted(idx_sorted, torch.tensor([lo, hi], device=dev)))
        outs, gouts = [], []
        for k, d in enumerate(dirs):
            o = out_of(d, Hc @ p[d.pre + "modlist.0.lin.weight"].t(), S[k], lo, hi, p)
            g = s1[k].grad.unsqueeze(0) + 2.0 * o.detach() * s2[k].grad.unsqueeze(0)
            if j1 > j0:
                g.index_add_(0, idx_sorted[j0:j1] - lo, sel[k].grad.index_select(0, order[j0:j1]))
            outs.append(o)
            gouts.append(g)
        torch.autograd.backward(outs, gouts)
        del Hc, outs, gouts
    # ---- F
    for lo, hi in chunks:
        if lo >= E:
            break
        Hc = H_of(hd, lo, hi)
        shares = [S_share(d, Hc @ p[d.pre + "modlist.0.lin.weight"].t(), lo, hi) for d in dirs]
        torch.autograd.backward(shares, [S[0].grad, S[1].grad])
        del Hc, shares
    # ---- G
    h.backward(hd.grad)
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in p.items()}
    return logits.detach(), loss.detach(), grads
