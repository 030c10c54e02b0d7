"""TEST INFRASTRUCTURE ONLY - CPU restatement (the "oracle") of the TwoWL hot path.

This file is the checker, never the product: only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import it. The product path
(link-prediction-gnn_b200/) never does and fails loudly without its CUDA library.

What it restates (all file:line relative to /root/reference):
  * the graph operators of TwoWL/utils.py:8-90 (degree, set_mul, check_in_set, get_ei2,
    blockei2, idx2mask, sample_block, reverse, double) - integer work, numpy, vectorised
    so that it also finishes at sizes where the reference's O(n*E) loops do not;
  * LocalWLNet.forward of TwoWL/model/model.py:68-84 with the PyG 2.3.1 semantics of
    GCNConv / GraphNorm the reference gets from torch-geometric==2.3.1
    (requirements.txt:20; source not under /root/reference) - torch CPU fp32, autograd
    gives the backward exactly as the reference's own backward is produced.

Pinning: the reference has no tests or golden vectors for this path.
The oracle is pinned instead (a) against the unmodified reference functions imported in
the build container (tests/test_oracle_vs_reference.py, oracle/ref_import.py) and (b)
against fixtures those functions produced, committed under tests/golden/ by
oracle/gen_golden.py. The GCNConv/GraphNorm arithmetic is third-party and absent from
this image, so for those two modules the pin is "restated from the PyG 2.3.1 docs and
consistent with the reference's recorded AUC band" - see DESIGN.md "Oracle".
"""
from __future__ import annotations

import numpy as np
import torch

# --------------------------------------------------------------------------------------
# integer operators (numpy in, numpy out; int64 everywhere like the reference)
# --------------------------------------------------------------------------------------


def _np(a):
    if isinstance(a, torch.Tensor):
        return a.detach().cpu().numpy()
    return np.asarray(a)


def degree(ei, num_node: int):
    """TwoWL/utils.py:8-10 - histogram of the TARGET row ei[1]."""
    ei = _np(ei)
    return np.bincount(ei[1].astype(np.int64), minlength=int(num_node)).astype(np.int64)


def set_mul(a, b):
    """TwoWL/utils.py:13-19 - Cartesian product, a-major, as a [p*q, 2] table."""
    a, b = _np(a).reshape(-1), _np(b).reshape(-1)
    return np.stack([np.repeat(a, b.size), np.tile(b, a.size)], axis=1).astype(np.int64)


def check_in_set(target, set_):
    """TwoWL/utils.py:22-33 - for each target value, HOW MANY entries of set equal it
    (the reference sums the equality matrix, so duplicates in set count twice)."""
    t, s = _np(target).reshape(-1), np.sort(_np(set_).reshape(-1))
    return (np.searchsorted(s, t, side="right") - np.searchsorted(s, t, side="left")).astype(np.int64)


def get_ei2(n_node: int, pos_edge, pred_edge):
    """TwoWL/utils.py:36-45 - the wedge join. For centre node i ascending, every observed
    edge id a with pos_edge[1][a]==i (ascending) x every pair id b with edge[0][b]==i
    (ascending, edge = cat(pos_edge, pred_edge)). Returns int64 [2, T] (row0=a, row1=b).

    Restated as stable sort by centre + prefix sum + segmented Cartesian fill; nodes
    >= n_node never match the reference's range(n_node) loop and are dropped."""
    pos_edge, pred_edge = _np(pos_edge), _np(pred_edge)
    edge_src = np.concatenate([pos_edge[0], pred_edge[0]]).astype(np.int64)
    in_key = pos_edge[1].astype(np.int64)
    n = int(n_node)
    in_ok = np.nonzero((in_key >= 0) & (in_key < n))[0]
    out_ok = np.nonzero((edge_src >= 0) & (edge_src < n))[0]
    in_list = in_ok[np.argsort(in_key[in_ok], kind="stable")]
    out_list = out_ok[np.argsort(edge_src[out_ok], kind="stable")]
    cin = np.bincount(in_key[in_ok], minlength=n).astype(np.int64)
    cout = np.bincount(edge_src[out_ok], minlength=n).astype(np.int64)
    in_ptr = np.concatenate([[0], np.cumsum(cin)])
    out_ptr = np.concatenate([[0], np.cumsum(cout)])
    seg = cin * cout
    off = np.concatenate([[0], np.cumsum(seg)])
    T = int(off[-1])
    if T == 0:
        return np.zeros((2, 0), dtype=np.int64)
    node = np.repeat(np.arange(n, dtype=np.int64), seg)
    local = np.arange(T, dtype=np.int64) - off[node]
    a = in_list[in_ptr[node] + local // cout[node]]
    b = out_list[out_ptr[node] + local % cout[node]]
    return np.stack([a, b]).astype(np.int64)


def get_ei2_loops(n_node: int, pos_edge, pred_edge):
    """The same join written as the reference's literal triple loop - small cases only;
    used to cross-check the vectorised restatement above."""
    pos_edge, pred_edge = _np(pos_edge), _np(pred_edge)
    src = list(pos_edge[0]) + list(pred_edge[0])
    a_out, b_out = [], []
    for i in range(int(n_node)):
        ins = [a for a in range(pos_edge.shape[1]) if pos_edge[1][a] == i]
        outs = [b for b in range(len(src)) if src[b] == i]
        for a in ins:
            for b in outs:
                a_out.append(a)
                b_out.append(b)
    return np.array([a_out, b_out], dtype=np.int64).reshape(2, -1)


def idx2mask(num: int, idx):
    """TwoWL/utils.py:53-57."""
    m = np.zeros(int(num), dtype=bool)
    m[_np(idx).reshape(-1)] = True
    return m


def blockei2(ei2, blocked_idx):
    """TwoWL/utils.py:48-50 - keep the wedges whose SOURCE edge id ei2[0] is not in
    blocked_idx; column order preserved."""
    ei2 = _np(ei2)
    keep = check_in_set(ei2[0], blocked_idx) == 0
    return np.ascontiguousarray(ei2[:, keep])


def sample_block(sample_idx, size: int, ei, ei2=None):
    """TwoWL/utils.py:61-68 - drop the sampled edge ids from ei, recount the degree by
    SOURCE (sparse row sum, dim=1, utils.py:66-67), filter ei2 by blockei2."""
    ei = _np(ei)
    keep = ~idx2mask(ei.shape[1], sample_idx)
    ei_new = np.ascontiguousarray(ei[:, keep])
    x_new = np.bincount(ei_new[0].astype(np.int64), minlength=int(size)).astype(np.int64)
    ei2_new = blockei2(ei2, sample_idx) if ei2 is not None else None
    return ei_new, x_new, ei2_new


def reverse(edge_index):
    """TwoWL/utils.py:71-78 - (+1 if id even else -1) == id ^ 1 on one row each."""
    e = _np(edge_index).astype(np.int64)
    edge = np.stack([e[0] ^ 1, e[1]])
    edge_r = np.stack([e[0], e[1] ^ 1])
    return edge, edge_r


def double(x, for_index: bool = False):
    """TwoWL/utils.py:81-90 - pair k -> directed ids 2k=(r,c), 2k+1=(c,r)."""
    x = _np(x).astype(np.int64)
    if not for_index:
        out = np.empty((2, 2 * x.shape[1]), dtype=np.int64)
        out[0, 0::2], out[0, 1::2] = x[0], x[1]
        out[1, 0::2], out[1, 1::2] = x[1], x[0]
        return out
    x = x.reshape(-1)
    out = np.empty(2 * x.shape[0], dtype=np.int64)
    out[0::2], out[1::2] = 2 * x, 2 * x + 1
    return out


# --------------------------------------------------------------------------------------
# floating-point model (torch CPU fp32; autograd supplies the backward)
# --------------------------------------------------------------------------------------


def gcn_norm(edge_index: torch.Tensor, num_nodes: int, dtype=torch.float32):
    """PyG 2.3.1 gcn_norm with add_remaining_self_loops."""
    keep = edge_index[0] != edge_index[1]
    loop = torch.arange(num_nodes, dtype=torch.long)
    ei = torch.cat([edge_index[:, keep], torch.stack([loop, loop])], dim=1)
    deg = torch.zeros(num_nodes, dtype=dtype).scatter_add_(
        0, ei[1], torch.ones(ei.shape[1], dtype=dtype))
    dis = deg.pow(-0.5)
    dis = dis.masked_fill(dis == float("inf"), 0.0)
    return ei, dis[ei[0]] * dis[ei[1]]


def gcn_conv(x, edge_index, lin_weight, bias):
    """GCNConv.forward as called at TwoWL/model/model.py:37,73,77 (steps 4-5)."""
    ei, w = gcn_norm(edge_index, x.shape[0], x.dtype)
    z = x @ lin_weight.t()
    out = torch.zeros_like(z).index_add_(0, ei[1], w.unsqueeze(1) * z.index_select(0, ei[0]))
    return out + bias


def graph_norm(x, weight, bias, mean_scale, eps: float = 1e-5):
    """GraphNorm.forward with batch=None (TwoWL/model/model.py:38,54)."""
    mean = x.mean(dim=0, keepdim=True)
    out = x - mean * mean_scale
    var = out.pow(2).mean(dim=0, keepdim=True)
    return weight * out / (var + eps).sqrt() + bias


def _seq(sd, prefix, x, edge_index, act: bool):
    """Seq([GCNConv, GraphNorm, Dropout, ReLU|Identity]) of model.py:36-41,87-96 in eval
    mode / dropout 0 (parity runs never draw dropout masks: CPU and CUDA RNG differ)."""
    x = gcn_conv(x, edge_index, sd[prefix + "modlist.0.lin.weight"], sd[prefix + "modlist.0.bias"])
    x = graph_norm(x, sd[prefix + "modlist.1.weight"], sd[prefix + "modlist.1.bias"],
                   sd[prefix + "modlist.1.mean_scale"])
    return torch.relu(x) if act else x


def local_wl_forward(sd, x, edge1, pos, idx, ei2, act0: bool = True, act1: bool = True, node_feat=None):
    """LocalWLNet.forward (TwoWL/model/model.py:68-84), eval mode / dropout 0. ``sd`` maps the reference's
    state_dict keys to (possibly requires_grad) tensors. node_feat (use_node_feat=True, model.py:47-51,71):
    x = LayerNorm(no affine)(Linear(node_feat)) with the parameters lin1.1.0.{weight,bias}; otherwise the
    degree embedding + GraphNorm of model.py:53-55."""
    depth1 = len({k.split(".")[1] for k in sd if k.startswith("conv1s.")})
    depth2 = len({k.split(".")[1] for k in sd if k.startswith("conv2s.")})
    e = torch.as_tensor(_np(ei2)).to(torch.long)
    edge2 = torch.stack([e[0] ^ 1, e[1]])
    edge2_r = torch.stack([e[0], e[1] ^ 1])
    if node_feat is not None:
        u = node_feat.to(sd["lin1.1.0.weight"].dtype) @ sd["lin1.1.0.weight"].t() + sd["lin1.1.0.bias"]
        h = torch.nn.functional.layer_norm(u, (u.shape[1],), None, None, 1e-5)
    else:
        h = sd["emb.0.weight"].index_select(0, x)
        h = graph_norm(h, sd["emb.1.weight"], sd["emb.1.bias"], sd["emb.1.mean_scale"])
    for k in range(depth1):
        act = act0 if k < depth1 - 1 else act1
        h = _seq(sd, f"conv1s.{k}.", h, edge1, act)
    h = h.index_select(0, pos[:, 0]) * h.index_select(0, pos[:, 1])
    for k in range(depth2):
        h = _seq(sd, f"conv2s.{k}.", h, edge2, True) + _seq(sd, f"conv2s_r.{k}.", h, edge2_r, True)
    h = h.index_select(0, idx)
    h = h[0::2] * h[1::2]
    return h @ sd["pred.weight"].t() + sd["pred.bias"]


def init_state_dict(max_x: int, channels_1wl: int, channels_2wl: int, depth1: int, depth2: int,
                    seed: int = 0):
    """Random-init parameters with the reference's names, shapes and initialisers
    (nn.Embedding N(0,1); PyG Linear glorot-uniform; zeros bias; GraphNorm 1/0/1;
    nn.Linear kaiming-uniform) - SURVEY.md 3.4 state_dict list."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def gnorm(prefix, c):
        sd[prefix + "weight"] = torch.ones(c)
        sd[prefix + "bias"] = torch.zeros(c)
        sd[prefix + "mean_scale"] = torch.ones(c)

    def conv(prefix, cin, cout):
        a = (6.0 / (cin + cout)) ** 0.5
        sd[prefix + "modlist.0.bias"] = torch.zeros(cout)
        sd[prefix + "modlist.0.lin.weight"] = (torch.rand(cout, cin, generator=g) * 2 - 1) * a
        gnorm(prefix + "modlist.1.", cout)

    sd["emb.0.weight"] = torch.randn(max_x + 1, channels_1wl, generator=g)
    gnorm("emb.1.", channels_1wl)
    for k in range(depth1):
        conv(f"conv1s.{k}.", channels_1wl, channels_1wl if k < depth1 - 1 else channels_2wl)
    for k in range(depth2):
        conv(f"conv2s.{k}.", channels_2wl, channels_2wl)
    for k in range(depth2):
        conv(f"conv2s_r.{k}.", channels_2wl, channels_2wl)
    b = 1.0 / channels_2wl ** 0.5
    sd["pred.weight"] = (torch.rand(1, channels_2wl, generator=g) * 2 - 1) * b
    sd["pred.bias"] = (torch.rand(1, generator=g) * 2 - 1) * b
    return sd


def fwd_bwd(sd, x, edge1, pos, idx, ei2, y, act0=True, act1=True, node_feat=None):
    """One reference train step minus the optimiser (TwoWL/model/train.py:36-38):
    forward, BCE-with-logits, backward. Returns (logits, loss, {name: grad})."""
    leaf = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    pred = local_wl_forward(leaf, x, edge1, pos, idx, ei2, act0, act1, node_feat)
    loss = torch.nn.functional.binary_cross_entropy_with_logits(pred, y)
    loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaf.items()}
    return pred.detach(), loss.detach(), grads


# --------------------------------------------------------------------------------------
# synthetic graphs shared by tests and bench (seeded numpy; not part of the reference)
# --------------------------------------------------------------------------------------


def synthetic_split(num_nodes: int, und_edges: np.ndarray, seed: int = 0):
    """Canonical (r<c) deduplicated undirected edge list -> (pos_edge [2,E], pred_edge
    [2,P]) in the reference's doubled layout, one seeded uniform non-edge per positive
    (SURVEY.md 8(d) "Configs 2-5")."""
    rng = np.random.default_rng(seed)
    r, c = np.minimum(und_edges[0], und_edges[1]), np.maximum(und_edges[0], und_edges[1])
    keep = r != c
    key = np.unique(r[keep].astype(np.int64) * num_nodes + c[keep].astype(np.int64))
    rng.shuffle(key)
    m = key.size
    neg = np.empty(0, dtype=np.int64)
    while neg.size < m:
        cand = rng.integers(0, num_nodes, size=(2, int(1.2 * (m - neg.size)) + 16))
        cr, cc = np.minimum(cand[0], cand[1]), np.maximum(cand[0], cand[1])
        ck = (cr * num_nodes + cc)[cr != cc]
        ck = ck[~np.isin(ck, key)]
        neg = np.unique(np.concatenate([neg, ck]))
    rng.shuffle(neg)
    neg = neg[:m]
    pos_und = np.stack([key // num_nodes, key % num_nodes])
    neg_und = np.stack([neg // num_nodes, neg % num_nodes])
    return double(pos_und), double(neg_und)
