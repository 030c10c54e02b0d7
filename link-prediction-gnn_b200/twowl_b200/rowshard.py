"""Row-block sharding of the pair table over several GPUs: ONE TwoWL step cut over the ranks.

Rank g owns the contiguous block [lo, hi) of pair rows (even boundaries: rows 2k / 2k+1 are the two directions of one pair,
utils.py:81-90, and stay together), i.e. its rows of every [R, C] activation of model.py:75-83, the observed edges and the
target links that fall into the block. The node-level part of the model (model.py:71-73, [N, C] tensors) is replicated.
On the factorised wedge path a pair row only talks to the other rows through per-NODE sums, so the exchange per step is

    forward   all_reduce  SH_f, SH_r   [2, N, C] fp32      in-list sums of the pair layer (partial over the rank's edges)
              all_reduce  moments      [2, 2C]   fp64      GraphNorm column (sum, sum of squares) of both branches
              all_reduce  logits       [L]       fp32      every rank ends with the full [L, 1] output of model.py:83
    backward  all_reduce  colsums      [6, C]    fp64      GraphNorm-backward / readout column sums over the selected rows
                                                           ([4, C] over all rows for a layer that is not the last)
              all_reduce  dS_f, dS_r   [2, N, C] fp32      gradient of the per-node sums
    (+ the caller's ONE all-reduce of the flat parameter gradients, dist.allreduce_grads)

instead of the all-gather / reduce-scatter of [R, C] an explicit wedge index would need. Every rank computes the same numbers
as the single-GPU path up to the summation order of those five reductions. Gradients that are complete on every rank after an
all-reduce (GraphNorm / readout parameters) are kept on rank 0 only, so that the caller's gradient SUM is exact.

Covers any depth2 >= 1 on the structured wedge path with the doubled pair layout and widths the tensor-core kernels take
(32 / 64 / 128); every non-last layer adds its own SH / moments / GraphNorm-backward column sums / dS exchanges. Dropout masks are keyed by the LOCAL row id: statistically the
same as the single-GPU run, not the same bits (parity tests run with dropout 0 / eval mode).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import torch
import torch.distributed as dist

from . import graph as G
from . import ops


def block_of(R: int, rank: int, world: int) -> Tuple[int, int]:
    """[lo, hi) of rank's block: equal numbers of undirected pairs (the factorised path costs the same for every row)."""
    if R % 2:
        raise ValueError("the pair table must have an even number of rows")
    pairs = R // 2
    return 2 * (pairs * rank // world), 2 * (pairs * (rank + 1) // world)


class RowShard:
    """Assign to ``LocalWLNet.row_shard`` to run forward / backward on this rank's block of pair rows."""

    def __init__(self, group=None, rank: Optional[int] = None, world: Optional[int] = None):
        self.group = group
        if rank is None:
            rank = dist.get_rank(group) if dist.is_initialized() else 0
            world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank, self.world = int(rank), int(world)

    def all_reduce(self, t: torch.Tensor) -> torch.Tensor:
        if self.world > 1:
            dist.all_reduce(t, group=self.group)
        return t


@dataclass
class _Local:
    lo: int
    hi: int
    E_loc: int             # observed edges inside the block (rows [lo, min(hi, E)))
    src: torch.Tensor      # int32 [Rl]
    dst: torch.Tensor
    in_ptr: torch.Tensor   # the block's observed edges grouped by target node (local ids)
    in_ids: torch.Tensor
    in_plan: torch.Tensor
    out_ptr: torch.Tensor  # the block's pair rows grouped by source node (local ids) = pair_init's backward CSR
    out_ids: torch.Tensor
    out_plan: torch.Tensor


def _local(struct: G.WedgeStruct, pt: G.PairTable, lo: int, hi: int) -> _Local:
    def build():
        n = struct.n_node
        hiE = max(lo, min(hi, struct.E))
        in_ptr, in_ids = ops.csr_build(struct.dst_e[lo:hiE].to(torch.int64), n)
        out_ptr, out_ids = ops.csr_build(struct.src[lo:hi].to(torch.int64), n)
        return (pt, _Local(lo, hi, hiE - lo, pt.src[lo:hi].contiguous(), pt.dst[lo:hi].contiguous(), in_ptr, in_ids,
                           ops.seg_plan(in_ptr, n, hiE - lo), out_ptr, out_ids, ops.seg_plan(out_ptr, n, hi - lo)))
    # the entry holds struct.src (key tensor) and pt (value): neither address can be recycled while the block is cached
    return G._cache.get(struct.src, ("rowshard", lo, hi) + G._Cache.key(pt.src), build)[1]


def _layer_forward(shard, loc, rows, blocked_l, R_total, n_node, eps, H, pf, pr):
    """Both directions' pre-GraphNorm outputs of one pair layer on this block + the GLOBAL statistics:
    -> (O_f, O_r, stats_f, stats_r, SH [2,N,C])."""
    centre, dinv, selfw, _ = rows
    (wf, bf, gmf), (wr, br, gmr) = pf, pr
    SHr, SHf = ops.seg_reduce(loc.in_ptr, loc.in_ids, n_node, H, plan=loc.in_plan, src_scale=dinv[1], skip_mask=blocked_l, dual=True,
                              src_scale2=dinv[0])
    SH = torch.stack((SHf, SHr))
    del SHf, SHr
    shard.all_reduce(SH)
    Sf, Sr = ops.linear_fwd(SH[0], wf), ops.linear_fwd(SH[1], wr)
    if ops.PAIR_CONV_DUAL and ops.pair_conv_dual_supported(H.shape[1], wf.shape[0]):
        Of, Or, mf, mr = ops.pair_conv_dual(H, wf, wr, selfw[0], selfw[1], (Sf, centre[0], dinv[0]), (Sr, centre[1], dinv[1]), bf, br,
                                            gmf, gmr, eps, want_moments=True)
        Os, moms = [Of, Or], [mf, mr]
    else:
        Os, moms = [], []
        for d, (w, b, gm, S) in enumerate(((wf, bf, gmf, Sf), (wr, br, gmr, Sr))):
            O, mom = ops.pair_conv([H], [w], [0], row_scale=[selfw[d]], gathers=[(S, centre[d], dinv[d])], bias=b,
                                   stats_mean_scale=gm, eps=eps, want_moments=True)
            Os.append(O)
            moms.append(mom)
    mom = shard.all_reduce(torch.stack(moms))
    sf = ops.graphnorm_stats_from_moments(mom[0], R_total, gmf, eps)
    sr = ops.graphnorm_stats_from_moments(mom[1], R_total, gmr, eps)
    return Os[0], Os[1], sf, sr, SH


def _layer_backward(shard, loc, rows, n_node, H, wf, wr, SH, dOf, dOr, dWf, dWr):
    """From the gradients of the two pre-GraphNorm outputs (and the (selfw*dO)^T H parts of the weight gradients) to
    (dH, dW_f, dW_r) on this block: the dS exchange and the input-gradient pass."""
    _, dinv, selfw, bnode = rows
    dOs = (dOf, dOr)
    dS = torch.stack([ops.seg_reduce(loc.out_ptr, loc.out_ids, n_node, dOs[d], plan=loc.out_plan, flip=d, src_scale=dinv[d])
                      for d in range(2)])
    # dW_d = (selfw_d * dO_d)^T H + dS_d^T SH_d: with THIS rank's partial dS the sum over ranks is the full product
    dWf = dWf + ops.linear_bwd_weight(dS[0], SH[0])
    dWr = dWr + ops.linear_bwd_weight(dS[1], SH[1])
    shard.all_reduce(dS)
    dSW = [ops.linear_bwd_input(dS[0], wf), ops.linear_bwd_input(dS[1], wr)]
    dh = ops.pair_conv([dOf, dOr], [wf, wr], [1, 1], row_scale=[selfw[0], selfw[1]],
                       gathers=[(dSW[0], bnode[0], dinv[0]), (dSW[1], bnode[1], dinv[1])])
    return dh, dWf, dWr


def _param_grads(shard, C, dpf, dpr, extra=()):
    """dparams = [d gn.weight | d gn.bias | d gn.mean_scale | d conv.bias], made from rank-summed column sums, i.e. complete on
    every rank: kept on rank 0 so that the caller's SUM over ranks is exact."""
    if shard.rank != 0:
        for t in (dpf, dpr) + tuple(extra):
            t.zero_()
    return (dpf[3 * C:], dpf[:C], dpf[C:2 * C], dpf[2 * C:3 * C]), (dpr[3 * C:], dpr[:C], dpr[C:2 * C], dpr[2 * C:3 * C])


class _ShardedPairInit(torch.autograd.Function):
    """model.py:75 on one row block: x [N, C] (replicated) -> H [Rl, C]; backward = this block's part of dx."""

    @staticmethod
    def forward(ctx, x, loc, n_node):
        x = x.contiguous()
        ctx.save_for_backward(x)
        ctx.meta = (loc, n_node)
        return ops.pair_init_fwd(x, loc.src, loc.dst)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        loc, n_node = ctx.meta
        dx = ops.seg_reduce(loc.out_ptr, loc.out_ids, n_node, g.contiguous(), plan=loc.out_plan, X2=x, mul_idx=loc.dst, pair_sum=True)
        return dx, None, None


class _ShardedPairLayer(torch.autograd.Function):
    """A conv2s[i] / conv2s_r[i] layer that is NOT the last one (model.py:77) on one row block: H [Rl, C] -> H' [Rl, C]."""

    @staticmethod
    def forward(ctx, H, wf, bf, gwf, gbf, gmf, wr, br, gwr, gbr, gmr, shard, loc, rows, blocked_l, R_total, n_node, eps, p_drop,
                seed_f, seed_r):
        H = H.contiguous()
        Of, Or, sf, sr, SH = _layer_forward(shard, loc, rows, blocked_l, R_total, n_node, eps, H, (wf, bf, gmf), (wr, br, gmr))
        hn = ops.graphnorm_apply2(Of, Or, sf, sr, (gwf, gbf, gmf), (gwr, gbr, gmr), p_drop, seed_f, seed_r, True)
        ctx.save_for_backward(H, wf, gwf, gbf, gmf, wr, gwr, gbr, gmr, Of, Or, sf, sr, SH, *rows)
        ctx.meta = (shard, loc, R_total, n_node, p_drop, seed_f, seed_r)
        return hn

    @staticmethod
    def backward(ctx, g):
        H, wf, gwf, gbf, gmf, wr, gwr, gbr, gmr, Of, Or, sf, sr, SH, *rows = ctx.saved_tensors
        shard, loc, R_total, n_node, p_drop, seed_f, seed_r = ctx.meta
        C = wf.shape[0]
        dOf, dOr, dpf, dpr = ops.graphnorm_bwd2_sharded(Of, Or, g.contiguous(), sf, sr, (gwf, gbf, gmf), (gwr, gbr, gmr), p_drop, seed_f,
                                                        seed_r, True, R_total, shard.all_reduce)
        dWf, dWr = ops.pair_dw(dOf, dOr, rows[2][0], rows[2][1], H)
        dh, dWf, dWr = _layer_backward(shard, loc, rows, n_node, H, wf, wr, SH, dOf, dOr, dWf, dWr)
        gf, gr = _param_grads(shard, C, dpf, dpr)
        return (dh, dWf) + gf + (dWr,) + gr + (None,) * 10


class _ShardedLastLayer(torch.autograd.Function):
    """The last conv2s / conv2s_r layer (model.py:77) + readout (model.py:78-83) on one row block: H [Rl, C] -> logits of the
    target links inside the block."""

    @staticmethod
    def forward(ctx, H, wf, bf, gwf, gbf, gmf, wr, br, gwr, gbr, gmr, pw, pb, shard, loc, rows, idx_l, blocked_l, R_total, n_node,
                eps, p_drop, seed_f, seed_r):
        H = H.contiguous()
        Of, Or, sf, sr, SH = _layer_forward(shard, loc, rows, blocked_l, R_total, n_node, eps, H, (wf, bf, gmf), (wr, br, gmr))
        if idx_l.numel():
            pred = ops.gn2_readout_fwd(Of, Or, sf, sr, (gwf, gbf, gmf), (gwr, gbr, gmr), p_drop, seed_f, seed_r, True, idx_l,
                                       pw.contiguous(), pb)
        else:
            pred = H.new_empty((0, 1))
        ctx.save_for_backward(H, wf, gwf, gbf, gmf, wr, gwr, gbr, gmr, pw, idx_l, Of, Or, sf, sr, SH, *rows)
        ctx.meta = (shard, loc, R_total, n_node, p_drop, seed_f, seed_r)
        return pred

    @staticmethod
    def backward(ctx, g):
        H, wf, gwf, gbf, gmf, wr, gwr, gbr, gmr, pw, idx_l, Of, Or, sf, sr, SH, *rows = ctx.saved_tensors
        shard, loc, R_total, n_node, p_drop, seed_f, seed_r = ctx.meta
        C = wf.shape[0]
        pf, pr = (gwf, gbf, gmf), (gwr, gbr, gmr)
        Gp, head, nxt, colsums = ops.gn2_readout_bwd_rows(Of, Or, sf, sr, pf, pr, p_drop, seed_f, seed_r, True, idx_l, pw.contiguous(),
                                                          g.reshape(-1))
        shard.all_reduce(colsums)
        consts, dpf, dpr, dpw, dpb = ops.gn2_readout_bwd_finish(colsums, R_total, sf, sr, pf, pr)
        dOf, dOr, dWf, dWr = ops.pair_dw_gn(Of, Or, consts, Gp, head, nxt, p_drop, seed_f, seed_r, True, rows[2][0], rows[2][1], H)
        dh, dWf, dWr = _layer_backward(shard, loc, rows, n_node, H, wf, wr, SH, dOf, dOr, dWf, dWr)
        gf, gr = _param_grads(shard, C, dpf, dpr, extra=(dpw, dpb))
        return (dh, dWf) + gf + (dWr,) + gr + (dpw.reshape(pw.shape), dpb) + (None,) * 11


class _ScatterLogits(torch.autograd.Function):
    """Every rank's [L_loc, 1] logits -> the full [L, 1] tensor of model.py:83 on every rank (one all-reduce of L floats);
    backward hands each rank the gradient of its own links."""

    @staticmethod
    def forward(ctx, pred_l, links_l, L, shard):
        full = pred_l.new_zeros((L, 1))
        full[links_l] = pred_l
        shard.all_reduce(full)
        ctx.save_for_backward(links_l)
        return full

    @staticmethod
    def backward(ctx, g):
        (links_l,) = ctx.saved_tensors
        return g[links_l].contiguous(), None, None, None


def supported(model, wedges, C: int) -> Optional[str]:
    """None if the sharded path covers this model / input, else the reason."""
    from . import functional as F2
    if len(model.conv2s) < 1:
        return "row sharding needs at least one pair layer"
    if not isinstance(wedges, G.WedgeStruct):
        return "row sharding needs the structured wedge index (TwoWL.utils.get_ei2 / sample_block)"
    if not all(F2.pair_layer_supported(wedges, C, f, r) for f, r in zip(model.conv2s, model.conv2s_r)) or not ops.pair_dw_supported(C):
        return f"pair width {C} is not covered by the tensor-core pair kernels"
    return None


def forward_pairs(model, x, pos, idx, ei2):
    """The pair-level part of LocalWLNet.forward (model.py:75-83) on this rank's row block; returns the full [L, 1] logits."""
    shard: RowShard = model.row_shard
    if idx is None:
        raise RuntimeError("row-sharded forward needs idx (the target links)")
    pt = G.pair_table(pos, x.shape[0])
    wedges = model._wedges(ei2, pt.R, pt)
    why = supported(model, wedges, x.shape[1])
    if why is None and not pt.mated:
        why = "row sharding needs the doubled pair layout (rows 2k / 2k+1 = (u,v) / (v,u))"
    if why is not None:
        raise NotImplementedError(why)
    idx = ops.index_guard(idx, pt.R, "idx")      # x[idx] of model.py:78: negative ids wrap, anything else out of range asserts
    if getattr(model, "pair_locality", False):
        # the blocks are cut from the pair table regrouped by hub (graph.LocalityView): same rows, streaming per-node gathers
        lv = G.locality_view(wedges, pos)
        pt = G.pair_table(lv.pos, x.shape[0])
        blocked = wedges.blocked
        wedges = lv.struct if blocked is None else lv.struct.with_blocked(ops.gather_u8(blocked, lv.perm[:wedges.E]))
        idx = lv.newid[idx]
    lo, hi = block_of(pt.R, shard.rank, shard.world)
    loc = _local(wedges, pt, lo, hi)
    _, centre, dinv, selfw, bnode = wedges.prepared()
    rows = tuple(t[:, lo:hi].contiguous() for t in (centre, dinv, selfw, bnode))
    blocked_l = wedges.blocked[lo:lo + loc.E_loc] if wedges.blocked is not None else None
    L = idx.numel() // 2
    inb = ((idx >= lo) & (idx < hi)).reshape(L, 2)
    links_l = torch.nonzero(inb[:, 0]).reshape(-1)          # one host read: the block's target links
    if bool((inb[:, 0] ^ inb[:, 1]).any().item()):
        raise RuntimeError("row-sharded readout: idx[2l] and idx[2l+1] must be the two rows of one pair (double(.., for_index=True))")
    idx_l = (idx.reshape(L, 2)[links_l] - lo).reshape(-1).contiguous()
    H = _ShardedPairInit.apply(x, loc, wedges.n_node)
    last = len(model.conv2s) - 1
    for i, (seq_f, seq_r) in enumerate(zip(model.conv2s, model.conv2s_r)):
        cf, gf, dpf = seq_f.modlist[0], seq_f.modlist[1], seq_f.modlist[2]
        cr, gr = seq_r.modlist[0], seq_r.modlist[1]
        p = dpf.p if (model.training and dpf.p > 0.0) else 0.0
        seeds = [ops.next_seed() for _ in range(2)] if p > 0.0 else [0, 0]
        par = (cf.lin.weight, cf.bias, gf.weight, gf.bias, gf.mean_scale, cr.lin.weight, cr.bias, gr.weight, gr.bias, gr.mean_scale)
        if i < last:
            H = _ShardedPairLayer.apply(H, *par, shard, loc, rows, blocked_l, pt.R, wedges.n_node, gf.eps, p, seeds[0], seeds[1])
        else:
            pred_l = _ShardedLastLayer.apply(H, *par, model.pred.weight, model.pred.bias, shard, loc, rows, idx_l, blocked_l, pt.R,
                                             wedges.n_node, gf.eps, p, seeds[0], seeds[1])
    return _ScatterLogits.apply(pred_l, links_l, L, shard)
