"""Drop-in for the reference's TwoWL/model/model.py: ``LocalWLNet`` and ``Seq`` with the same
constructor / forward signatures and the same state_dict keys (emb.0.weight, emb.1.*,
conv{1s,2s,2s_r}.K.modlist.0.{bias,lin.weight}, conv*.K.modlist.1.{weight,bias,mean_scale},
pred.{weight,bias}), plus ``GCNConv`` / ``GraphNorm`` replacing the torch_geometric 2.3.1 modules the
reference imports at model.py:3. Every tensor op in forward and backward is a libtwowl_b200.so kernel.
"""
from __future__ import annotations

import math

import torch
from torch import nn
from torch.nn.modules.dropout import Dropout

from TwoWL.utils import *  # noqa: F401,F403  (the reference does the same, model.py:5)
from TwoWL.utils import _struct_of
from twowl_b200 import functional as F2
from twowl_b200 import graph as G
from twowl_b200 import ops


def _seed() -> int:
    # drawn from torch's default CPU generator: reproducible under torch.manual_seed, no device sync
    return ops.next_seed()


class _Lin(nn.Module):
    """torch_geometric.nn.dense.linear.Linear(bias=False, weight_initializer='glorot')."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels))
        a = math.sqrt(6.0 / (in_channels + out_channels))
        nn.init.uniform_(self.weight, -a, a)

    def forward(self, x):
        return F2.linear(x, self.weight)


class GCNConv(nn.Module):
    """torch_geometric.nn.GCNConv(in_channels, out_channels) with its defaults (add_self_loops, normalize,
    bias, not cached) - reference call site model.py:37. ``edge_index`` is an int64 [2,E] tensor."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin = _Lin(in_channels, out_channels)
        self.bias = nn.Parameter(torch.zeros(out_channels))

    def forward(self, x, edge_index):
        g = G.node_graph(edge_index, x.shape[0])
        z = self.lin(x)
        return F2.gcn_aggregate(z, self.bias, g.dinv, g.ptr, g.col, g.plan, g.tptr, g.tcol, g.tplan, None, 0, 0, g.emask, g.temask)

    def forward_pairs(self, x, wedges, direction: int):
        """The pair-level call of model.py:77: direction 0 = conv(x, edge2), 1 = conv_r(x, edge2_r)."""
        z = self.lin(x)
        if isinstance(wedges, G.WedgeStruct):
            _, centre, dinv, selfw, _ = wedges.prepared()
            return F2.wedge_aggregate(z, self.bias, wedges.in_ptr, wedges.in_ids, wedges.in_plan, wedges.out_ptr,
                                      wedges.out_ids, wedges.out_plan, centre[direction], dinv[direction], selfw[direction], wedges.dst_e,
                                      wedges.blocked, wedges.E, wedges.n_node, direction)
        flip, row_flip = (1, 0) if direction == 0 else (0, 1)
        return F2.gcn_aggregate(z, self.bias, wedges.dinv[direction], wedges.ptr_b, wedges.col_b, wedges.plan_b,
                                wedges.ptr_a, wedges.col_a, wedges.plan_a, None, flip, row_flip)


class GraphNorm(nn.Module):
    """torch_geometric.nn.GraphNorm(in_channels, eps=1e-5), batch=None - reference model.py:38,54."""

    def __init__(self, in_channels, eps=1e-5):
        super().__init__()
        self.in_channels, self.eps = in_channels, eps
        self.weight = nn.Parameter(torch.ones(in_channels))
        self.bias = nn.Parameter(torch.zeros(in_channels))
        self.mean_scale = nn.Parameter(torch.ones(in_channels))

    def forward(self, x, batch=None):
        if batch is not None:
            raise NotImplementedError("the TwoWL path never passes a batch vector")
        return self.fused(x, 0.0, False, None)

    def fused(self, x, p_drop: float, relu: bool, addend):
        seed = _seed() if p_drop > 0.0 else 0
        out, _ = F2.graphnorm_act(x, self.weight, self.bias, self.mean_scale, addend, self.eps, p_drop, seed, relu)
        return out


def _norm_tail(mods, x, training: bool, addend=None):
    """[GraphNorm, Dropout, ReLU|Identity] of the reference's relu_conv / emb blocks as ONE fused kernel pair."""
    gn, dp, act = mods
    p = dp.p if (training and dp.p > 0.0) else 0.0
    return gn.fused(x, p, isinstance(act, nn.ReLU), addend)


class Seq(nn.Module):
    """model.py:87-96. When modlist is the reference's [GCNConv, GraphNorm, Dropout, ReLU|Identity] block the
    tail runs fused; any other modlist runs module by module exactly like the reference."""

    def __init__(self, modlist):
        super().__init__()
        self.modlist = nn.ModuleList(modlist)

    def _is_conv_block(self):
        m = self.modlist
        return (len(m) == 4 and isinstance(m[0], GCNConv) and isinstance(m[1], GraphNorm)
                and isinstance(m[2], Dropout) and isinstance(m[3], (nn.ReLU, nn.Identity)))

    def forward(self, *args, **kwargs):
        out = self.modlist[0](*args, **kwargs)
        if self._is_conv_block():
            return _norm_tail(list(self.modlist)[1:], out, self.training)
        for i in range(1, len(self.modlist)):
            out = self.modlist[i](out)
        return out

    def forward_pairs(self, x, wedges, direction: int, addend=None):
        out = self.modlist[0].forward_pairs(x, wedges, direction)
        return _norm_tail(list(self.modlist)[1:], out, self.training, addend)


class LocalWLNet(nn.Module):
    def __init__(self,
                 max_x,
                 use_node_feat,
                 node_feat,
                 channels_1wl=256,
                 channels_2wl=32,
                 depth1=1,
                 depth2=1,
                 dp_lin0=0.7,
                 dp_lin1=0.7,
                 dp_emb=0.5,
                 dp_1wl0=0.5,
                 dp_2wl=0.5,
                 dp_1wl1=0.5,
                 act0=True,
                 act1=True,
                 ):
        super().__init__()
        use_affine = False

        relu_lin = lambda a, b, dp, lnx, actx: nn.Sequential(  # noqa: E731
            nn.Linear(a, b),
            nn.LayerNorm(b, elementwise_affine=use_affine) if lnx else nn.Identity(),
            nn.Dropout(p=dp, inplace=True),
            nn.ReLU(inplace=True) if actx else nn.Identity())

        relu_conv = lambda insize, outsize, dp, act: Seq([  # noqa: E731
            GCNConv(insize, outsize),
            GraphNorm(outsize),
            Dropout(p=dp, inplace=True),
            nn.ReLU(inplace=True) if act else nn.Identity()
        ])

        self.max_x = max_x
        self.use_node_feat = use_node_feat
        self.node_feat = node_feat
        # "structured" | "explicit" | "auto": how the pair-level GCNConv consumes ei2 (see forward)
        self.pair_path = "auto"
        # one fused custom op per pair layer (tcgen05 GEMM + structured aggregation + GraphNorm statistics) when the
        # wedges are structured and the widths are supported; False = the op-by-op path
        self.fused_pair_layer = True
        # the LAST pair layer only feeds x[idx] (model.py:77-83): fuse it with the readout so its GraphNorm / ReLU run
        # on the selected rows only (needs fused_pair_layer and an idx)
        self.fused_readout = True
        # depth2 = 1 on the doubled pair layout: the pair init joins that op too (twowl::pair_init_layer_readout)
        self.fused_pair_init = True
        # a twowl_b200.rowshard.RowShard: forward / backward run on this rank's block of pair rows and node block (multi-GPU,
        # one step cut over the ranks). None = the whole pair table on this GPU.
        self.row_shard = None
        # regroup the pair rows internally by their higher-degree endpoint (graph.LocalityView): same terms, streaming gathers
        self.pair_locality = True

        if use_node_feat:
            self.lin1 = nn.Sequential(
                nn.Dropout(dp_lin0),
                relu_lin(node_feat.shape[-1], channels_1wl, dp_lin1, True, False)
            )
        else:
            self.emb = nn.Sequential(nn.Embedding(max_x + 1, channels_1wl),
                                     GraphNorm(channels_1wl),
                                     Dropout(p=dp_emb, inplace=True))

        self.conv1s = nn.ModuleList(
            [relu_conv(channels_1wl, channels_1wl, dp_1wl0, act0) for _ in range(depth1 - 1)] +
            [relu_conv(channels_1wl, channels_2wl, dp_1wl1, act1)])

        self.conv2s = nn.ModuleList(
            [relu_conv(channels_2wl, channels_2wl, dp_2wl, True) for _ in range(depth2)])
        self.conv2s_r = nn.ModuleList(
            [relu_conv(channels_2wl, channels_2wl, dp_2wl, True) for _ in range(depth2)])

        self.pred = nn.Linear(channels_2wl, 1)

    def _wedges(self, ei2, R: int, pt=None):
        """Pick the representation of ei2 the pair-level kernels read.
        structured: ei2 came from this package's get_ei2 / blockei2 / sample_block (or is a WedgeIndex) and
                    matches the pair table -> the factorised kernels, O(E + R) instead of O(T);
        explicit:   any int64 [2,T] tensor -> two CSRs over the wedges (built once per tensor, cached)."""
        struct = ei2.struct if isinstance(ei2, G.WedgeIndex) else _struct_of(ei2)
        ok = struct is not None and struct.R == R and struct.E % 2 == 0 and R % 2 == 0
        if ok and pt is not None:
            ok = G.wedges_match_table(struct, pt)      # the index was built for THIS pair table
        if self.pair_path == "structured" and not ok:
            raise RuntimeError("pair_path='structured' needs an ei2 produced by TwoWL.utils.get_ei2/sample_block "
                               "for this pair table")
        if ok and self.pair_path in ("auto", "structured"):
            return struct
        if isinstance(ei2, G.WedgeIndex):
            ei2 = ei2.materialize()
        return G.explicit_wedges(ei2, R)

    def forward(self, x, edge1, pos, idx=None, ei2=None, test=False):
        """model.py:68-84. x: int64 [N] degrees; edge1: int64 [2,E']; pos: int64 [R,2]; idx: int64 [2L];
        ei2: int64 [2,T'] (or a WedgeIndex). Returns fp32 [L,1] logits. ``test`` is unused, as in the
        reference. reverse(ei2) (model.py:69) is folded into the aggregation kernels."""
        if self.row_shard is not None and self.row_shard.world > 1:
            from twowl_b200 import rowshard
            return rowshard.forward(self, x, edge1, pos, idx, ei2)
        if self.use_node_feat:
            # model.py:47-51: lin1 = [Dropout(dp_lin0), [Linear, LayerNorm(no affine), Dropout(dp_lin1), Identity]] - the modules keep
            # the parameters (state_dict keys lin1.1.0.{weight,bias}); the arithmetic is one fused op on this package's kernels
            dp0, blk = self.lin1[0], self.lin1[1]
            lin, ln, dp1 = blk[0], blk[1], blk[2]
            p0 = dp0.p if (self.training and dp0.p > 0.0) else 0.0
            p1 = dp1.p if (self.training and dp1.p > 0.0) else 0.0
            x = F2.node_feat_input(self.node_feat, lin.weight, lin.bias, ln.eps, p0, _seed() if p0 > 0.0 else 0, p1,
                                   _seed() if p1 > 0.0 else 0)[0]
        else:
            emb, gn, dp = self.emb[0], self.emb[1], self.emb[2]
            x = F2.embedding(emb.weight, x)
            x = gn.fused(x, dp.p if (self.training and dp.p > 0.0) else 0.0, False, None)
        for conv1 in self.conv1s:
            x = conv1(x, edge1)

        pt = G.pair_table(pos, x.shape[0])             # validates pos against x's rows once per table (IndexError as model.py:75)
        if idx is not None:
            # x[idx] of model.py:78: ids in [-R, 0) wrap, anything else out of range is a device-side assertion (no host sync)
            idx = ops.index_guard(idx, pt.R, "idx")
        wedges = self._wedges(ei2, pt.R, pt) if len(self.conv2s) else None
        if self.pair_locality and isinstance(wedges, G.WedgeStruct) and pt.mated and idx is not None:
            lv = G.locality_view(wedges, pos)
            pt = G.pair_table(lv.pos, x.shape[0])
            blocked = wedges.blocked
            wedges = lv.struct if blocked is None else lv.struct.with_blocked(ops.gather_u8(blocked, lv.perm[:wedges.E]))
            idx = lv.newid[idx.reshape(-1)]
        if (self.fused_pair_init and len(self.conv2s) == 1 and self.fused_pair_layer and self.fused_readout and idx is not None
                and pt.mated and F2.pair_layer_supported(wedges, x.shape[1], self.conv2s[0], self.conv2s_r[0])):
            # depth2 = 1: pair init + the only pair layer + readout as ONE op - the gradient of the pair features never exists
            # at full height (one row per pair: the sum of its two directions, which is all the pair-init backward reads)
            return F2.pair_init_layer_readout_apply(x, pt, wedges, self.conv2s[0], self.conv2s_r[0], self.training, idx, self.pred)
        x = F2.pair_init(x, pt.src, pt.dst, pt.ptr_s, pt.ids_s, pt.plan_s, pt.ptr_d, pt.ids_d, pt.plan_d, pt.mated)
        if len(self.conv2s):
            last = len(self.conv2s) - 1
            for i in range(len(self.conv2s)):
                if self.fused_pair_layer and F2.pair_layer_supported(wedges, x.shape[1], self.conv2s[i], self.conv2s_r[i]):
                    if i == last and self.fused_readout and idx is not None:
                        return F2.pair_layer_readout_apply(x, wedges, self.conv2s[i], self.conv2s_r[i], self.training, idx,
                                                           self.pred)
                    x = F2.pair_layer_apply(x, wedges, self.conv2s[i], self.conv2s_r[i], self.training)
                else:
                    a = self.conv2s[i].forward_pairs(x, wedges, 0)
                    x = self.conv2s_r[i].forward_pairs(x, wedges, 1, addend=a)
        return F2.readout(x, idx, self.pred.weight, self.pred.bias)
