"""Drop-in for the reference's TwoWL/utils.py (same names, positional signatures, return layouts and
dtypes - reference lines cited per function), running on libtwowl_b200.so's sm_100a kernels.

Differences a caller can see: tensors must be CUDA tensors (there is no CPU path), and the wedge
tensors returned by get_ei2 / blockei2 / sample_block carry a hidden ``_twowl_wedge`` attribute with
the factored form of the join, which LocalWLNet.forward uses to skip the O(T) index entirely.
"""
from __future__ import annotations

import math

import torch
from torch import Tensor

from twowl_b200 import ops
from twowl_b200.graph import WedgeIndex, WedgeStruct, build_wedge_struct

__all__ = ["degree", "set_mul", "check_in_set", "get_ei2", "blockei2", "idx2mask", "sample_block", "reverse",
           "double", "random_split_edges", "get_ei2_implicit", "get_ei2_shard", "WedgeIndex", "torch", "Tensor", "math"]


def _tag(t: Tensor, struct: WedgeStruct) -> Tensor:
    t._twowl_wedge = (struct, t._version)
    return t


def _struct_of(t) -> "WedgeStruct | None":
    tag = getattr(t, "_twowl_wedge", None)
    if tag is None or tag[1] != t._version:   # tensor was modified in place after we built it
        return None
    return tag[0]


def degree(ei: Tensor, num_node: int):
    """utils.py:8-10 - scatter_add(ones, ei[1], dim_size=num_node): histogram of the target row."""
    return ops.degree(ei[1], num_node)


def set_mul(a: Tensor, b: Tensor):
    """utils.py:13-19 - Cartesian product, a-major, int64 [p*q, 2]."""
    return ops.set_mul(a, b)


def check_in_set(target, set):
    """utils.py:22-33 - out[t] = number of entries of ``set`` equal to target[t] (int64)."""
    return ops.check_in_set(target, set)


def _materialize(struct: WedgeStruct) -> Tensor:
    n = struct.n_node
    off = ops.ei2_offsets(struct.in_ptr, struct.out_ptr, n)
    T = int(off[-1].item())  # data-dependent size: the reference's torch.cat synchronises here too
    full = ops.ei2_fill(struct.in_ptr, struct.in_ids, struct.out_ptr, struct.out_ids, off, n, 0, T).t()
    if struct.blocked is not None:
        full = ops.select_columns(full, struct.blocked, 1)
    return full


def _materialize_unblocked(struct: WedgeStruct) -> Tensor:
    """blockei2 of an index get_ei2 made = the join over the in-lists with the blocked edges taken out (same order: centre
    ascending, a ascending, b ascending), written once as the contiguous [2,T'] tensor utils.py:50 returns - instead of reading
    the [2,T] tensor twice to compact it."""
    n = struct.n_node
    live = ops.gather_u8(struct.blocked, struct.in_ids) == 0           # per in-list ENTRY, in list order
    in_ids = struct.in_ids[live]
    cnt = struct.prepared()[0][:n].to(torch.int64)                      # live in-edges per node
    in_ptr = torch.zeros(n + 1, dtype=torch.int64, device=cnt.device)
    torch.cumsum(cnt, 0, out=in_ptr[1:])
    off = ops.ei2_offsets(in_ptr, struct.out_ptr, n)
    T = int(off[-1].item())
    return ops.ei2_fill_rows(in_ptr, in_ids, struct.out_ptr, struct.out_ids, off, n, T)


def get_ei2(n_node: int, pos_edge, pred_edge):
    """utils.py:36-45 - the wedge join: for centre node i ascending, every observed edge id a with
    pos_edge[1][a]==i (ascending) x every pair id b with cat(pos_edge,pred_edge)[0][b]==i (ascending).
    Returns int64 [2,T] as the same transposed view of a [T,2] buffer the reference returns."""
    struct = build_wedge_struct(int(n_node), pos_edge, pred_edge)
    return _tag(_materialize(struct), struct)


def get_ei2_shard(n_node: int, pos_edge, pred_edge, rank: int, world: int):
    """Rank `rank`'s contiguous slice of get_ei2's columns when `world` ranks build the index together (extension):
    every rank holds the small int edge lists and fills only wedges [T*rank/world, T*(rank+1)/world) - equal work whatever the
    degree skew, no exchange - and the slices concatenated in rank order ARE get_ei2(n_node, pos_edge, pred_edge), because the
    layout is one global prefix sum. Returns (int64 [2, T_rank] in the reference's transposed layout, first column index, T)."""
    struct = build_wedge_struct(int(n_node), pos_edge, pred_edge)
    n = struct.n_node
    off = ops.ei2_offsets(struct.in_ptr, struct.out_ptr, n)
    T = int(off[-1].item())
    t0, t1 = T * int(rank) // int(world), T * (int(rank) + 1) // int(world)
    part = ops.ei2_fill(struct.in_ptr, struct.in_ids, struct.out_ptr, struct.out_ids, off, n, t0, t1).t()
    return part, t0, T


def get_ei2_implicit(n_node: int, pos_edge, pred_edge) -> WedgeIndex:
    """get_ei2 without materialising [2,T] (extension; T = sum deg^2 does not fit at R-MAT scale)."""
    return WedgeIndex(build_wedge_struct(int(n_node), pos_edge, pred_edge))


def blockei2(ei2, blocked_idx):
    """utils.py:48-50 - keep the wedges whose source edge id ei2[0] is not in blocked_idx (order kept)."""
    struct = ei2.struct if isinstance(ei2, WedgeIndex) else _struct_of(ei2)
    blocked_idx = blocked_idx.reshape(-1)
    if struct is not None:
        mask = ops.mask_from_idx(blocked_idx, struct.E)  # ids outside [0,E) can match no source edge
        new_struct = struct.with_blocked(mask)
        if isinstance(ei2, WedgeIndex):
            return WedgeIndex(new_struct)
        if struct.E % 2 == 0 and struct.R % 2 == 0:
            return _tag(_materialize_unblocked(new_struct), new_struct)
        return _tag(ops.select_columns(ei2, mask, 1), new_struct)
    num = int(blocked_idx.max().item()) + 1 if blocked_idx.numel() else 1
    mask = ops.mask_from_idx(blocked_idx, max(num, 1))
    return ops.select_columns(ei2, mask, 1)


def idx2mask(num: int, idx):
    """utils.py:53-57 - bool mask of length num, True at idx."""
    return ops.mask_from_idx(idx, num).view(torch.bool)


def sample_block(sample_idx, size, ei, ei2=None):
    """utils.py:61-68 - remove the sampled edge ids from ei, recount the degree BY SOURCE (the sparse
    row-sum of utils.py:66-67), filter ei2 with blockei2. Returns (ei_new, x_new, ei2_new)."""
    mask = ops.mask_from_idx(sample_idx, ei.shape[1])
    ei_new = ops.select_columns(ei, mask, 0)
    ei_new._twowl_edges = (ei, mask, ei_new._version)   # lets GCNConv reuse the cached CSRs of `ei` (graph.node_graph)
    x_new = ops.degree(ei_new[0], int(size))
    ei2_new = blockei2(ei2, sample_idx) if ei2 is not None else None
    return ei_new, x_new, ei2_new


def reverse(edge_index):
    """utils.py:71-78 - edge = [a^1; b], edge_r = [a; b^1] (+1 for even ids, -1 for odd)."""
    if isinstance(edge_index, WedgeIndex):
        raise TypeError("reverse() needs a materialised [2,T] tensor; LocalWLNet.forward folds it into the kernels")
    return ops.reverse(edge_index)


def double(x, for_index=False):
    """utils.py:81-90 - pair k -> directed ids 2k=(r,c), 2k+1=(c,r); for_index: k -> 2k, 2k+1."""
    if not for_index:
        return ops.double_edges(x)
    return ops.double_index(x)


def random_split_edges(data, val_ratio: float = 0.05, test_ratio: float = 0.1):
    """utils.py:93-147 - keep row<col, random train/val/test split of the positives, and n_v+n_t negatives
    drawn uniformly from the non-edges of the upper triangle. Same attribute names and shapes as the
    reference; the dense N x N mask of utils.py:130 (1 TB at 1M nodes) is replaced by the hash-set sampler kernel
    (csrc/sampler.cu), and ``train_neg_adj_mask`` (never read by the TwoWL path) is omitted."""
    num_nodes = int(data.num_nodes)
    row, col = data.edge_index
    edge_attr = getattr(data, "edge_attr", None)
    data.edge_index = data.edge_attr = None
    mask = row < col
    row, col = row[mask], col[mask]
    if edge_attr is not None:
        edge_attr = edge_attr[mask]
    n_v = int(math.floor(val_ratio * row.size(0)))
    n_t = int(math.floor(test_ratio * row.size(0)))
    perm = torch.randperm(row.size(0), device=row.device)
    row, col = row[perm], col[perm]
    if edge_attr is not None:
        edge_attr = edge_attr[perm]
    data.val_pos_edge_index = torch.stack([row[:n_v], col[:n_v]], dim=0)
    data.test_pos_edge_index = torch.stack([row[n_v:n_v + n_t], col[n_v:n_v + n_t]], dim=0)
    data.train_pos_edge_index = torch.stack([row[n_v + n_t:], col[n_v + n_t:]], dim=0)
    if edge_attr is not None:
        data.val_pos_edge_attr = edge_attr[:n_v]
        data.test_pos_edge_attr = edge_attr[n_v:n_v + n_t]

    # n_v + n_t distinct uniform non-edges of the upper triangle (utils.py:127-139): twowl_nonedge_sample, a seeded hash-set
    # sampler, instead of the dense N x N mask; a graph with fewer non-edges than that returns what exists, like the reference
    neg_row, neg_col = ops.sample_non_edges(row.to(torch.int64), col.to(torch.int64), num_nodes, n_v + n_t, undirected=True)
    data.val_neg_edge_index = torch.stack([neg_row[:n_v], neg_col[:n_v]], dim=0)
    data.test_neg_edge_index = torch.stack([neg_row[n_v:n_v + n_t], neg_col[n_v:n_v + n_t]], dim=0)
    return data
