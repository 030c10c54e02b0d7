"""Host-side plumbing of the multi-GPU path (one process per GPU, torch.distributed; NCCL on GPUs, gloo in the CPU tests).

The TwoWL step shards over TARGET LINKS: every rank holds the graph (the reference's `dataset` object, read-only), takes
its own disjoint slice of the global batch of target links, runs sample_block + forward + backward for that slice, and
the parameter gradients are summed with ONE all-reduce of a flat ~10^4-10^5 float buffer (train.py:18-38 per rank).
There is no data-path collective: target links are independent units of work given the graph. Nothing in this module
touches the C ABI, so it is importable (and tested) without a GPU.
"""
from __future__ import annotations

from typing import Iterable, List, Sequence, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_batch(n_items: int, batch: int, step: int, seed: int = 0, replicate: bool = False) -> torch.Tensor:
    """This rank's slice of step `step`'s global batch: `batch` ids per rank, drawn without replacement from ONE
    seeded global permutation of range(n_items), so the slices of different ranks are disjoint and their union is
    a uniform sample of world*batch ids (train.py:18-23 draws one such permutation per epoch). CPU int64 tensor.
    replicate=True: every rank gets the SAME `batch` ids (row-sharded steps: the ranks cut one batch's pair rows)."""
    rank, ws = (0, 1) if replicate else world()
    if batch * ws > n_items:
        raise ValueError(f"global batch {batch}*{ws} exceeds the {n_items} available ids")
    g = torch.Generator().manual_seed(seed * 1_000_003 + step)
    perm = torch.randperm(n_items, generator=g)
    return perm[rank * batch:(rank + 1) * batch].clone()


def allreduce_grads(params: Iterable[torch.nn.Parameter], average: bool = False) -> int:
    """Sum (or average) the .grad of every parameter over the ranks with one all-reduce of a flat buffer, written back in
    place. Parameters without a gradient take part as zeros so every rank reduces the same layout. Returns the number of
    floats reduced."""
    params = [p for p in params if p.requires_grad]
    if not params:
        return 0
    _, ws = world()
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params])
    if ws > 1:
        dist.all_reduce(flat)
        if average:
            flat /= ws
    off = 0
    for p in params:
        n = p.numel()
        g = flat[off:off + n].view_as(p)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += n
    return int(flat.numel())


def max_over_ranks(value: float, device=None) -> float:
    """Slowest rank's value - multi-GPU times are reported as the max over ranks."""
    _, ws = world()
    if ws == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def row_blocks(weights: Sequence[int], parts: int) -> List[Tuple[int, int]]:
    """Contiguous blocks [lo, hi) of pair rows, one per rank, with EVEN boundaries (rows 2k / 2k+1 are the two directions
    of one pair and must stay together, utils.py:81-90) and balanced by the given per-row weights (e.g. wedge counts:
    degree skew makes equal row counts unbalanced). Used to shard full-graph inference (train.py:54-60: every prediction
    pair is a target) over ranks; the concatenation of the per-rank logits in rank order is the single-GPU result."""
    w = torch.as_tensor(weights, dtype=torch.float64)
    R = int(w.numel())
    if R % 2:
        raise ValueError("the pair table must have an even number of rows")
    pw = w.reshape(-1, 2).sum(1).cumsum(0)              # per undirected pair
    total = float(pw[-1]) if R else 0.0
    cuts = [0]
    for k in range(1, parts):
        target = total * k / parts
        j = int(torch.searchsorted(pw, torch.tensor([target], dtype=torch.float64)).item())   # first pair with cum >= target
        if j < pw.numel():
            before = float(pw[j - 1]) if j > 0 else 0.0
            if float(pw[j]) - target < target - before:   # the boundary after pair j is the nearer one
                j += 1
        cuts.append(max(cuts[-1], min(2 * j, R)))
    cuts.append(R)
    return [(cuts[i], cuts[i + 1]) for i in range(parts)]


def merge_column_stats(n: torch.Tensor, mean: torch.Tensor, m2: torch.Tensor) -> Tuple[float, torch.Tensor, torch.Tensor]:
    """Chan's parallel merge of per-rank column statistics (count n[g], mean[g, C], M2[g, C] = sum of squared deviations)
    into the global (count, mean, M2) - what a row-sharded GraphNorm all-gathers (2C+1 numbers per rank) instead of
    reducing raw sums (hard part 4). Fixed rank order: deterministic."""
    N = float(n[0])
    mu = mean[0].double().clone()
    M2 = m2[0].double().clone()
    for g in range(1, n.numel()):
        nb = float(n[g])
        if nb == 0:
            continue
        delta = mean[g].double() - mu
        tot = N + nb
        mu = mu + delta * (nb / tot)
        M2 = M2 + m2[g].double() + delta * delta * (N * nb / tot)
        N = tot
    return N, mu, M2
