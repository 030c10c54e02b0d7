"""Input formats and generators either side of the hot path: seeded on-device R-MAT / uniform graphs in
the reference's edge-list conventions, and an edge-list reader for CSV (datasets.py:154-168's format), .npy and raw binary.

The generators are plain torch on the device they are given (input generation, not the measured path); the negative draw is
the package's hash-set sampler kernel (CUDA only, like every operator here). The outputs feed
``TwoWL.utils.double`` / ``get_ei2`` / ``operators.datasets.graph_from_split`` exactly like the reference's CSV does.
"""
from __future__ import annotations

import os
from typing import Dict, Optional, Tuple

import numpy as np
import torch


def rmat_edges(scale: int, samples: int, abcd: Tuple[float, float, float, float], seed: int, device) -> Tuple[torch.Tensor, torch.Tensor, int]:
    """`samples` seeded R-MAT edge samples on 2**scale nodes (quadrant probabilities a, b, c, d) with a random vertex
    relabelling (so that node id carries no degree information) -> (src, dst, n). Duplicates and self loops still inside."""
    g = torch.Generator(device=device).manual_seed(seed)
    a, b, c, _ = abcd
    src = torch.zeros(samples, dtype=torch.int64, device=device)
    dst = torch.zeros(samples, dtype=torch.int64, device=device)
    for _ in range(scale):
        r = torch.rand(samples, generator=g, device=device)
        src = src * 2 + (r >= a + b).to(torch.int64)
        dst = dst * 2 + (((r >= a) & (r < a + b)) | (r >= a + b + c)).to(torch.int64)
    n = 1 << scale
    perm = torch.randperm(n, generator=g, device=device)
    return perm[src], perm[dst], n


def canonical_undirected(src: torch.Tensor, dst: torch.Tensor, n: int) -> torch.Tensor:
    """Sorted unique keys lo*n + hi of the undirected simple graph: self loops dropped, (u,v) and (v,u) merged, duplicates
    merged - the `row < col` canonical form random_split_edges keeps (utils.py:99-101)."""
    lo, hi = torch.minimum(src, dst), torch.maximum(src, dst)
    return torch.unique((lo * n + hi)[lo != hi])


def sample_non_edges(keys_sorted: torch.Tensor, n: int, count: int, generator: torch.Generator) -> torch.Tensor:
    """`count` distinct uniform undirected non-edges (keys lo*n + hi, lo < hi) - the contract of the reference's negative draw
    (utils.py:127-139) without its dense N x N mask - from the hash-set sampler kernel (twowl_nonedge_sample), seeded from
    `generator`."""
    from twowl_b200 import ops
    dev = keys_sorted.device
    seed = int(torch.randint(0, 2 ** 62, (1,), generator=generator, device=dev).item())
    r, c = ops.sample_non_edges(keys_sorted // n, keys_sorted % n, n, count, seed=seed, undirected=True)
    if r.numel() < count:
        raise ValueError(f"only {r.numel()} of the {count} requested non-edges exist")
    return r * n + c


def synthetic_link_graph(n: int, src: torch.Tensor, dst: torch.Tensor, seed: int) -> Dict[str, torch.Tensor]:
    """Edge samples -> the undirected positives (shuffled, as a dataset's rows would be) and one uniform non-edge per positive:
    dict(n, pos_und int64 [2,m], neg_und int64 [2,m]) with lo < hi in every column (SURVEY 8(d) "Configs 2-5")."""
    dev = src.device
    g = torch.Generator(device=dev).manual_seed(seed + 1)
    skeys = canonical_undirected(src, dst, n)
    m = skeys.numel()
    keys = skeys[torch.randperm(m, generator=g, device=dev)]
    neg = sample_non_edges(skeys, n, m, g)
    return {"n": n, "pos_und": torch.stack((keys // n, keys % n)), "neg_und": torch.stack((neg // n, neg % n))}


def load_edge_list(path: str, device="cpu", dtype: Optional[str] = None) -> torch.Tensor:
    """int64 [2,M] edge list from `path`:
      .csv / .txt   two comma- (or whitespace-) separated integer columns, no header - the reference's raw_data format
                    (datasets.py:156-158 reads columns 0 and 1 with pandas);
      .npy          an integer array of shape [M,2] or [2,M];
      .bin / other  raw little-endian pairs (u0 v0 u1 v1 ...), `dtype` = "int32" (default) or "int64"."""
    ext = os.path.splitext(path)[1].lower()
    if ext in (".csv", ".txt"):
        import pandas as pd
        with open(path) as f:
            first = f.readline()
        df = pd.read_csv(path, header=None, sep="," if "," in first else r"\s+", usecols=[0, 1])
        arr = df.to_numpy(dtype="int64")
    elif ext == ".npy":
        arr = np.load(path, allow_pickle=False)
        if arr.ndim != 2 or 2 not in arr.shape:
            raise ValueError(f"{path}: expected an [M,2] or [2,M] integer array, got {arr.shape}")
        if arr.shape[1] != 2:
            arr = arr.T
        arr = arr.astype("int64")
    else:
        raw = np.fromfile(path, dtype=np.dtype(dtype or "int32").newbyteorder("<"))
        if raw.size % 2:
            raise ValueError(f"{path}: odd number of ids")
        arr = raw.reshape(-1, 2).astype("int64")
    if arr.size and arr.min() < 0:
        raise ValueError(f"{path}: negative node id")
    return torch.from_numpy(np.ascontiguousarray(arr.T)).to(device)
