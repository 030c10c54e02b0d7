"""TEST INFRASTRUCTURE ONLY - the two PyG 2.3.1 utils used by datasets.py:173-197."""
import random

import numpy as np
import torch


def add_self_loops(edge_index, edge_attr=None, fill_value=None, num_nodes=None):
    """Append one (i, i) column per node; N = max id + 1 unless given."""
    n = int(edge_index.max()) + 1 if num_nodes is None else int(num_nodes)
    loop = torch.arange(n, dtype=torch.long, device=edge_index.device)
    return torch.cat([edge_index, loop.unsqueeze(0).repeat(2, 1)], dim=1), edge_attr


def negative_sampling(edge_index, num_nodes=None, num_neg_samples=None,
                      method="sparse", force_undirected=False):
    """Uniform non-edges. PyG 2.3.1 draws candidate linear ids with python ``random``
    (hence the reference is unseeded unless ``random.seed`` is set), rejects ids that
    are existing edges, and retries up to three times."""
    n = int(edge_index.max()) + 1 if num_nodes is None else int(num_nodes)
    population = n * n
    idx = (edge_index[0] * n + edge_index[1]).cpu().numpy()
    k = int(num_neg_samples if num_neg_samples is not None else edge_index.size(1))
    prob = 1.0 - idx.size / population
    sample_size = int(1.1 * k / prob)
    found = None
    for _ in range(3):
        if population <= sample_size:
            rnd = np.arange(population)
        else:
            rnd = np.asarray(random.sample(range(population), sample_size), dtype=np.int64)
        rnd = rnd[~np.isin(rnd, idx)]
        found = rnd if found is None else np.concatenate([found, rnd])
        if found.size >= k:
            found = found[:k]
            break
    found = torch.from_numpy(found).to(edge_index.device)
    return torch.stack([found // n, found % n], dim=0)
