"""the seeded hash-set non-edge sampler (csrc/sampler.cu, twowl_nonedge_sample) behind random_split_edges
(utils.py:93-147), negative_sampling / do_edge_split (datasets.py:171-206) and the synthetic generators - the OUTPUT CONTRACT of
the reference's negative draws: the requested counts, row < col for the split's negatives, no positive edge, no self loop, no
duplicate, reproducible under the seed, and what exists when the graph has fewer non-edges than asked for."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from twowl_b200 import ops
    return ops


def _graph(n, m, seed, skew=True):
    rng = np.random.default_rng(seed)
    e = np.minimum((rng.pareto(1.2, size=(2, m)) * n / 40).astype(np.int64), n - 1) if skew else rng.integers(0, n, size=(2, m))
    return torch.from_numpy(e).cuda()


@pytest.mark.parametrize("n,m,count,undirected", [(50, 200, 300, True), (2000, 30000, 30000, True), (2000, 30000, 50000, False),
                                                  (1 << 20, 2_000_000, 2_000_000, True), (7, 0, 10, True), (3000, 5000, 0, False)])
def test_sampler_contract(ops, n, m, count, undirected):
    e = _graph(n, m, n + m)
    r, c = ops.sample_non_edges(e[0], e[1], n, count, seed=123, undirected=undirected)
    assert r.dtype == torch.int64 and r.shape == c.shape == (count,)
    if count == 0:
        return
    assert int(r.min()) >= 0 and int(c.min()) >= 0 and int(r.max()) < n and int(c.max()) < n
    assert bool((r != c).all())                                                      # no self loops
    if undirected:
        assert bool((r < c).all())                                                   # upper triangle, as utils.py:101-102 keeps
        pos = torch.unique(torch.minimum(e[0], e[1]) * n + torch.maximum(e[0], e[1]))
    else:
        pos = torch.unique(e[0] * n + e[1])
    key = r * n + c
    assert torch.unique(key).numel() == count                                        # distinct
    assert not bool(torch.isin(key, pos).any())                                      # non-edges only
    r2, c2 = ops.sample_non_edges(e[0], e[1], n, count, seed=123, undirected=undirected)
    assert torch.equal(r, r2) and torch.equal(c, c2)                                 # a function of (seed, edges) only
    r3, c3 = ops.sample_non_edges(e[0], e[1], n, count, seed=124, undirected=undirected)
    assert not (torch.equal(r, r3) and torch.equal(c, c3))
    # seed from torch's default generator: reproducible under torch.manual_seed
    torch.manual_seed(9)
    a = ops.sample_non_edges(e[0], e[1], n, count, undirected=undirected)
    torch.manual_seed(9)
    b = ops.sample_non_edges(e[0], e[1], n, count, undirected=undirected)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])


def test_sampler_is_uniform_over_the_non_edges(ops):
    """Every non-edge of a small graph is drawn about equally often over many seeds (chi-square-ish bound)."""
    n = 9
    e = torch.tensor([[0, 1, 2, 3, 4, 0], [1, 2, 3, 4, 5, 8]], device="cuda")
    hits = torch.zeros(n * n, dtype=torch.int64, device="cuda")
    trials, k = 3000, 4
    for s in range(trials):
        r, c = ops.sample_non_edges(e[0], e[1], n, k, seed=s, undirected=True)
        hits.index_add_(0, r * n + c, torch.ones(k, dtype=torch.int64, device="cuda"))
    free = n * (n - 1) // 2 - 6
    assert int((hits > 0).sum()) == free
    want = trials * k / free
    got = hits[hits > 0].double()
    assert float((got - want).abs().max()) < 6 * want ** 0.5                         # ~6 sigma of a binomial count


def test_dense_graph_returns_what_exists(ops):
    n = 12
    iu = torch.triu_indices(n, n, 1).cuda()
    r, c = ops.sample_non_edges(iu[0][:-5], iu[1][:-5], n, 9, seed=1, undirected=True, rounds=64)   # all but 5 pairs are edges
    assert r.numel() == 5
    assert torch.equal(torch.sort(r * n + c).values, torch.sort(iu[0][-5:] * n + iu[1][-5:]).values)
    r, c = ops.sample_non_edges(iu[0], iu[1], n, 3, seed=1, undirected=True)          # complete graph: nothing to draw
    assert r.numel() == 0


def test_random_split_edges_and_do_edge_split_contract(ops):
    """utils.py:93-147 / datasets.py:171-206 on a skewed graph: attribute names, shapes, disjoint positive splits, negatives that
    are non-edges of the whole graph (val / test of random_split_edges) or of the split's own graph (do_edge_split)."""
    import TwoWL.operators.datasets as D
    import TwoWL.utils as U
    n = 3000
    e = _graph(n, 20000, 5)
    e = e[:, e[0] != e[1]]
    und = torch.unique(torch.minimum(e[0], e[1]) * n + torch.maximum(e[0], e[1]))
    both = torch.cat((torch.stack((und // n, und % n)), torch.stack((und % n, und // n))), dim=1)     # both directions, like a dataset
    torch.manual_seed(0)
    data = U.random_split_edges(D._Data(both.clone()), 0.05, 0.1)
    m = und.numel()
    n_v, n_t = int(0.05 * m), int(0.1 * m)
    assert data.val_pos_edge_index.shape == (2, n_v) and data.test_pos_edge_index.shape == (2, n_t)
    assert data.train_pos_edge_index.shape == (2, m - n_v - n_t)
    assert data.val_neg_edge_index.shape == (2, n_v) and data.test_neg_edge_index.shape == (2, n_t)
    key = lambda t: t[0] * n + t[1]                                                   # noqa: E731
    allpos = torch.cat((key(data.train_pos_edge_index), key(data.val_pos_edge_index), key(data.test_pos_edge_index)))
    assert torch.equal(torch.sort(allpos).values, und)                                # a partition of the row < col edges
    neg = torch.cat((key(data.val_neg_edge_index), key(data.test_neg_edge_index)))
    assert torch.unique(neg).numel() == neg.numel() and not bool(torch.isin(neg, und).any())
    assert bool((data.val_neg_edge_index[0] < data.val_neg_edge_index[1]).all())
    torch.manual_seed(0)
    again = U.random_split_edges(D._Data(both.clone()), 0.05, 0.1)
    assert torch.equal(again.val_neg_edge_index, data.val_neg_edge_index) and torch.equal(again.train_pos_edge_index, data.train_pos_edge_index)

    torch.manual_seed(1)
    split = D.do_edge_split(D._Data(both.clone()), 0.05, 0.1)
    for name, cnt in (("train", m - n_v - n_t), ("valid", n_v), ("test", n_t)):
        pos_s, neg_s = split[name]["edge"], split[name]["edge_neg"]
        assert pos_s.shape == (2, cnt) and neg_s.shape == (2, cnt)
        assert bool((neg_s[0] != neg_s[1]).all()) and torch.unique(key(neg_s)).numel() == cnt
    train_k = key(split["train"]["edge"])
    assert not bool(torch.isin(key(split["train"]["edge_neg"]), train_k).any())
    seen = torch.cat((train_k, key(split["valid"]["edge"]), key(split["test"]["edge"])))
    assert not bool(torch.isin(key(split["test"]["edge_neg"]), seen).any())           # datasets.py:192-197: against train+val+test


def test_synthetic_link_graph_on_the_sampler(ops):
    from TwoWL.operators.synthetic import canonical_undirected, rmat_edges, synthetic_link_graph
    s, d, n = rmat_edges(12, 60000, (0.57, 0.19, 0.19, 0.05), 0, "cuda")
    g = synthetic_link_graph(n, s, d, 0)
    g2 = synthetic_link_graph(n, s, d, 0)
    pos, neg = g["pos_und"], g["neg_und"]
    assert torch.equal(neg, g2["neg_und"]) and torch.equal(pos, g2["pos_und"])        # seeded
    assert pos.shape == neg.shape and pos.shape[0] == 2
    assert bool((pos[0] < pos[1]).all()) and bool((neg[0] < neg[1]).all())
    kp, kn = pos[0] * n + pos[1], neg[0] * n + neg[1]
    assert kp.unique().numel() == kp.numel() and kn.unique().numel() == kn.numel()
    assert not bool(torch.isin(kn, kp).any())
    assert torch.equal(torch.sort(kp).values, canonical_undirected(s, d, n))
    assert not torch.equal(kp, torch.sort(kp).values)                                 # rows are shuffled like a dataset's
