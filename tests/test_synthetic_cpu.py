"""Input generators and edge-list reader - plain torch / numpy, run on CPU."""
import numpy as np
import pytest
import torch

from TwoWL.operators.synthetic import canonical_undirected, load_edge_list, rmat_edges, sample_non_edges, synthetic_link_graph


def test_rmat_graph_is_seeded_simple_and_skewed():
    s, d, n = rmat_edges(12, 60000, (0.57, 0.19, 0.19, 0.05), 0, "cpu")
    s2, d2, _ = rmat_edges(12, 60000, (0.57, 0.19, 0.19, 0.05), 0, "cpu")
    assert n == 4096 and torch.equal(s, s2) and torch.equal(d, d2)                  # seeded
    assert int(s.min()) >= 0 and int(s.max()) < n and int(d.max()) < n
    g = synthetic_link_graph(n, s, d, 0)
    pos, neg = g["pos_und"], g["neg_und"]
    assert pos.shape == neg.shape and pos.shape[0] == 2
    assert bool((pos[0] < pos[1]).all()) and bool((neg[0] < neg[1]).all())           # canonical, no self loops
    kp, kn = pos[0] * n + pos[1], neg[0] * n + neg[1]
    assert kp.unique().numel() == kp.numel() and kn.unique().numel() == kn.numel()   # simple graph, distinct negatives
    assert not bool(torch.isin(kn, kp).any())                                        # negatives are non-edges
    assert torch.equal(torch.sort(kp).values, canonical_undirected(s, d, n))
    deg = torch.bincount(torch.cat((pos[0], pos[1])), minlength=n)
    assert int(deg.max()) > 20 * float(deg.float().mean())                           # R-MAT skew: hubs exist
    assert not torch.equal(kp, torch.sort(kp).values)                                # rows are shuffled like a dataset's


def test_non_edge_sampler_on_a_dense_graph():
    n = 12
    iu = torch.triu_indices(n, n, 1)
    keys = (iu[0] * n + iu[1])[:-5]                                                  # all but 5 pairs are edges
    gen = torch.Generator().manual_seed(1)
    neg = sample_non_edges(torch.sort(keys).values, n, 5, gen)
    assert torch.equal(torch.sort(neg).values, torch.sort((iu[0] * n + iu[1])[-5:]).values)
    assert sample_non_edges(torch.empty(0, dtype=torch.int64), 4, 6, gen).unique().numel() == 6   # empty graph: every pair is free


def test_edge_list_formats_agree(tmp_path):
    rng = np.random.default_rng(0)
    e = rng.integers(0, 1000, size=(500, 2)).astype(np.int64)
    np.savetxt(tmp_path / "a.csv", e, fmt="%d", delimiter=",")
    np.savetxt(tmp_path / "a.txt", e, fmt="%d", delimiter=" ")
    np.save(tmp_path / "a.npy", e)
    np.save(tmp_path / "t.npy", e.T.astype(np.int32))
    e.astype("<i4").tofile(tmp_path / "a.bin")
    e.astype("<i8").tofile(tmp_path / "b.bin")
    ref = torch.from_numpy(e.T.copy())
    for name, kw in (("a.csv", {}), ("a.txt", {}), ("a.npy", {}), ("t.npy", {}), ("a.bin", {}), ("b.bin", {"dtype": "int64"})):
        got = load_edge_list(str(tmp_path / name), **kw)
        assert got.dtype == torch.int64 and torch.equal(got, ref), name
    (tmp_path / "odd.bin").write_bytes(b"\x00" * 12)
    with pytest.raises(ValueError):
        load_edge_list(str(tmp_path / "odd.bin"))
    np.save(tmp_path / "neg.npy", np.array([[0, -1]]))
    with pytest.raises(ValueError):
        load_edge_list(str(tmp_path / "neg.npy"))
