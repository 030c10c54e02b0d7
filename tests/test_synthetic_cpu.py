"""Input generators and edge-list reader - plain torch / numpy, run on CPU (the negative draw is a CUDA kernel:
tests/test_gpu_sampler.py)."""
import numpy as np
import pytest
import torch

from TwoWL.operators.synthetic import canonical_undirected, load_edge_list, rmat_edges


def test_rmat_edge_samples_are_seeded_in_range_and_skewed():
    s, d, n = rmat_edges(12, 60000, (0.57, 0.19, 0.19, 0.05), 0, "cpu")
    s2, d2, _ = rmat_edges(12, 60000, (0.57, 0.19, 0.19, 0.05), 0, "cpu")
    assert n == 4096 and torch.equal(s, s2) and torch.equal(d, d2)                  # seeded
    assert int(s.min()) >= 0 and int(s.max()) < n and int(d.max()) < n
    keys = canonical_undirected(s, d, n)
    assert bool((keys[1:] > keys[:-1]).all()) and bool((keys // n < keys % n).all())  # sorted, unique, lo < hi, no self loops
    deg = torch.bincount(torch.cat((keys // n, keys % n)), minlength=n)
    assert int(deg.max()) > 20 * float(deg.float().mean())                           # R-MAT skew: hubs exist


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
