"""Pins tests/ref64.py (the chunked float64 evaluation the full-size GPU tests use as their reference) to the CPU oracle:
same inputs, the oracle on the explicit [2,T'] wedge index with plain autograd (float64), ref64 on the blocked-edge mask with
chunks far smaller than the graph. Logits, loss and every parameter gradient must agree to float64 rounding."""
import numpy as np
import pytest
import torch

import ref64
from oracle import twowl_oracle as O


def _case(n, m, seed, skew):
    rng = np.random.default_rng(seed)
    if skew:
        und = np.minimum((rng.pareto(1.5, size=(2, m)) * n / 30).astype(np.int64), n - 1)
    else:
        und = rng.integers(0, n, size=(2, m))
    pos, pred = O.synthetic_split(n, und, seed)
    E = pos.shape[1]
    nb = max(2, (E // 2) // 10)
    idx1 = O.double(rng.permutation(E // 2)[:nb], for_index=True)
    idx2 = O.double(rng.permutation(pred.shape[1] // 2)[:nb], for_index=True) + E
    pos1 = np.concatenate([pos.T, pred.T])
    y = torch.cat((torch.ones(nb), torch.zeros(nb))).unsqueeze(-1)
    return pos, pred, pos1, idx1, idx2, y


@pytest.mark.parametrize("n,m,hidden,skew,chunk", [(300, 900, 16, False, 64), (500, 2500, 24, True, 250), (200, 700, 8, True, 1 << 20)])
def test_chunked_fp64_step_equals_the_oracle(n, m, hidden, skew, chunk):
    pos, pred, pos1, idx1, idx2, y = _case(n, m, 5, skew)
    E = pos.shape[1]
    ei2 = O.get_ei2(n, pos, pred)
    ei_new, x_new, ei2_new = O.sample_block(idx1, n, pos, ei2)
    idx = np.concatenate([idx1, idx2])
    # duplicates in idx (rows selected more than once) must accumulate
    idx = np.concatenate([idx, idx[:6]])
    y = torch.cat((y, y[:3]))
    sd = O.init_state_dict(int(O.degree(pos, n).max()), hidden, hidden, 1, 1, seed=2)
    g = torch.Generator().manual_seed(7)
    sd = {k: (v + 0.2 * torch.randn(v.shape, generator=g) if v.dim() == 1 else v) for k, v in sd.items()}
    sd64 = {k: v.double() for k, v in sd.items()}
    p_o, l_o, g_o = O.fwd_bwd(sd64, torch.from_numpy(x_new), torch.from_numpy(ei_new), torch.from_numpy(pos1), torch.from_numpy(idx),
                              ei2_new, y.double())
    blocked = torch.zeros(E, dtype=torch.bool)
    blocked[torch.from_numpy(idx1)] = True
    p_r, l_r, g_r = ref64.step(sd64, torch.from_numpy(x_new), torch.from_numpy(ei_new), torch.from_numpy(pos1), torch.from_numpy(idx),
                               E, blocked, y, chunk_rows=chunk)
    assert torch.allclose(p_r, p_o, rtol=1e-11, atol=1e-12), float((p_r - p_o).abs().max())
    assert abs(float(l_r) - float(l_o)) < 1e-12
    assert g_r.keys() == g_o.keys()
    for k in g_o:
        scale = float(g_o[k].abs().max()) + 1e-30
        assert float((g_r[k] - g_o[k]).abs().max()) <= 1e-10 * scale + 1e-14, (k, float((g_r[k] - g_o[k]).abs().max()), scale)
