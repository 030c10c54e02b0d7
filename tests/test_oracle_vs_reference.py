"""Pin the oracle against the UNMODIFIED reference functions imported from
/root/reference (build container only; skipped where the reference is absent)."""
import numpy as np
import pytest
import torch

from oracle import twowl_oracle as O
from oracle.ref_import import load_reference, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="/root/reference not present")


@pytest.fixture(scope="module")
def ref():
    return load_reference()


@pytest.mark.parametrize("seed", range(12))
def test_index_operators_random_graphs(ref, seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 40))
    e = int(rng.integers(0, 120))
    p = int(rng.integers(0, 60))
    doubled = seed % 2 == 0
    if doubled:
        pos = O.double(rng.integers(0, n, size=(2, e // 2)))
        pred = O.double(rng.integers(0, n, size=(2, p // 2)))
    else:
        pos, pred = rng.integers(0, n, size=(2, e)), rng.integers(0, n, size=(2, p))
    tp, tq = torch.from_numpy(pos), torch.from_numpy(pred)
    if pos.shape[1] + pred.shape[1] == 0:
        return
    ei2_ref = ref.utils.get_ei2(n, tp, tq)
    ei2 = O.get_ei2(n, pos, pred)
    assert np.array_equal(ei2, ei2_ref.numpy().reshape(2, -1))
    assert np.array_equal(O.degree(pos, n), ref.utils.degree(tp, n).numpy())
    if pos.shape[1] == 0:
        return
    blk = rng.choice(pos.shape[1], size=max(1, pos.shape[1] // 5), replace=False)
    if ei2.shape[1]:
        assert np.array_equal(O.blockei2(ei2, blk), ref.utils.blockei2(ei2_ref, torch.from_numpy(blk)).numpy())
        a, b = O.reverse(ei2)
        ra, rb = ref.utils.reverse(ei2_ref)
        assert np.array_equal(a, ra.numpy()) and np.array_equal(b, rb.numpy())
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        r_ei, r_x, r_ei2 = ref.utils.sample_block(torch.from_numpy(blk), n, tp, ei2_ref if ei2.shape[1] else None)
    o_ei, o_x, o_ei2 = O.sample_block(blk, n, pos, ei2 if ei2.shape[1] else None)
    assert np.array_equal(o_ei, r_ei.numpy()) and np.array_equal(o_x, r_x.numpy())
    if o_ei2 is not None:
        assert np.array_equal(o_ei2, r_ei2.numpy())
    k = rng.integers(0, 50, size=7)
    assert np.array_equal(O.double(k, True), ref.utils.double(torch.from_numpy(k), True).numpy())


def test_index_operators_hypothesis_graphs(ref):
    """The same pin over hypothesis-generated inputs: multi-edges, self loops, isolated nodes, an empty prediction list, blocked ids
    that repeat - the oracle must equal the UNMODIFIED reference functions (TorchScript, utils.py:8-90) bit for bit."""
    import warnings
    from hypothesis import given, settings, strategies as st, HealthCheck

    @st.composite
    def graphs(draw):
        n = draw(st.integers(1, 20))
        node = st.integers(0, n - 1)
        und_e = draw(st.lists(st.tuples(node, node), min_size=1, max_size=30))
        und_p = draw(st.lists(st.tuples(node, node), min_size=0, max_size=15))
        return n, und_e, und_p, draw(st.booleans()), draw(st.integers(0, 2 ** 16))

    @settings(max_examples=40, deadline=None, suppress_health_check=list(HealthCheck))
    @given(graphs())
    def run(case):
        n, und_e, und_p, doubled, seed = case
        rng = np.random.default_rng(seed)
        e = np.array(und_e, dtype=np.int64).reshape(-1, 2).T
        p = np.array(und_p, dtype=np.int64).reshape(-1, 2).T
        pos, pred = (O.double(e), O.double(p)) if doubled else (e, p)
        tp, tq = torch.from_numpy(pos), torch.from_numpy(pred)
        ei2_ref = ref.utils.get_ei2(n, tp, tq)
        ei2 = O.get_ei2(n, pos, pred)
        assert np.array_equal(ei2, ei2_ref.numpy().reshape(2, -1))
        assert np.array_equal(O.degree(pos, n), ref.utils.degree(tp, n).numpy())
        blk = rng.integers(0, pos.shape[1], size=int(rng.integers(1, pos.shape[1] + 1)))      # repeats allowed
        if ei2.shape[1]:
            assert np.array_equal(O.blockei2(ei2, blk), ref.utils.blockei2(ei2_ref, torch.from_numpy(blk)).numpy())
            a, b = O.reverse(ei2)
            ra, rb = ref.utils.reverse(ei2_ref)
            assert np.array_equal(a, ra.numpy()) and np.array_equal(b, rb.numpy())
        uniq = np.unique(blk)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            r_ei, r_x, r_ei2 = ref.utils.sample_block(torch.from_numpy(uniq), n, tp, ei2_ref if ei2.shape[1] else None)
        o_ei, o_x, o_ei2 = O.sample_block(uniq, n, pos, ei2 if ei2.shape[1] else None)
        assert np.array_equal(o_ei, r_ei.numpy()) and np.array_equal(o_x, r_x.numpy())
        if o_ei2 is not None:
            assert np.array_equal(o_ei2, r_ei2.numpy())
        assert np.array_equal(O.double(e), ref.utils.double(torch.from_numpy(e)).numpy())

    run()


@pytest.mark.parametrize("cfg", [dict(c1=32, c2=16, d1=1, d2=1, a0=True, a1=True),
                                 dict(c1=24, c2=24, d1=3, d2=2, a0=False, a1=True)])
def test_model_forward_backward_matches_reference_module(ref, cfg):
    rng = np.random.default_rng(7)
    n = 50
    pos_e, pred_e = O.synthetic_split(n, rng.integers(0, n, size=(2, 160)), seed=3)
    ei2 = O.get_ei2(n, pos_e, pred_e)
    blk = O.double(rng.choice(pos_e.shape[1] // 2, size=12, replace=False), True)
    ei_new, x_new, ei2_new = O.sample_block(blk, n, pos_e, ei2)
    negs = O.double(rng.choice(pred_e.shape[1] // 2, size=12, replace=False), True) + pos_e.shape[1]
    idx = torch.from_numpy(np.concatenate([blk, negs]))
    pos1 = torch.from_numpy(np.concatenate([pos_e.T, pred_e.T]))
    y = torch.cat((torch.ones(12), torch.zeros(12))).unsqueeze(-1)
    torch.manual_seed(0)
    mod = ref.model.LocalWLNet(int(x_new.max()), False, None, channels_1wl=cfg["c1"], channels_2wl=cfg["c2"],
                               depth1=cfg["d1"], depth2=cfg["d2"], dp_lin0=0., dp_lin1=0., dp_emb=0., dp_1wl0=0.,
                               dp_2wl=0., dp_1wl1=0., act0=cfg["a0"], act1=cfg["a1"])
    x, e1 = torch.from_numpy(x_new), torch.from_numpy(ei_new)
    pred = mod(x, e1, pos1, idx, torch.from_numpy(ei2_new))
    loss = torch.nn.functional.binary_cross_entropy_with_logits(pred, y)
    loss.backward()
    sd = {k: v.detach().clone() for k, v in mod.state_dict().items()}
    o_pred, o_loss, o_grads = O.fwd_bwd(sd, x, e1, pos1, idx, ei2_new, y, cfg["a0"], cfg["a1"])
    assert torch.allclose(o_pred, pred.detach(), rtol=1e-5, atol=1e-6)
    for k, p in mod.named_parameters():
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        assert torch.allclose(o_grads[k], g, rtol=1e-5, atol=1e-6), k
    # init_state_dict produces the reference's key set and shapes
    mine = O.init_state_dict(int(x_new.max()), cfg["c1"], cfg["c2"], cfg["d1"], cfg["d2"])
    assert {k: tuple(v.shape) for k, v in mine.items()} == {k: tuple(v.shape) for k, v in sd.items()}


def test_node_feature_branch_matches_reference_module(ref):
    """use_node_feat=True (model.py:47-51,71): x = lin1(node_feat) = Dropout -> Linear -> LayerNorm(no affine) -> Dropout; with the
    dropouts at 0 the oracle's node_feat branch must reproduce the reference module's logits and gradients."""
    rng = np.random.default_rng(11)
    n, F = 40, 13
    pos_e, pred_e = O.synthetic_split(n, rng.integers(0, n, size=(2, 120)), seed=5)
    ei2 = O.get_ei2(n, pos_e, pred_e)
    blk = O.double(rng.choice(pos_e.shape[1] // 2, size=8, replace=False), True)
    ei_new, x_new, ei2_new = O.sample_block(blk, n, pos_e, ei2)
    negs = O.double(rng.choice(pred_e.shape[1] // 2, size=8, replace=False), True) + pos_e.shape[1]
    idx = torch.from_numpy(np.concatenate([blk, negs]))
    pos1 = torch.from_numpy(np.concatenate([pos_e.T, pred_e.T]))
    y = torch.cat((torch.ones(8), torch.zeros(8))).unsqueeze(-1)
    torch.manual_seed(1)
    feat = torch.randn(n, F)
    mod = ref.model.LocalWLNet(0, True, feat, channels_1wl=24, channels_2wl=16, depth1=2, depth2=1, dp_lin0=0., dp_lin1=0., dp_emb=0.,
                               dp_1wl0=0., dp_2wl=0., dp_1wl1=0.)
    x, e1 = torch.from_numpy(x_new), torch.from_numpy(ei_new)
    pred = mod(x, e1, pos1, idx, torch.from_numpy(ei2_new))
    torch.nn.functional.binary_cross_entropy_with_logits(pred, y).backward()
    sd = {k: v.detach().clone() for k, v in mod.state_dict().items()}
    assert "lin1.1.0.weight" in sd and "lin1.1.0.bias" in sd and not any(k.startswith("emb.") for k in sd)
    o_pred, _, o_grads = O.fwd_bwd(sd, x, e1, pos1, idx, ei2_new, y, True, True, node_feat=feat)
    assert torch.allclose(o_pred, pred.detach(), rtol=1e-5, atol=1e-6)
    for k, p in mod.named_parameters():
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        assert torch.allclose(o_grads[k], g, rtol=1e-5, atol=1e-6), k
