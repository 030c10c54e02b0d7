"""The oracle (oracle/twowl_oracle.py) against the committed golden vectors, which are
outputs of the unmodified reference functions (oracle/gen_golden.py). CPU only."""
import numpy as np
import torch

from oracle import twowl_oracle as O
from helpers import assert_close, fb_split, sha, state_dict_from


def test_three_node_worked_example(golden):
    g = golden("three_node.npz")
    pos, pred = g["pos"], g["pred"]
    assert np.array_equal(O.double(np.array([[0, 1], [1, 2]])), pos)
    assert np.array_equal(O.double(np.array([0, 2]), for_index=True), g["double_index"])
    ei2 = O.get_ei2(3, pos, pred)
    assert np.array_equal(ei2, g["ei2"])
    assert np.array_equal(ei2, [[1, 1, 0, 0, 3, 3, 2, 2], [0, 4, 1, 2, 1, 2, 3, 5]])
    assert np.array_equal(O.get_ei2_loops(3, pos, pred), g["ei2"])
    e, er = O.reverse(ei2)
    assert np.array_equal(e, g["edge"]) and np.array_equal(er, g["edge_r"])
    assert np.array_equal(O.degree(pos, 3), g["degree"])
    ei_new, x_new, ei2_new = O.sample_block(np.array([0, 1]), 3, pos, ei2)
    assert np.array_equal(ei_new, g["sb_ei"]) and np.array_equal(x_new, g["sb_x"])
    assert np.array_equal(ei2_new, g["sb_ei2"])
    assert np.array_equal(O.set_mul([5, 7], [1, 2, 3]), g["set_mul"])
    assert np.array_equal(O.check_in_set([1, 2, 3, 2], [2, 2, 9]), g["check_in_set"])
    assert np.array_equal(O.idx2mask(5, [1, 3]), g["idx2mask"])


def test_ragged_cases(golden):
    g = golden("ragged.npz")
    k = 0
    while f"c{k}_n" in g.files:
        n = int(g[f"c{k}_n"][0])
        pos, pred = g[f"c{k}_pos"], g[f"c{k}_pred"]
        ei2 = O.get_ei2(n, pos, pred)
        assert np.array_equal(ei2, g[f"c{k}_ei2"]), f"case {k}"
        assert np.array_equal(O.get_ei2_loops(n, pos, pred), g[f"c{k}_ei2"]), f"case {k}"
        if f"c{k}_blockei2" in g.files:
            assert np.array_equal(O.blockei2(ei2, g[f"c{k}_blocked"]), g[f"c{k}_blockei2"])
        if f"c{k}_sb_ei" in g.files:
            nn = int(max(pos.max(initial=0), pred.max(initial=0))) + 1
            ei_new, x_new, _ = O.sample_block(g[f"c{k}_blocked"], len(g[f"c{k}_sb_x"]), pos, None)
            assert np.array_equal(ei_new, g[f"c{k}_sb_ei"]) and np.array_equal(x_new, g[f"c{k}_sb_x"])
            assert np.array_equal(O.degree(pos, len(g[f"c{k}_degree"])), g[f"c{k}_degree"])
        k += 1
    assert k >= 6


def test_fb_pages_food_indices(fb):
    n = int(fb["num_nodes"][0])
    for s in range(3):
        ei, pred, pos1 = fb_split(fb, s)
        assert sha(pos1) == str(fb[f"pos1_{s}_sha"])
        ei2 = O.get_ei2(n, ei, pred)
        assert tuple(ei2.shape) == tuple(fb[f"ei2_{s}_shape"])
        assert sha(ei2) == str(fb[f"ei2_{s}_sha"])
        assert np.array_equal(ei2[:, :64], fb[f"ei2_{s}_head"])
        assert np.array_equal(ei2[:, -64:], fb[f"ei2_{s}_tail"])
        # T = sum_i in_E(i) * out_{E+P}(i)  
        cin = np.bincount(ei[1], minlength=n)
        cout = np.bincount(np.concatenate([ei[0], pred[0]]), minlength=n)
        assert ei2.shape[1] == int((cin * cout).sum())
    ei, pred, _ = fb_split(fb, 0)
    assert np.array_equal(O.degree(ei, n), fb["x0"]) and np.array_equal(O.degree(ei, n), fb["x1"])
    assert np.array_equal(O.degree(fb_split(fb, 1)[0], n), fb["x2"])


def test_fb_pages_food_sample_block_and_reverse(fb):
    n = int(fb["num_nodes"][0])
    ei, pred, _ = fb_split(fb, 0)
    ei2 = O.get_ei2(n, ei, pred)
    ei_new, x_new, ei2_new = O.sample_block(fb["idx1"], n, ei, ei2)
    assert sha(ei_new) == str(fb["sb_ei_sha"]) and tuple(ei_new.shape) == tuple(fb["sb_ei_shape"])
    assert np.array_equal(x_new, fb["sb_x"])
    assert sha(ei2_new) == str(fb["sb_ei2_sha"]) and tuple(ei2_new.shape) == tuple(fb["sb_ei2_shape"])
    e, er = O.reverse(ei2_new)
    assert sha(e) == str(fb["rev_edge_sha"]) and sha(er) == str(fb["rev_edge_r_sha"])


def _batch(fb):
    n = int(fb["num_nodes"][0])
    ei, pred, pos1 = fb_split(fb, 0)
    ei2 = O.get_ei2(n, ei, pred)
    ei_new, x_new, ei2_new = O.sample_block(fb["idx1"], n, ei, ei2)
    pos2 = np.concatenate([fb["idx1"], fb["idx2"]]).astype(np.int64)
    bs = len(fb["idx1"]) // 2
    y = torch.cat((torch.ones(bs), torch.zeros(bs))).unsqueeze(-1)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    return t(x_new), t(ei_new), t(pos1), t(pos2), ei2_new, y


def test_fb_pages_food_model_forward_backward(fb):
    """Logits, loss and every parameter gradient of the reference LocalWLNet train step."""
    x, e1, pos, idx, ei2, y = _batch(fb)
    for prefix, acts in (("", (True, True)), ("m2/", (False, False))):
        sd = state_dict_from(fb, prefix + "sd/")
        pred, loss, grads = O.fwd_bwd(sd, x, e1, pos, idx, ei2, y, *acts)
        assert_close(pred, fb[prefix + "train_logits"], what=prefix + "logits")
        if prefix == "":
            assert_close(loss, fb["train_loss"], what="loss")
        for k, g in grads.items():
            assert_close(g, fb[prefix + "grad/" + k], what=prefix + "grad " + k)


def test_fb_pages_food_test_inference(fb):
    n = int(fb["num_nodes"][0])
    ei, pred, pos1 = fb_split(fb, 2)
    ei2 = O.get_ei2(n, ei, pred)
    sd = state_dict_from(fb)
    idx = ei.shape[1] + torch.arange(fb["test_y"].shape[0])
    with torch.no_grad():
        out = O.local_wl_forward(sd, torch.from_numpy(fb["x2"].astype(np.int64)), torch.from_numpy(ei),
                                 torch.from_numpy(pos1), idx, ei2)
    assert_close(out, fb["test_logits"], what="test logits")
