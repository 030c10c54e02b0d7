"""TEST INFRASTRUCTURE ONLY - regenerate tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (needs /root/reference):  python oracle/gen_golden.py
Every array below is an output of the reference's own functions (TwoWL/utils.py,
TwoWL/operators/datasets.py, TwoWL/model/model.py imported through oracle/ref_import.py)
on seeded inputs; the fixtures travel to the GPU box, the reference does not.
Large index tensors are stored as sha256 of their int64 little-endian bytes plus shape.
"""
import hashlib
import os
import random
import sys
import warnings

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ref_import import load_reference, reference_cwd  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
warnings.filterwarnings("ignore")


def sha(t) -> str:
    a = np.ascontiguousarray(t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else t)
    return hashlib.sha256(a.astype("<i8").tobytes()).hexdigest()


def seed_all(s):
    torch.manual_seed(s)
    random.seed(s)
    np.random.seed(s)


def three_node(R):
    """SURVEY.md 8(c) worked example: path graph 0-1-2, prediction pair (0,2)."""
    u = R.utils
    pos = u.double(torch.tensor([[0, 1], [1, 2]]))
    pred = u.double(torch.tensor([[0], [2]]))
    ei2 = u.get_ei2(3, pos, pred)
    edge, edge_r = u.reverse(ei2)
    ei_new, x_new, ei2_new = u.sample_block(torch.tensor([0, 1]), 3, pos, ei2)
    np.savez(os.path.join(OUT, "three_node.npz"),
             pos=pos.numpy(), pred=pred.numpy(), ei2=ei2.contiguous().numpy(),
             edge=edge.numpy(), edge_r=edge_r.numpy(), degree=u.degree(pos, 3).numpy(),
             double_index=u.double(torch.tensor([0, 2]), for_index=True).numpy(),
             sb_ei=ei_new.numpy(), sb_x=x_new.numpy(), sb_ei2=ei2_new.numpy(),
             set_mul=u.set_mul(torch.tensor([5, 7]), torch.tensor([1, 2, 3])).numpy(),
             check_in_set=u.check_in_set(torch.tensor([1, 2, 3, 2]), torch.tensor([2, 2, 9])).numpy(),
             idx2mask=u.idx2mask(5, torch.tensor([1, 3])).numpy())


def ragged(R):
    """Seeded irregular inputs (NOT in doubled layout): duplicates, self-loop edges,
    isolated nodes, empty pred list, n_node smaller than the largest id."""
    u = R.utils
    rng = np.random.default_rng(1234)
    out = {}
    cases = [(7, 12, 5, 7), (16, 40, 0, 16), (5, 9, 9, 4), (30, 64, 31, 30), (3, 1, 1, 3), (9, 0, 6, 9)]
    for k, (n, e, p, n_arg) in enumerate(cases):
        pos = torch.from_numpy(rng.integers(0, n, size=(2, e)))
        pred = torch.from_numpy(rng.integers(0, n, size=(2, p)))
        ei2 = u.get_ei2(n_arg, pos, pred) if e + p > 0 else torch.zeros(2, 0, dtype=torch.long)
        out[f"c{k}_n"] = np.array([n_arg])
        out[f"c{k}_pos"], out[f"c{k}_pred"] = pos.numpy(), pred.numpy()
        out[f"c{k}_ei2"] = ei2.contiguous().numpy().reshape(2, -1)
        if e > 0:
            blk = torch.from_numpy(rng.choice(e, size=max(1, e // 4), replace=False))
            out[f"c{k}_blocked"] = blk.numpy()
            if ei2.numel() > 0:
                out[f"c{k}_blockei2"] = u.blockei2(ei2, blk).numpy()
            ei_new, x_new, _ = u.sample_block(blk, n, pos, None)
            out[f"c{k}_sb_ei"], out[f"c{k}_sb_x"] = ei_new.numpy(), x_new.numpy()
            out[f"c{k}_degree"] = u.degree(pos, n).numpy()
    np.savez(os.path.join(OUT, "ragged.npz"), **out)


def fb_pages_food(R):
    """configs[0]: raw_data/fb-pages-food through load_dataset('2wl_l') with the three
    RNGs seeded to 0 BEFORE the call (the reference seeds nothing)."""
    u, d, m = R.utils, R.datasets, R.model
    seed_all(0)
    with reference_cwd():
        bg = d.load_dataset("2wl_l")
    bg.preprocess()
    bg.setPosDegreeFeature()
    fix = dict(edge_pos=bg.edge_pos.numpy().astype(np.int16), edge_neg=bg.edge_neg.numpy().astype(np.int16),
               num_pos=bg.num_pos.numpy(), num_neg=bg.num_neg.numpy(),
               num_nodes=np.array([bg.num_nodes]), max_x=np.array([bg.max_x]))
    for s in range(3):
        fix[f"x{s}"] = bg.x[s].numpy().astype(np.int16)
        fix[f"ei2_{s}_shape"] = np.array(bg.ei2s[s].shape)
        fix[f"ei2_{s}_sha"] = np.array(sha(bg.ei2s[s]))
        fix[f"ei2_{s}_head"] = bg.ei2s[s][:, :64].contiguous().numpy()
        fix[f"ei2_{s}_tail"] = bg.ei2s[s][:, -64:].contiguous().numpy()
        fix[f"pos1_{s}_sha"] = np.array(sha(bg.pos1s[s]))

    # one train batch exactly as TwoWL/model/train.py:16-33 draws it
    trn = d.dataset(*bg.split(0))
    tst = d.dataset(*bg.split(2))
    trn.x, tst.x = bg.x[0], bg.x[2]
    seed_all(1)
    bs = bg.ys[1].shape[0] // 2
    perm1 = torch.randperm(trn.ei.shape[1] // 2)
    perm2 = torch.randperm((trn.pos1.shape[0] - trn.ei.shape[1]) // 2)
    idx1 = u.double(perm1[:bs], for_index=True)
    idx2 = u.double(perm2[:bs], for_index=True) + trn.ei.shape[1]
    ei_new, x_new, ei2_new = u.sample_block(idx1, trn.x.shape[0], trn.ei, trn.ei2)
    pos2 = torch.cat((idx1, idx2))
    y = torch.cat((torch.ones(bs), torch.zeros(bs))).unsqueeze(-1)
    fix.update(idx1=idx1.numpy().astype(np.int32), idx2=idx2.numpy().astype(np.int32),
               sb_ei_sha=np.array(sha(ei_new)), sb_ei_shape=np.array(ei_new.shape),
               sb_x=x_new.numpy().astype(np.int16),
               sb_ei2_sha=np.array(sha(ei2_new)), sb_ei2_shape=np.array(ei2_new.shape))
    edge, edge_r = u.reverse(ei2_new)
    fix.update(rev_edge_sha=np.array(sha(edge)), rev_edge_r_sha=np.array(sha(edge_r)))

    # model: an Optuna-space point (TwoWL_work.py:67-79) with every dropout at 0
    seed_all(2)
    cfg = dict(channels_1wl=64, channels_2wl=24, depth1=2, depth2=2, dp_lin0=0., dp_lin1=0., dp_emb=0.,
               dp_1wl0=0., dp_2wl=0., dp_1wl1=0., act0=True, act1=True)
    mod = m.LocalWLNet(bg.max_x, False, None, **cfg)
    mod.train()
    pred = mod(x_new, ei_new, trn.pos1, pos2, ei2_new)
    loss = torch.nn.functional.binary_cross_entropy_with_logits(pred, y)
    loss.backward()
    fix["train_logits"] = pred.detach().numpy()
    fix["train_loss"] = loss.detach().numpy()
    for k, v in mod.state_dict().items():
        fix["sd/" + k] = v.numpy()
    for k, p in mod.named_parameters():
        fix["grad/" + k] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy()
    mod.eval()
    with torch.no_grad():
        tp = mod(tst.x, tst.ei, tst.pos1, tst.ei.shape[1] + torch.arange(tst.y.shape[0]), tst.ei2, True)
    fix["test_logits"] = tp.numpy()
    fix["test_y"] = tst.y.numpy()

    # a second point: widths not a multiple of 32, act flags off, depth 1/3
    seed_all(3)
    cfg2 = dict(channels_1wl=24, channels_2wl=16, depth1=1, depth2=3, dp_lin0=0., dp_lin1=0., dp_emb=0.,
                dp_1wl0=0., dp_2wl=0., dp_1wl1=0., act0=False, act1=False)
    mod2 = m.LocalWLNet(bg.max_x, False, None, **cfg2)
    pred2 = mod2(x_new, ei_new, trn.pos1, pos2, ei2_new)
    loss2 = torch.nn.functional.binary_cross_entropy_with_logits(pred2, y)
    loss2.backward()
    fix["m2/train_logits"] = pred2.detach().numpy()
    for k, v in mod2.state_dict().items():
        fix["m2/sd/" + k] = v.numpy()
    for k, p in mod2.named_parameters():
        fix["m2/grad/" + k] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy()
    np.savez_compressed(os.path.join(OUT, "fb_pages_food_seed0.npz"), **fix)
    print("fb-pages-food:", {k: tuple(v.tolist()) for k, v in fix.items() if k.endswith("_shape")},
          "loss", float(loss))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    R = load_reference()
    three_node(R)
    ragged(R)
    fb_pages_food(R)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
