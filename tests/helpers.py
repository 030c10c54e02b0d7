"""Helpers shared by CPU and GPU tests (not collected by pytest)."""
import hashlib

import numpy as np
import torch


def sha(t) -> str:
    a = np.ascontiguousarray(t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else t)
    return hashlib.sha256(a.astype("<i8").tobytes()).hexdigest()


def fb_split(fb, s):
    """(pos_edge, pred_edge, pos1, y) of split s rebuilt from the golden edge lists the
    same way BaseGraph.preprocess does (TwoWL/operators/datasets.py:44-92)."""
    ep, en = fb["edge_pos"].astype(np.int64), fb["edge_neg"].astype(np.int64)
    npos, nneg = fb["num_pos"], fb["num_neg"]
    ei = ep[:, :npos[0]] if s < 2 else ep[:, :npos[0] + npos[1]]
    pos_e = [ep[:, :npos[0]], ep[:, npos[0]:npos[0] + npos[1]], ep[:, ep.shape[1] - npos[2]:]]
    neg_e = [en[:, :nneg[0]], en[:, nneg[0]:nneg[0] + nneg[1]], en[:, en.shape[1] - nneg[2]:]]
    pred = neg_e[0] if s == 0 else np.concatenate([pos_e[s], neg_e[s]], axis=1)
    pos1 = np.concatenate([ei.T, pred.T], axis=0)
    return ei, pred, pos1


def state_dict_from(fb, prefix="sd/"):
    return {k[len(prefix):]: torch.from_numpy(v) for k, v in fb.items() if k.startswith(prefix)}


def assert_close(a, b, rtol=1e-5, atol=1e-6, what=""):
    """The north_star fp32 tolerance: |a-b| <= atol + rtol*|b| (SURVEY.md 7 hard part 3)."""
    a = torch.as_tensor(a).detach().cpu().double()
    b = torch.as_tensor(b).detach().cpu().double()
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    err = (a - b).abs()
    bound = atol + rtol * b.abs()
    bad = err > bound
    assert not bool(bad.any()), (f"{what}: {int(bad.sum())}/{bad.numel()} outside tolerance; "
                                 f"max err {float(err.max()):.3e}, max |ref| {float(b.abs().max()):.3e}")
