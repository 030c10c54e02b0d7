"""Drop-in for the reference's TwoWL/model/train.py: ``train``, ``test``, ``train_routine`` with the
same signatures, return values and record-file formats, minus the Streamlit / matplotlib imports the reference pulls in
through ``from assets.theme import *`` (train.py:8).

What changes underneath:
  * ROC-AUC is computed on the device (``ops.auc``: radix sort + tie-aware rank sum) instead of
    ``pred.sigmoid().cpu().numpy()`` + sklearn every step (train.py:41-43, :61-66); ``train`` makes ONE host read per step
    (loss and score together) where the reference makes two.
  * ``test`` computes the ROC curve points only when asked (``curve=True``, the default, keeps the reference's
    ``(auc, fpr, tpr)`` return value); ``train_routine`` asks for them only where the reference uses them.
"""
import json
import os
import time

import torch
import torch.nn.functional as F

from TwoWL.utils import sample_block, double
from twowl_b200 import ops

PATH_SAVE_TEST_AUC = "records_auc/"   # constant.py:8 of the reference

_state = {}


def train(mod, opt, dataset, batch_size, i, step=None):
    """train.py:11-47: one batch = batch_size/2 positive + batch_size/2 negative undirected target links, their edges
    blocked from the graph (sample_block), forward, BCE-with-logits, backward, optimizer step.
    -> (loss: float, train AUC: float, next batch index: int).
    step: an optional twowl_b200.graphed.GraphedTrainStep for this dataset / batch size - edge blocking, forward, loss and
    backward are then ONE replayed CUDA graph (the small graphs are launch-bound otherwise)."""
    mod.train()
    if i == 0:
        _state["pos_bs"] = batch_size // 2
        _state["neg_bs"] = batch_size // 2
        _state["perm1"] = torch.randperm(dataset.ei.shape[1] // 2, device=dataset.x.device)
        _state["perm2"] = torch.randperm((dataset.pos1.shape[0] - dataset.ei.shape[1]) // 2, device=dataset.x.device)
    pb, nb, perm1, perm2 = _state["pos_bs"], _state["neg_bs"], _state["perm1"], _state["perm2"]
    idx1 = perm1[i * pb:(i + 1) * pb]
    idx2 = perm2[i * nb:(i + 1) * nb]
    y = torch.cat((torch.ones_like(idx1, dtype=torch.float), torch.zeros_like(idx2, dtype=torch.float)), dim=0).unsqueeze(-1)
    idx1 = double(idx1, for_index=True)
    idx2 = double(idx2, for_index=True) + dataset.ei.shape[1]
    pos2 = torch.cat((idx1, idx2), dim=0)

    opt.zero_grad()
    if step is not None:
        loss = step(idx1, pos2, y)
        pred = step.logits
    else:
        ei_new, x_new, ei2_new = sample_block(idx1, dataset.x.shape[0], dataset.ei, dataset.ei2)
        pred = mod(x_new, ei_new, dataset.pos1, pos2, ei2_new)
        loss = F.binary_cross_entropy_with_logits(pred, y)
        loss.backward()
    opt.step()

    with torch.no_grad():
        # sigmoid is monotone: the AUC of the logits is the AUC of the probabilities (ties included up to fp32 saturation,
        # so rank the probabilities like the reference does)
        res = torch.cat((loss.detach().double().reshape(1), ops.auc(pred.sigmoid(), y)[:1])).cpu()
    i += 1
    if (i + 1) * pb > perm1.shape[0]:
        i = 0
    return float(res[0]), float(res[1]), i


@torch.no_grad()
def test(mod, dataset, test=False, curve=True):
    """train.py:50-68: full-graph inference, every prediction pair is a target; AUC over one label per undirected pair.
    -> (auc, fpr, tpr); fpr / tpr are None with curve=False (no device->host copy of the scores then)."""
    mod.eval()
    pred = mod(dataset.x, dataset.ei, dataset.pos1,
               dataset.ei.shape[1] + torch.arange(dataset.y.shape[0], device=dataset.x.device), dataset.ei2, True)
    sig = pred.sigmoid()
    yy = dataset.y.reshape(-1)[0::2][: sig.shape[0]]          # the reference's interleaved True/False mask (train.py:62-64)
    result = float(ops.auc(sig, yy)[0].item())
    fpr = tpr = None
    if curve:
        from sklearn.metrics import roc_curve
        fpr, tpr, _ = roc_curve(yy.cpu().numpy(), sig.cpu().numpy().reshape(-1))
    return result, fpr, tpr


def train_routine(dsname, mod, opt, trn_ds, val_ds, tst_ds, epoch, verbose=True, record_dir=PATH_SAVE_TEST_AUC, cuda_graph=False):
    """train.py:71-135: one batch per epoch (the reference resets train_idx every epoch, train.py:87), validation every
    epoch, test on every validation improvement, early stop after 800 epochs without one; appends
    'AUC:<auc>   Time:<s>   ' to <record_dir><dsname>_auc_record_twowl.txt and keeps fpr.json / tpr.json of the best run."""
    def vprint(*args, **kwargs):
        if verbose:
            print(*args, **kwargs)

    trn_ds.pos1 = trn_ds.pos1.to(torch.long)
    val_ds.pos1 = val_ds.pos1.to(torch.long)
    tst_ds.pos1 = tst_ds.pos1.to(torch.long)
    batch_size = val_ds.y.shape[0]
    vprint(f"batch size{batch_size}")
    step = None
    if cuda_graph:   # the training step as one replayed CUDA graph (dropout included: device-resident seeds)
        from twowl_b200.graphed import GraphedTrainStep
        pb = batch_size // 2
        step = GraphedTrainStep(mod, trn_ds.x.shape[0], trn_ds.ei, trn_ds.pos1, trn_ds.ei2, n_block=2 * pb, n_links=2 * pb)

    best_val, tst_score, early_stop, early_stop_thd = 0, 0, 0, 800
    fpr = tpr = None
    t0 = t1 = 0.0
    for i in range(epoch):
        train_idx = 0
        t0 = time.time()
        loss, trn_score, train_idx = train(mod, opt, trn_ds, batch_size, train_idx, step)
        t1 = time.time()
        val_score, _, _ = test(mod, val_ds, curve=False)
        vprint(f"epoch: {i:03d}, trn: time {t1 - t0:.2f} s, loss {loss:.4f}, trn {trn_score:.4f}, val {val_score:.4f}", end=" ")
        early_stop += 1
        if val_score > best_val:
            early_stop = 0
            best_val = val_score
            if verbose:
                t0 = time.time()
                tst_score, fpr, tpr = test(mod, tst_ds, True)
                t1 = time.time()
            vprint(f"tst {tst_score:.4f}")
        else:
            vprint()
        if early_stop > early_stop_thd:
            break
    vprint(f"end test {tst_score:.3f}")
    if verbose and record_dir is not None:
        os.makedirs(record_dir, exist_ok=True)
        rec = os.path.join(record_dir, f"{dsname}_auc_record_twowl.txt")
        with open(rec, "a") as f:
            f.write("AUC:" + str(round(tst_score, 4)) + "   " + "Time:" + str(round(t1 - t0, 4)) + "   " + "\n")
        values_auc = []
        with open(rec) as f1:
            for line in f1:
                line = line.strip()
                if line:
                    values_auc.append(float(line.split()[0].split(":")[1]))
        if fpr is not None and values_auc and tst_score >= max(values_auc):
            with open("fpr.json", "w") as f:
                json.dump(fpr.tolist(), f)
            with open("tpr.json", "w") as f:
                json.dump(tpr.tolist(), f)
    return best_val
