"""Headless replacement for the reference's TwoWL/TwoWL_work.py: the same ``work(args, device)`` /
``read_results_twowl()`` entry points and the same result files (records_auc/<ds>_auc_record_twowl.txt,
<time_dir>/time_twowl.txt, logs.json = the best trial's parameters as one flat dict, as `json.dump(study.best_params)` writes it
at TwoWL_work.py:140-144), without Streamlit widgets and without Optuna (neither is installable offline):
hyper-parameters are drawn from the reference's search space (TwoWL_work.py:67-79) with a seeded ``random.Random``.

    python -m TwoWL.TwoWL_work --csv raw_data/fb-pages-food/fb-pages-food.csv --trials 10 --epoch 1000 --device cuda
"""
import argparse
import json
import os
import random
import time

import torch

from TwoWL.model import train
from TwoWL.model.model import LocalWLNet
from TwoWL.operators.datasets import load_dataset, dataset
from twowl_b200.optim import FusedAdam

PATH_TIME_TWOWL = "./assets/"          # constant.py:10 of the reference
SEARCH_SPACE = {                        # TwoWL_work.py:67-79
    "lr": [0.0005, 0.001, 0.005, 0.01, 0.05],
    "depth1": [1, 2, 3], "depth2": [1, 2, 3],
    "channels_1wl": [24, 32, 64], "channels_2wl": [16, 24],
    "dp_lin0": [round(0.1 * k, 1) for k in range(9)], "dp_lin1": [round(0.1 * k, 1) for k in range(9)],
    "dp_emb": [round(0.1 * k, 1) for k in range(6)], "dp_1wl0": [round(0.1 * k, 1) for k in range(6)],
    "dp_1wl1": [round(0.1 * k, 1) for k in range(6)], "dp_2wl": [round(0.1 * k, 1) for k in range(6)],
    "act0": [True, False], "act1": [True, False],
}


def _datasets(args, device):
    bg = load_dataset(args.pattern, csv=getattr(args, "csv", None), device=device)
    bg.to(device)
    bg.preprocess()
    bg.setPosDegreeFeature()
    return bg, dataset(*bg.split(0)), dataset(*bg.split(1)), dataset(*bg.split(2))


def work(args, device="cuda"):
    """TwoWL_work.py:18-149. args: pattern, epoch, and optionally csv, trials, seed, dataset, record_dir, time_dir.
    Returns {"best_params": ..., "best_val": ...} and writes the reference's result files."""
    device = torch.device(device)
    seed = getattr(args, "seed", None)
    rng = random.Random(seed)
    if seed is not None:
        torch.manual_seed(seed)
    bg, trn_ds, val_ds, tst_ds = _datasets(args, device)
    max_degree = int(torch.max(bg.x[2]).item())
    use_node_attr = trn_ds.na is not None
    dsname = getattr(args, "dataset", "fb-pages-food")
    record_dir = getattr(args, "record_dir", train.PATH_SAVE_TEST_AUC)
    time_dir = getattr(args, "time_dir", PATH_TIME_TWOWL)
    best = {"value": -1.0, "params": None}
    for trial in range(int(getattr(args, "trials", 10))):
        time_start = time.time()
        if rng.random() < 0.1:                       # TwoWL_work.py:59-66: occasionally redraw the split
            bg, trn_ds, val_ds, tst_ds = _datasets(args, device)
            max_degree = int(torch.max(bg.x[2]).item())
        setting = {k: rng.choice(v) for k, v in SEARCH_SPACE.items()}
        params = dict(setting)
        lr = setting.pop("lr")
        mod = LocalWLNet(max_degree, use_node_attr, trn_ds.na, **setting).to(device)
        opt = FusedAdam(mod.parameters(), lr=lr)       # Adam(mod.parameters(), lr) of TwoWL_work.py:100 as one kernel per step
        val = train.train_routine(dsname, mod, opt, trn_ds, val_ds, tst_ds, args.epoch, verbose=True, record_dir=record_dir,
                                  cuda_graph=getattr(args, "cuda_graph", False))
        os.makedirs(time_dir, exist_ok=True)
        with open(os.path.join(time_dir, "time_twowl.txt"), "a") as f:
            f.write("Time:" + str(round(time.time() - time_start, 4)) + "\n")
        if val > best["value"]:
            best = {"value": val, "params": params}
    with open("logs.json", "w") as f:                # TwoWL_work.py:140-144: study.best_params, dumped flat
        json.dump(best["params"], f)
    return {"best_params": best["params"], "best_val": best["value"]}


def read_results(dsname="fb-pages-food", record_dir=train.PATH_SAVE_TEST_AUC, time_dir=PATH_TIME_TWOWL, log_file="logs.json"):
    """TwoWL_work.py:152-176 with the reference's return value: (best parameters from logs.json, best test AUC of the record
    file, average trial wall time)."""
    with open(log_file) as f:
        logs = json.load(f)
    aucs, _, walls = read_results_twowl(dsname, record_dir, time_dir)
    return logs, max(aucs, default=0.0), sum(walls) / len(walls)


def read_results_twowl(dsname="fb-pages-food", record_dir=train.PATH_SAVE_TEST_AUC, time_dir=PATH_TIME_TWOWL):
    """The raw series behind read_results: (test AUCs, inference times, trial wall times) from the record files."""
    aucs, infer, walls = [], [], []
    rec = os.path.join(record_dir, f"{dsname}_auc_record_twowl.txt")
    if os.path.isfile(rec):
        for line in open(rec):
            parts = line.split()
            if len(parts) >= 2:
                aucs.append(float(parts[0].split(":")[1]))
                infer.append(float(parts[1].split(":")[1]))
    tf = os.path.join(time_dir, "time_twowl.txt")
    if os.path.isfile(tf):
        walls = [float(line.split(":")[1]) for line in open(tf) if ":" in line]
    return aucs, infer, walls


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--pattern", default="2wl_l")
    ap.add_argument("--csv", default=None)
    ap.add_argument("--epoch", type=int, default=1000)
    ap.add_argument("--trials", type=int, default=10)
    ap.add_argument("--seed", type=int, default=None)
    ap.add_argument("--device", default="cuda")
    ap.add_argument("--cuda-graph", action="store_true", help="replay the training step as one captured CUDA graph")
    a = ap.parse_args()
    print(work(a, a.device))
