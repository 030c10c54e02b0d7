"""bench.py JSON lines -> the hidden-width sweep table of BASELINE configs[4] (profiles/r2_model_sweep_rmat20.txt).
    python tools/sweep_table.py out.txt bench_h32.json bench_h64.json ..."""
import json
import sys


def load(f):
    txt = open(f).read()
    return json.loads([ln for ln in txt.splitlines() if ln.startswith("{")][-1])


out, files = sys.argv[1], sys.argv[2:]
rows = [load(f) for f in files]
ops = ["pair_conv", "seg_reduce", "pair_dw_gn", "pair_dw", "pair_init_fwd", "gn2_readout_fwd", "gn2_readout_bwd_prepare", "nccl_all_reduce", "nccl_all_gather"]
with open(out, "w") as f:
    f.write("# bench.py --hidden H on R-MAT 1M/16M (scale 20, R = 60 M pair rows, 3.0 M target links per step): one fwd+bwd step, CUDA events per op\n")
    f.write("# per op: ms per step / algorithmic GB/s / fraction of the measured copy peak (6547 GB/s); '-' = op not on this path\n")
    f.write(f"# {'hidden':>6} {'gpus':>4} {'mode':>6} {'ms/step':>9} {'links/s':>12} {'e2e ms':>8} " + " ".join(f"{o[:22]:>24}" for o in ops) + "\n")
    for d in rows:
        po = d["roofline"]["per_op"]
        cells = []
        for o in ops:
            if o in po and po[o]["ms"]:
                cells.append(f"{po[o]['ms'] / d['steps']:7.2f}/{po[o]['GBps'] or 0:6.0f}/{(po[o].get('frac') or 0):4.2f}")
            else:
                cells.append("-")
        f.write(f"  {d['config']['hidden']:>6} {d['n_gpus']:>4} {d['scaling']:>6} {d['ms_per_step']:9.2f} {d['value']:12.0f} {d['e2e']['ms_per_step']:8.2f} "
                + " ".join(f"{c:>24}" for c in cells) + "\n")
print(open(out).read())
