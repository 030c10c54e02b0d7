"""Summarise ncu artefacts brought back in gpurun_out/ into small text files under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/launches_r1.csv profiles/r1_launches_rmat20.txt [last_n]
  python tools/summarize_ncu.py full gpurun_out/prof_seg_r1.ncu-rep profiles/r1_seg_reduce_full.txt
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name[:90]


def launches(path, out, last_n=None):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(io.StringIO("".join(lines))):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            rows.append((int(r["ID"]), r["Kernel Name"], float(r["Metric Value"].replace(",", "")) / 1e3))  # us
    if last_n:
        rows = rows[-int(last_n):]
    agg = OrderedDict()
    for _, k, us in rows:
        a = agg.setdefault(short(k), [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised): {len(rows)} launches, "
                f"{tot / 1e3:.2f} ms total\n# share%   total_ms   launches   avg_us   kernel\n")
        for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{100 * us / tot:7.2f} {us / 1e3:10.3f} {n:10d} {us / n:9.1f}   {k}\n")
    print(open(out).read()[:3000])


def full(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rd[0], rd[1], rd[2:]
    want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__occupancy_limit_registers", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__cycles_active.avg", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"]
    idx = [(w, hdr.index(w)) for w in want if w in hdr]
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none, from {path}\n")
        for row in data:
            f.write("-" * 100 + "\n")
            for w, i in idx:
                v = row[i]
                if w == "Kernel Name":
                    v = short(v)
                f.write(f"{w:80s} {v} {units[i]}\n")
            try:
                t = float(row[hdr.index('gpu__time_duration.sum')].replace(",", ""))
                tu = units[hdr.index('gpu__time_duration.sum')]
                t_s = t * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(tu, 1e-9)
                def b(name):
                    v = float(row[hdr.index(name)].replace(",", ""))
                    u = units[hdr.index(name)]
                    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
                tr = b("dram__bytes_read.sum") + b("dram__bytes_write.sum")
                f.write(f"{'=> DRAM traffic (read+write) bytes / achieved DRAM GB/s':80s} {tr:.4g} / {tr / t_s / 1e9:.1f}\n")
            except Exception as e:  # noqa
                f.write(f"(could not derive traffic: {e})\n")
    print(open(out).read()[:6000])


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(*sys.argv[2:5])
    else:
        full(sys.argv[2], sys.argv[3])
