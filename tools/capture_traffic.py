"""profiles/traffic.json from an `ncu --set full` capture: measured DRAM bytes per launch of the big kernels, stamped with the hash
of the kernel's source so that bench.py drops the number the moment the kernel changes.

    python tools/capture_traffic.py <capture.ncu-rep> <workload/hiddenH> [summary.txt]

Kernel -> op: k_pair_conv* -> pair_conv; k_dw_tc<C, true> -> pair_dw_gn, k_dw_tc<C, false> -> pair_dw; k_seg_rows / k_seg_chunks /
k_seg_long -> seg_reduce (the kernels of ONE ops.seg_reduce call are summed: a call is a k_seg_rows launch plus the chunk / long
passes that follow it). Also writes a small text summary (time, DRAM bytes, DRAM GB/s, L2 hit rate per launch) for profiles/.
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rows_of(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(raw)))
    hdr, data = rd[0], rd[2:]
    col = {h: i for i, h in enumerate(hdr)}

    def num(row, key):
        try:
            return float(row[col[key]].replace(",", ""))
        except (KeyError, ValueError):
            return float("nan")
    out = []
    for r in data:
        unit_scale = 1.0
        out.append(dict(name=r[col["Kernel Name"]], ns=num(r, "gpu__time_duration.sum"), rd=num(r, "dram__bytes_read.sum"),
                        wr=num(r, "dram__bytes_write.sum"), l2=num(r, "lts__t_sector_hit_rate.pct"),
                        tensor=num(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                        regs=num(r, "launch__registers_per_thread"), units=unit_scale))
    # units row: ncu prints bytes in the unit of the second header row (byte / Kbyte / Mbyte / Gbyte), time in ns / us / ms
    units = rd[1]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6,
             "nsecond": 1.0, "second": 1e9}
    for key, field in (("dram__bytes_read.sum", "rd"), ("dram__bytes_write.sum", "wr"), ("gpu__time_duration.sum", "ns")):
        if key in col:
            f = scale.get(units[col[key]], 1.0)
            for o in out:
                o[field] *= f
    return out


def op_of(name):
    if "k_pair_conv" in name:
        return "pair_conv"
    if "k_dw_tc_final" in name:
        return None
    if "k_dw_tc" in name:
        return "pair_dw_gn" if ("true" in name or ", 1>" in name or "(bool)1" in name) else "pair_dw"
    if "k_seg_rows" in name or "k_seg_chunks" in name or "k_seg_long" in name:
        return "seg_reduce"
    return None


def main():
    path, key = sys.argv[1], sys.argv[2]
    import bench
    rows = rows_of(path)
    calls = {}          # op -> list of [bytes, ns] per op call
    for r in rows:
        op = op_of(r["name"])
        if op is None or (op == "pair_conv" and "k_pair_conv<0" in r["name"]):
            continue        # k_pair_conv<0, 0> = the plain linear layers (ops.linear_*), not the pair_conv op
        lst = calls.setdefault(op, [])
        if op == "seg_reduce" and "k_seg_rows" not in r["name"] and lst:
            lst[-1][0] += r["rd"] + r["wr"]
            lst[-1][1] += r["ns"]
        else:
            lst.append([r["rd"] + r["wr"], r["ns"]])
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        table = json.load(open(tpath))
    except Exception:
        table = {}
    head = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    for op, lst in calls.items():
        b = sum(x[0] for x in lst) / len(lst)
        ns = sum(x[1] for x in lst) / len(lst)
        table.setdefault(op, {})[key] = {
            "bytes_per_launch": int(b), "launches_captured": len(lst), "avg_ms_under_ncu": round(ns / 1e6, 3),
            "dram_GBps_under_ncu": round(b / ns, 1), "source_sha16": bench.source_sha16(op), "commit": head,
            "capture": os.path.basename(path), "how": "mean over the captured launches of dram__bytes_read.sum + dram__bytes_write.sum"}
    table["_comment"] = ("dram__bytes_read.sum + dram__bytes_write.sum per launch (per ops.* call) from ncu --set full captures, written "
                         "by tools/capture_traffic.py; bench.py uses an entry only while source_sha16 matches the kernel's current source")
    with open(tpath, "w") as f:
        json.dump(table, f, indent=1, sort_keys=True)
    if len(sys.argv) > 3:
        with open(sys.argv[3], "w") as f:
            f.write(f"# ncu --set full --clock-control none, {os.path.basename(path)}, commit {head}, {key}\n")
            f.write("# time_ms   dram_read_GB   dram_write_GB   dram_GB/s   L2_hit%   tensor_pipe%   regs   kernel\n")
            for r in rows:
                if op_of(r["name"]) is None:
                    continue
                f.write(f"{r['ns'] / 1e6:8.3f} {r['rd'] / 1e9:12.2f} {r['wr'] / 1e9:14.2f} {(r['rd'] + r['wr']) / r['ns']:11.1f} "
                        f"{r['l2']:9.1f} {r['tensor']:13.1f} {int(r['regs']) if r['regs'] == r['regs'] else 0:6d}   {r['name'][:110]}\n")
        print(open(sys.argv[3]).read())
    print(json.dumps({op: table[op][key] for op in calls}, indent=1))


if __name__ == "__main__":
    main()
