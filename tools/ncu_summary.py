#!/usr/bin/env python
"""Summarise Nsight Compute reports for profiles/:
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep --out profiles/r1_top_kernels   -> <out>.csv (per launch) + <out>_traffic.json
    python tools/ncu_summary.py --launches gpurun_out/launches.csv --out profiles/r1_launches   -> per-kernel share of the step
Runs here (no GPU needed): `ncu -i` only reads the report."""
import argparse
import csv
import io
import json
import re
import subprocess
from collections import OrderedDict

KEEP = OrderedDict([
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex_pct"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
    ("launch__registers_per_thread", "regs"),
    ("smsp__inst_executed.sum", "warp_insts"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads_per_inst"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct"),
])


def short_name(k: str) -> str:
    m = re.search(r"(k_\w+)(<[^>]*>)?", k)
    return (m.group(1) + (m.group(2) or "")) if m else k[:60]


def to_bytes(v: str, unit: str) -> float:
    x = float(v.replace(",", ""))
    u = unit.lower()
    return x * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)


def to_us(v: str, unit: str) -> float:
    x = float(v.replace(",", ""))
    return x * {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "s": 1e6, "second": 1e6}.get(unit.lower(), 1.0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report", nargs="?")
    ap.add_argument("--launches")
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    if a.launches:
        rows = [r for r in csv.reader(open(a.launches)) if len(r) > 10]
        hdr = rows[0]
        ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
        agg = OrderedDict()
        for r in rows[1:]:
            n = short_name(r[ki])
            agg.setdefault(n, []).append(to_us(r[vi], r[ui]))
        total = sum(sum(v) for k, v in agg.items() if "flush" not in k)
        with open(a.out + ".csv", "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["kernel", "launches", "mean_us", "total_us", "share_of_step_pct"])
            for k, v in agg.items():
                w.writerow([k, len(v), round(sum(v) / len(v), 2), round(sum(v), 2), "" if "flush" in k else round(100 * sum(v) / total, 1)])
        print(open(a.out + ".csv").read())
        return
    raw = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    cols = [(hdr.index(m), n, m) for m, n in KEEP.items() if m in hdr]
    traffic = {}
    with open(a.out + ".csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + [f"{n}" for _, n, _ in cols] + ["dram_bytes_per_launch"])
        for r in rows[2:]:
            name = short_name(r[ki])
            vals, rd, wr = [], 0.0, 0.0
            for i, n, m in cols:
                if n == "duration":
                    vals.append(round(to_us(r[i], units[i]), 2))
                elif n in ("dram_read", "dram_write"):
                    b = to_bytes(r[i], units[i])
                    vals.append(int(b))
                    rd, wr = (b, wr) if n == "dram_read" else (rd, b)
                else:
                    vals.append(r[i])
            w.writerow([name] + vals + [int(rd + wr)])
            traffic.setdefault(name, []).append(rd + wr)
    json.dump({k: sum(v) / len(v) for k, v in traffic.items()}, open(a.out + "_traffic.json", "w"), indent=1)
    print(open(a.out + ".csv").read())


if __name__ == "__main__":
    main()
