#!/usr/bin/env python3
"""Summarise an ncu report (already exported with `ncu -i X.ncu-rep --page raw --csv`) or
a launch list (`--metrics gpu__time_duration.sum --csv`) into a small text table for profiles/.

  tools/ncu_summary.py raw gpurun_out/prof_raw.csv > profiles/rNN_<tag>_full.txt
  tools/ncu_summary.py launches gpurun_out/launches.csv > profiles/rNN_<tag>_launches.txt
"""
import csv
import re
import sys
from collections import OrderedDict

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("dram__bytes_read.sum.pct_of_peak_sustained_elapsed", "dram_rd%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("lts__t_sector_hit_rate.pct", "l2hit%"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "ld_sectors"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
]


def short(name):
    name = re.sub(r"\(.*", "", name)
    return name.replace("void ", "").replace("sfftb::", "").strip()[:44]


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print("kernel".ljust(46) + " ".join(lbl.rjust(12) for _, lbl in KEYS))
    for r in rows[2:]:
        out = short(r[idx["Kernel Name"]]).ljust(46)
        for k, lbl in KEYS:
            if k in idx:
                v = r[idx[k]]
                u = units[idx[k]]
                try:
                    f = float(v.replace(",", ""))
                    v = f"{f:.4g}"
                except ValueError:
                    pass
                if u in ("ms", "us", "Mbyte", "Gbyte", "Kbyte", "byte", "ns"):
                    v += u.replace("byte", "B")
                out += v.rjust(13)
            else:
                out += "-".rjust(13)
        print(out)


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.reader(lines))
    hdr = rows[0]
    idx = {h: i for i, h in enumerate(hdr)}
    agg = OrderedDict()
    total = 0.0
    for r in rows[1:]:
        if len(r) < len(hdr) or r[idx["Metric Name"]] != "gpu__time_duration.sum":
            continue
        val = float(r[idx["Metric Value"]].replace(",", ""))
        unit = r[idx["Metric Unit"]]
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
        us = val * scale
        nm = short(r[idx["Kernel Name"]])
        a = agg.setdefault(nm, [0, 0.0])
        a[0] += 1
        a[1] += us
        total += us
    print(f"{'kernel':46s} {'launches':>9s} {'total_us':>12s} {'avg_us':>10s} {'share%':>8s}")
    for nm, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{nm:46s} {c:9d} {us:12.2f} {us / c:10.2f} {100 * us / total:8.2f}")
    print(f"{'TOTAL':46s} {sum(c for c, _ in agg.values()):9d} {total:12.2f}")
    # the same list restricted to the kernels of a transform (the plan builder's FFTs and
    # sequential chains, torch's input synthesis and the legacy dense scatter run outside the
    # timed region): these shares are the ones to compare with bench.py's stage_ms
    xform = {nm: v for nm, v in agg.items()
             if re.match(r"(gather_kernel|select_|vote_kernel|estimate_|v2_|comb_|fft_pass_kernel<[01]>|"
                         r"(<unnamed>::)?v3_)", nm)}
    xt = sum(us for _, us in xform.values())
    if xt > 0:
        print()
        print(f"{'transform kernels only':46s} {'launches':>9s} {'total_us':>12s} {'avg_us':>10s} {'share%':>8s}")
        for nm, (c, us) in sorted(xform.items(), key=lambda kv: -kv[1][1]):
            print(f"{nm:46s} {c:9d} {us:12.2f} {us / c:10.2f} {100 * us / xt:8.2f}")


if __name__ == "__main__":
    {"raw": raw, "launches": launches}[sys.argv[1]](sys.argv[2])
