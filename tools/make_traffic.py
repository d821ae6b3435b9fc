#!/usr/bin/env python3
"""profiles/rNN_traffic.json from `ncu --set full` raw exports: DRAM bytes (read + write) per launch of
every kernel of a transform, summed per stage of bench.py's `stage_ms`.

    tools/make_traffic.py r02 C2=gpurun_out/z_prof_C2_raw.csv C4=... > profiles/r02_traffic.json
"""
import csv
import json
import re
import sys

STAGE_OF = [
    (r"gather_kernel", "gather"),
    (r"v2_regroup_kernel|v2_fused_kernel|estimate_", "estimate"),
    (r"fft_pass_kernel", "bucket_fft"),
    (r"select_", "select"),
    (r"vote_", "vote"),
    (r"comb_", "comb"),
    (r"v3_peel_kernel", "peel"),
    (r"v3_mansour_kernel|v3_gauss", "bucketise"),
    (r"shard_", "exchange"),
]


def stage_of(name):
    for pat, st in STAGE_OF:
        if re.search(pat, name):
            return st
    return None


def one(path):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    idx = {h: i for i, h in enumerate(hdr)}
    per_kernel = {}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        try:
            rd = float(r[idx["dram__bytes_read.sum"]].replace(",", ""))
            wr = float(r[idx["dram__bytes_write.sum"]].replace(",", ""))
        except (ValueError, KeyError):
            continue
        units = rows[1]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        rd *= scale.get(units[idx["dram__bytes_read.sum"]], 1)
        wr *= scale.get(units[idx["dram__bytes_write.sum"]], 1)
        per_kernel.setdefault(name, []).append(rd + wr)
    out, kernels = {}, {}
    for name, vals in per_kernel.items():
        st = stage_of(name)
        avg = sum(vals) / len(vals)
        short = re.sub(r"\(.*", "", name).replace("void ", "").replace("sfftb::", "")
        kernels[short] = {"launches_captured": len(vals), "dram_bytes_per_launch": avg}
        if st:
            out[st] = out.get(st, 0) + avg
    return {"per_stage_dram_bytes": out, "kernels": kernels}


def main():
    tag = sys.argv[1]
    res = {"source": "ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum per launch "
                     "(cold-cache, serialised replays); profiles/%s_*_full.txt summarise the same reports" % tag}
    for arg in sys.argv[2:]:
        wl, path = arg.split("=", 1)
        res[wl] = one(path)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
