#!/usr/bin/env python3
"""Sweep the gather's launch form in ONE process (the library reads the SFFTB_GATHER_* knobs at
every launch when SFFTB_TUNE is set):  python tools/gather_sweep.py C5
One JSON line per setting: gather stage ms (CUDA events, 256 MiB L2 flush between transforms)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import sfft_b200.sfft as m  # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "C2"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    version, n, k, snr, desc = bench.WORKLOADS[wl]
    batch = bench.BATCH.get(wl, 1)
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    os.environ["SFFTB_TUNE"] = "1"
    plan = m.sfft(n, k, version, strict_parameters=False)
    st = torch.cuda.Stream()
    torch.cuda.set_stream(st)
    plan.set_stream(st.cuda_stream)
    if batch > 1:
        x = torch.stack([bench.device_signal(torch, n, k, 9 + i, snr, dev) for i in range(batch)])
    else:
        x = bench.device_signal(torch, n, k, 9, snr, dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    info = plan.info()
    samples = batch * (info["gather_samples"] - (info["Comb_loops"] * info["W_Comb"] if version == 2 else 0))
    plan.stage_timing(True)
    settings = [("grid", u, 0) for u in (4, 8)] + [("persist", u, c) for u in (4, 8) for c in (2, 3, 4, 5)]
    for mode, unroll, ctas in settings:
        os.environ["SFFTB_GATHER_MODE"] = mode
        os.environ["SFFTB_GATHER_UNROLL"] = str(unroll)
        os.environ["SFFTB_GATHER_CTAS"] = str(ctas)
        acc = []
        for i in range(reps + 2):
            flush.zero_()
            if batch > 1:
                plan.execute_many_device(x, None, sync=False)
            else:
                plan.execute_device(x, None, sync=False)
            t = plan.stage_times()
            if i >= 2:
                acc.append(t["gather"])
        g = sum(acc) / len(acc)
        print(json.dumps({"workload": wl, "mode": mode, "unroll": unroll, "ctas_per_sm": ctas, "gather_ms": g,
                          "min_ms": min(acc), "gsamples_gathered_per_s": samples / (g * 1e-3) / 1e9}), flush=True)
    plan.close()


if __name__ == "__main__":
    main()
