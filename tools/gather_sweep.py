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
    # dynamic shared memory per CTA caps the resident CTAs per SM: 0 (registers decide), 56 KB (4), 75 KB (3), 113 KB (2)
    settings = [(u, pad) for u in (2, 4, 8) for pad in (0, 56 * 1024, 75 * 1024, 113 * 1024)]
    for unroll, pad in settings:
        os.environ["SFFTB_GATHER_UNROLL"] = str(unroll)
        os.environ["SFFTB_GATHER_SMEM"] = str(pad)
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
        print(json.dumps({"workload": wl, "unroll": unroll, "smem_pad": pad, "gather_ms": g,
                          "min_ms": min(acc), "gsamples_gathered_per_s": samples / (g * 1e-3) / 1e9}), flush=True)
    plan.close()


if __name__ == "__main__":
    main()
