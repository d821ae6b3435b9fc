#!/usr/bin/env python3
"""A/B harness for the permuted-filter gather (K1): one process per setting, because the knobs
(SFFTB_GATHER_FILL = 64 | 128, SFFTB_L2_FETCH = 32 | 64 | 128) are read once per process.

    SFFTB_GATHER_FILL=64 SFFTB_L2_FETCH=32 python tools/gather_ab.py C2

Prints one JSON line: event-timed gather stage (L2 flushed between transforms), samples and
Gsamples/s.  Run it under `ncu --metrics dram__bytes_read.sum,... -k regex:gather_kernel` for the
bytes per sample."""
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
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    version, n, k, snr, desc = bench.WORKLOADS[wl]
    batch = bench.BATCH.get(wl, 1)
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
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
    plan.stage_timing(True)
    acc = {}
    for i in range(reps + 2):
        flush.zero_()
        if batch > 1:
            plan.execute_many_device(x, None, sync=False)
        else:
            plan.execute_device(x, None, sync=False)
        if i >= 2:
            for nm, ms in plan.stage_times().items():
                acc.setdefault(nm, []).append(ms)
    samples = batch * (info["gather_samples"] - (info["Comb_loops"] * info["W_Comb"] if version == 2 else 0))
    extra = {k: os.environ[k] for k in ("SFFTB_GATHER_UNROLL", "SFFTB_V3_TEAM", "SFFTB_HOST_CHAINS", "SFFTB_NO_SELECT_CLUSTER") if k in os.environ}
    g = sum(acc["gather"]) / len(acc["gather"]) if "gather" in acc else None
    out = {"workload": wl, "fill": os.environ.get("SFFTB_GATHER_FILL", "auto"),
           "l2_fetch": os.environ.get("SFFTB_L2_FETCH", "default"), "gather_ms": g, "samples": samples,
           "gsamples_gathered_per_s": samples / (g * 1e-3) / 1e9 if g else None,
           "env": extra, "total_ms": sum(sum(v) / len(v) for v in acc.values()),
           "stages_ms": {nm: sum(v) / len(v) for nm, v in acc.items()}}
    print(json.dumps(out))
    plan.close()


if __name__ == "__main__":
    main()
