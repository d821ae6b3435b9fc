#!/usr/bin/env python3
"""Diagnostic: v3 on the bench's C3 signals, CUDA path against the oracle, draw by draw
(rounds, recovered location sets).  Needs a GPU and ~3 min (the oracle's plan at n = 2^26)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import sfft_b200.sfft as m  # noqa: E402
from oracle import oracle  # noqa: E402


def main():
    n, k = 1 << 26, 2000
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    p = m.sfft(n, k, 3)
    op = oracle.Plan(n, k, 3)
    xs = [bench.host_signal(n, k, 1000 + i) for i in range(4)]
    xd = [torch.from_numpy(x).cuda() for x in xs]
    for step in range(steps):
        i = step % 4
        oracle.seed(17, 5000 + step)
        d = p.draw()
        cnt = p.execute_device(xd[i], d)
        loc, val = p.result()
        rounds = int(p.debug_fetch("rounds", np.int32, 1)[0])
        oracle.seed(17, 5000 + step)
        out = op.exec(xs[i])
        want = np.flatnonzero(out)
        nz = val != 0
        got = np.sort(loc[nz])
        same = np.array_equal(got, want)
        extra = np.setdiff1d(got, want).size
        missing = np.setdiff1d(want, got).size
        print(f"step {step} sig {i}: rounds gpu {rounds} oracle {op.v3_rounds}; locations {'same' if same else 'DIFFER'} "
              f"(gpu {got.size}, oracle {want.size}, gpu-only {extra}, oracle-only {missing})", flush=True)


if __name__ == "__main__":
    main()
