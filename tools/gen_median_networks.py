#!/usr/bin/env python3
"""Generate sfft_b200/csrc/median_networks.inc: for each loop count L, a straight-line
median-selection network on L registers returning v[(L-1)/2] of the ascending order -- the
element the reference takes after std::nth_element (src/computefourier-1.0-2.0.cc:406-412).

Two constructions are costed per L and the cheaper one is emitted:

  A. Batcher's merge-exchange sorting network for arbitrary N, pruned backwards to the
     comparators that can influence the wanted output wire.
  B. Sort the two halves (best known small sorter, else Batcher), then select: with
     k = (L-1)/2 + 1, the k smallest of the union are min(a[i], b[k-1-i]) (bitonic
     half-cleaner; out-of-range partners count as +inf), and the answer is their maximum
     (a tree of max operations).  Half of the last two stages' outputs are never computed.

Cost = selects a GPU needs (2 per live 64-bit output) + one compare per operation.
Everything is verified by the 0/1 principle: sorters exhaustively on 2^n inputs, complete
networks exhaustively for L <= 20, and for larger L on every pair of sorted 0/1 halves
(construction B; sufficient because all operations are monotone) or at random (A).
"""
import os
import random

import numpy as np

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                   "sfft_b200", "csrc", "median_networks.inc")
LMIN, LMAX = 2, 32

# best known small sorting networks (Knuth TAOCP 5.3.4 and later); each is verified below
KNOWN = {
    5: [(0, 1), (3, 4), (2, 4), (2, 3), (0, 3), (0, 2), (1, 4), (1, 3), (1, 2)],
    6: [(1, 2), (4, 5), (0, 2), (3, 5), (0, 1), (3, 4), (2, 5), (0, 3), (1, 4), (2, 4), (1, 3), (2, 3)],
    7: [(1, 2), (3, 4), (5, 6), (0, 2), (3, 5), (4, 6), (0, 1), (4, 5), (2, 6), (0, 4), (1, 5), (0, 3),
        (2, 5), (1, 3), (2, 4), (2, 3)],
    9: [(0, 1), (3, 4), (6, 7), (1, 2), (4, 5), (7, 8), (0, 1), (3, 4), (6, 7), (0, 3), (3, 6), (0, 3),
        (1, 4), (4, 7), (1, 4), (2, 5), (5, 8), (2, 5), (1, 3), (5, 7), (2, 6), (4, 6), (2, 4), (2, 3),
        (5, 6)],
    10: [(4, 9), (3, 8), (2, 7), (1, 6), (0, 5), (1, 4), (6, 9), (0, 3), (5, 8), (0, 2), (3, 6), (7, 9),
         (0, 1), (2, 4), (5, 7), (8, 9), (1, 2), (4, 6), (7, 8), (3, 5), (2, 5), (6, 8), (1, 3), (4, 7),
         (2, 3), (6, 7), (3, 4), (5, 6), (4, 5)],
    12: [(0, 1), (2, 3), (4, 5), (6, 7), (8, 9), (10, 11), (1, 3), (5, 7), (9, 11), (0, 2), (4, 6), (8, 10),
         (1, 2), (5, 6), (9, 10), (0, 4), (7, 11), (1, 5), (6, 10), (3, 7), (4, 8), (5, 9), (2, 6), (0, 4),
         (7, 11), (3, 8), (1, 5), (6, 10), (2, 3), (8, 9), (1, 4), (7, 10), (3, 5), (6, 8), (2, 4), (7, 9),
         (5, 6), (3, 4), (7, 8)],
}
KNOWN[11] = [c for c in KNOWN[12] if 11 not in c]


def batcher(n):
    comps = []
    p = 1
    while p < n:
        k = p
        while k >= 1:
            j = k % p
            while j <= n - 1 - k:
                for i in range(0, min(k - 1, n - j - k - 1) + 1):
                    if (i + j) // (2 * p) == (i + j + k) // (2 * p):
                        comps.append((i + j, i + j + k))
                j += 2 * k
            k //= 2
        p *= 2
    return comps


def all01(n):
    """(2^n, n) matrix of every 0/1 input."""
    idx = np.arange(1 << n, dtype=np.uint32)
    return ((idx[:, None] >> np.arange(n, dtype=np.uint32)[None, :]) & 1).astype(np.int8)


def run_ops(ops, v):
    """v: (cases, slots) array; ops act on columns.  Same tie rule as the CUDA macros."""
    v = v.copy()
    for op in ops:
        if op[0] == "cs":
            _, i, j = op
            lo = np.minimum(v[:, i], v[:, j]); hi = np.maximum(v[:, i], v[:, j])
            v[:, i], v[:, j] = lo, hi
        elif op[0] == "min":
            _, d, a, b = op
            v[:, d] = np.minimum(v[:, a], v[:, b])
        else:
            _, d, a, b = op
            v[:, d] = np.maximum(v[:, a], v[:, b])
    return v


def is_sorter(n, comps):
    v = run_ops([("cs", i, j) for i, j in comps], all01(n))
    return bool((np.diff(v, axis=1) >= 0).all())


def best_sorter(n):
    if n <= 1:
        return []
    cand = KNOWN.get(n)
    if cand is not None and is_sorter(n, cand) and len(cand) <= len(batcher(n)):
        return cand
    b = batcher(n)
    assert n > 16 or is_sorter(n, b)
    return b


def liveness(ops, want):
    """drop dead operations / dead halves of compare-exchanges; returns (ops, selects, compares)"""
    live = {want}
    out = []
    for op in reversed(ops):
        if op[0] == "cs":
            _, i, j = op
            li, lj = i in live, j in live
            if li and lj:
                out.append(op)
            elif li:
                out.append(("min", i, i, j))
            elif lj:
                out.append(("max", j, i, j))
            else:
                continue
            live.add(i); live.add(j)
        else:
            _, d, a, b = op
            if d not in live:
                continue
            out.append(op)
            live.discard(d)
            live.add(a); live.add(b)
    out.reverse()
    selects = sum(4 if op[0] == "cs" else 2 for op in out)
    return out, selects, len(out)


def construction_a(n):
    want = (n - 1) // 2
    ops = [("cs", i, j) for i, j in batcher(n)]
    return liveness(ops, want) + (want, n)


def construction_b(n):
    p = (n + 1) // 2
    q = n - p
    k = (n - 1) // 2 + 1
    ops = [("cs", i, j) for i, j in best_sorter(p)]
    ops += [("cs", p + i, p + j) for i, j in best_sorter(q)]
    slot = n
    vals = []
    for i in range(k):
        ai = i if i < p else None
        bj = p + (k - 1 - i) if 0 <= k - 1 - i < q else None
        if ai is not None and bj is not None:
            ops.append(("min", slot, ai, bj)); vals.append(slot); slot += 1
        elif ai is not None:
            vals.append(ai)
        elif bj is not None:
            vals.append(bj)
    while len(vals) > 1:
        nxt = []
        for t in range(0, len(vals) - 1, 2):
            ops.append(("max", slot, vals[t], vals[t + 1])); nxt.append(slot); slot += 1
        if len(vals) % 2:
            nxt.append(vals[-1])
        vals = nxt
    want = vals[0]
    return liveness(ops, want) + (want, slot)


def verify(n, ops, want, nslots, sorted_halves_only):
    k = (n - 1) // 2
    if n <= 20:
        v = all01(n)
    elif sorted_halves_only:
        p = (n + 1) // 2
        q = n - p
        rows = []
        rng = random.Random(n)
        for za in range(p + 1):
            for zb in range(q + 1):
                for _ in range(8):      # any arrangement inside a half sorts to the same thing
                    a = [0] * za + [1] * (p - za); b = [0] * zb + [1] * (q - zb)
                    rng.shuffle(a); rng.shuffle(b)
                    rows.append(a + b)
        v = np.array(rows, dtype=np.int8)
    else:
        rng = np.random.default_rng(n)
        v = (rng.random((400000, n)) < rng.random((400000, 1))).astype(np.int8)
    full = np.zeros((v.shape[0], nslots), dtype=v.dtype)
    full[:, :n] = v
    got = run_ops(ops, full)[:, want]
    assert (got == np.sort(v, axis=1)[:, k]).all(), n
    rng = np.random.default_rng(1000 + n)
    for data in (rng.random((20000, n)), rng.integers(0, 5, (20000, n)).astype(np.float64)):
        full = np.zeros((data.shape[0], nslots)); full[:, :n] = data
        assert (run_ops(ops, full)[:, want] == np.sort(data, axis=1)[:, k]).all(), n


def name(s, n):
    return f"v[{s}]" if s < n else f"t{s - n}"


def main(out_path=OUT):
    lines = ["// GENERATED by tools/gen_median_networks.py -- do not edit.\n",
             "// MedianNet<L>::run(v): v[(L-1)/2] of the ascending order of v[0..L), straight-line\n",
             "// compare-exchange / min / max operations on registers (see the generator for the\n",
             "// two constructions and their verification).\n",
             "#define SFFTB_CSWAP(a, b) { const double x_ = (a), y_ = (b); const bool s_ = y_ < x_; "
             "(a) = s_ ? y_ : x_; (b) = s_ ? x_ : y_; }\n",
             "#define SFFTB_MIN(a, b) ((b) < (a) ? (b) : (a))\n",
             "#define SFFTB_MAX(a, b) ((b) < (a) ? (a) : (b))\n",
             "// run_hi: the same network on 32-bit keys (the doubles' high words read as floats: one\n",
             "// FMNMX per min / max instead of a 64-bit compare and two selects per output).\n",
             "#define SFFTB_FCSWAP(a, b) { const float x_ = (a), y_ = (b); (a) = fminf(x_, y_); (b) = fmaxf(x_, y_); }\n",
             "template <int L> struct MedianNet;\n"]
    for n in range(LMIN, LMAX + 1):
        a = construction_a(n)
        b = construction_b(n)
        pick, tag = (b, "B") if (b[1] + b[2], b[2]) < (a[1] + a[2], a[2]) else (a, "A")
        ops, selects, compares, want, nslots = pick
        verify(n, ops, want, nslots, tag == "B")
        lines.append(f"template <> struct MedianNet<{n}> {{  // construction {tag}: {compares} compares, "
                     f"{selects} selects (A: {a[2]}/{a[1]}, B: {b[2]}/{b[1]})\n")
        lines.append(f"  static __device__ __forceinline__ double run(double (&v)[{n}]) {{\n")
        declared = set()
        for op in ops:
            if op[0] == "cs":
                lines.append(f"    SFFTB_CSWAP({name(op[1], n)}, {name(op[2], n)})\n")
            else:
                d = name(op[1], n)
                macro = "SFFTB_MIN" if op[0] == "min" else "SFFTB_MAX"
                decl = ""
                if op[1] >= n and op[1] not in declared:
                    declared.add(op[1]); decl = "const double "
                lines.append(f"    {decl}{d} = {macro}({name(op[2], n)}, {name(op[3], n)});\n")
        lines.append(f"    return {name(want, n)};\n  }}\n")
        # the same operations on float keys
        lines.append(f"  static __device__ __forceinline__ float run_hi(float (&v)[{n}]) {{\n")
        declared = set()
        for op in ops:
            if op[0] == "cs":
                lines.append(f"    SFFTB_FCSWAP({name(op[1], n)}, {name(op[2], n)})\n")
            else:
                d = name(op[1], n)
                fn = "fminf" if op[0] == "min" else "fmaxf"
                decl = ""
                if op[1] >= n and op[1] not in declared:
                    declared.add(op[1]); decl = "const float "
                lines.append(f"    {decl}{d} = {fn}({name(op[2], n)}, {name(op[3], n)});\n")
        lines.append(f"    return {name(want, n)};\n  }}\n}};\n")
        print(n, tag, "compares", compares, "selects", selects, "| A", a[2], a[1], "| B", b[2], b[1])
    lines.append("#undef SFFTB_CSWAP\n#undef SFFTB_FCSWAP\n#undef SFFTB_MIN\n#undef SFFTB_MAX\n")
    open(out_path, "w").writelines(lines)


if __name__ == "__main__":
    import sys
    main(sys.argv[1] if len(sys.argv) > 1 else OUT)
