"""The index algebra behind the fused v2 estimation kernel (DESIGN.md section 4, K7'),
checked against the definition it replaces.  Pure Python, no GPU.

Definition (cf12.cc:366-385): hit loc = jj*W + r reads, in loop j, bucket
b = (ai*loc mod n) / seg, moved to the next bucket (mod B) when the in-bucket offset
exceeds seg/2.  Claim used by `v2_fused_kernel` (sfft_b200/csrc/v12_kernels.cu): for the
tile of hits jj = c + u*2^s (u < T) all those buckets agree modulo 2^t = q*2^s
(q = W/seg, t = logB - logT), i.e. they are ONE run of the class-major copy
xt[(b mod 2^t)*T + (b >> t)], element (p0 + m*u) mod T of it, with the same in-bucket
offset for every u -- where, with bucket0/dist0 those of jj = 0:
    m = ai mod n/W,  P = (bucket0/q + m*c) mod n/W,
    run = (bucket0 mod q) | ((P mod 2^s) * q),  p0 = P >> s.
"""
import random


def reference_bucket(n, logB, ai, loc):
    seg = n >> logB
    pos = (ai * loc) % n
    bucket, dist = divmod(pos, seg)
    if dist > seg // 2:
        bucket = (bucket + 1) % (1 << logB)
        dist -= seg
    return bucket, dist


def test_tile_reads_one_run_with_constant_offset():
    rng = random.Random(7)
    cases = 0
    while cases < 400:
        logn = rng.randint(12, 22)
        logB = rng.randint(5, logn - 2)
        logW = rng.randint(logn - logB, logn - 3)          # W a multiple of the bucket width
        logT = rng.randint(2, min(9, logn - logW, logB))
        n, W, T = 1 << logn, 1 << logW, 1 << logT
        logNW = logn - logW
        sbits = logNW - logT
        logseg = logn - logB
        logq = logW - logseg
        t = logB - logT
        assert t == logq + sbits
        ai = rng.randrange(1, n, 2)
        r = rng.randrange(W)
        c = rng.randrange(1 << sbits)
        bucket0, dist0 = reference_bucket(n, logB, ai, r)
        m = ai % (1 << logNW)
        P = ((bucket0 >> logq) + m * c) % (1 << logNW)
        run = (bucket0 & ((1 << logq) - 1)) | ((P & ((1 << sbits) - 1)) << logq)
        p0 = P >> sbits
        for u in range(T):
            jj = c + (u << sbits)
            b, d = reference_bucket(n, logB, ai, jj * W + r)
            assert d == dist0
            assert b & ((1 << t) - 1) == run
            assert b >> t == (p0 + m * u) % T
        cases += 1


def test_tiles_partition_the_prefilled_list():
    """(i, c, u) -> jj*W + r enumerates every pre-filled index exactly once (cf12.cc:505-512)."""
    logn, logW, logT = 14, 8, 3
    n, W, T = 1 << logn, 1 << logW, 1 << logT
    sbits = (logn - logW) - logT
    approved = sorted(random.Random(3).sample(range(W), 37))
    seen = set()
    for i, r in enumerate(approved):
        for c in range(1 << sbits):
            for u in range(T):
                loc = ((c + (u << sbits)) << logW) + r
                assert loc not in seen and 0 <= loc < n
                seen.add(loc)
    want = {jj * W + r for r in approved for jj in range(n // W)}
    assert seen == want
