"""Host-side model of the 32-bit-key median of the v2 estimation kernel
(sfft_b200/csrc/v12_kernels.cu:v2_median; the reference takes the middle element after
std::nth_element, src/computefourier-1.0-2.0.cc:406-412).

The kernel runs its selection network on the quotients' high words read as floats, takes the low
word from the one quotient that carries the selected high word, and falls back to the exact 64-bit
network when several quotients carry it.  Checked here: the order claim the shortcut rests on, and
that shortcut + fallback return the same bits as a full sort -- including the tie cases.
"""
import numpy as np


def hi_lo(v):
    w = np.ascontiguousarray(v, dtype=np.float64).view(np.uint64)
    return (w >> np.uint64(32)).astype(np.uint32), (w & np.uint64(0xFFFFFFFF)).astype(np.uint32)


def in_band(rng, shape):
    """doubles as the kernel's safe tiles produce them: +0, or magnitude in [2^-758, 2^765)"""
    mant = 1.0 + rng.random(shape)
    expo = rng.integers(-758, 765, shape)
    sign = np.where(rng.random(shape) < 0.5, -1.0, 1.0)
    v = sign * np.ldexp(mant, expo)
    v[rng.random(shape) < 0.02] = 0.0
    return v


def median_by_keys(v):
    """v2_median restated: (value, used_fallback)"""
    L = v.size
    hi, lo = hi_lo(v)
    key = hi.view(np.float32)
    assert np.isfinite(key).all()
    mh = np.sort(key, kind="stable")[(L - 1) // 2].view(np.uint32)      # what MedianNet<L>::run_hi returns
    carriers = np.flatnonzero(hi == mh)
    if carriers.size != 1:
        return np.sort(v)[(L - 1) // 2], True                            # MedianNet<L>::run on the doubles
    word = (np.uint64(mh) << np.uint64(32)) | np.uint64(lo[carriers[0]])
    return np.array([word], dtype=np.uint64).view(np.float64)[0], False


def test_high_words_read_as_floats_order_like_the_doubles():
    rng = np.random.default_rng(1)
    v = in_band(rng, 400000)
    # same binade, mantissas that differ only below bit 20: equal keys, never a wrong order
    v[:1000] = 1.5 + rng.random(1000) * 2.0 ** -30
    hi, _ = hi_lo(v)
    key = hi.view(np.float32)
    assert np.isfinite(key).all() and not (np.abs(key[key != 0]) < np.finfo(np.float32).tiny).any()
    a, b = rng.integers(0, v.size, 10 ** 6), rng.integers(0, v.size, 10 ** 6)
    differ = hi[a] != hi[b]
    assert np.array_equal(key[a][differ] < key[b][differ], v[a][differ] < v[b][differ])
    assert np.array_equal(key[a][~differ] == key[b][~differ], np.ones((~differ).sum(), bool))


def test_shortcut_plus_fallback_equals_the_sorted_median():
    rng = np.random.default_rng(2)
    fallbacks = 0
    for L in (2, 3, 7, 12, 19, 20, 21, 32):
        for trial in range(400):
            v = in_band(rng, L)
            kind = trial % 4
            if kind == 1:       # the loops' estimates of a real coefficient: agree to ~1e-8
                v = 3.25 * (1.0 + 1e-8 * rng.standard_normal(L))
            elif kind == 2:     # exact duplicates of the median candidate
                v[rng.integers(0, L, 3)] = v[0]
            elif kind == 3:     # one binade, random mantissas: ties in the high word are rare
                v = 1.0 + rng.random(L)
            got, fb = median_by_keys(v)
            want = np.sort(v)[(L - 1) // 2]
            assert got.view(np.uint64) == want.view(np.uint64), (L, trial)
            fallbacks += fb
    assert fallbacks > 0          # the tie path was exercised
