"""Stage-level parity of the CUDA kernels against the oracle, through the C ABI's
stage hooks (include/sfft.h: sfftb_debug_*)."""
import numpy as np
import pytest

from util import bits_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    from sfft_b200 import _lib
    return _lib.load()


def _gpu_fft(L, x, log2n, batch, sign, table):
    x = np.ascontiguousarray(x, dtype=np.complex128)
    out = np.empty_like(x)
    rc = L.sfftb_debug_fft(x.ctypes.data, out.ctypes.data, log2n, batch, sign, table)
    assert rc == 0, L.sfftb_last_error()
    return out


@pytest.mark.parametrize("log2n", [1, 2, 5, 9, 11, 12, 13, 15, 17])
@pytest.mark.parametrize("sign", [-1, 1])
def test_bucket_fft_is_bit_identical_to_the_oracle(L, oracle_mod, log2n, sign):
    n, batch = 1 << log2n, 3
    rng = np.random.default_rng(log2n)
    x = rng.standard_normal((batch, n)) + 1j * rng.standard_normal((batch, n))
    got = _gpu_fft(L, x, log2n, batch, sign, 1).reshape(batch, n)
    for b in range(batch):
        want = oracle_mod.fft(x[b], sign)
        assert bits_equal(got[b], want), f"row {b}: max diff {np.abs(got[b] - want).max()}"


@pytest.mark.parametrize("log2n", [10, 16, 21])
def test_plan_builder_fft_accuracy(L, log2n):
    n = 1 << log2n
    rng = np.random.default_rng(7)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    got = _gpu_fft(L, x, log2n, 1, -1, 0)
    want = np.fft.fft(x)
    assert np.abs(got - want).max() <= 1e-14 * log2n * np.abs(want).max()


@pytest.mark.parametrize("n", [15, 607, 7507, 11765])
def test_bluestein_dft_of_odd_length(L, n):
    rng = np.random.default_rng(n)
    x = np.ascontiguousarray(rng.standard_normal(n) + 1j * rng.standard_normal(n))
    out = np.empty_like(x)
    assert L.sfftb_debug_dft_any(x.ctypes.data, out.ctypes.data, n) == 0, L.sfftb_last_error()
    want = np.fft.fft(x)
    assert np.abs(out - want).max() <= 1e-13 * np.abs(want).max()


@pytest.mark.parametrize("B,num", [(64, 10), (512, 100), (2048, 200), (8192, 100), (16384, 2000),
                                   (32768, 1000), (65536, 3000)])
def test_top_num_selection_matches_find_largest_indices(L, oracle_mod, B, num):
    rng = np.random.default_rng(B + num)
    batch = 4
    v = np.abs(rng.standard_normal((batch, B)))
    v[1] = np.round(v[1] * 8) / 8            # many exact ties, also at the cutoff
    v[2, :] = 0.25                            # all equal: pure tie rule
    v[3, : B // 2] = 0.0                      # half zeros
    v = np.ascontiguousarray(v)
    out = np.empty((batch, num), dtype=np.int32)
    assert L.sfftb_debug_select(v.ctypes.data, B, num, batch, out.ctypes.data) == 0, L.sfftb_last_error()
    for b in range(batch):
        want = oracle_mod.find_largest_indices(v[b] * v[b], num)
        assert np.array_equal(out[b], want), f"row {b}"


def test_reciprocal_division_is_ieee_exact(L):
    """The estimation kernel divides by a precomputed RN(1/den) plus two correction
    steps; it must equal IEEE division bit for bit (2^31 pseudo-random operand pairs,
    including all-ones / power-of-two divisors and equal mantissas)."""
    for seed in (1, 2):
        bad = L.sfftb_debug_div_check(seed, 1 << 30)
        assert bad == 0, f"{bad} quotients differ from __ddiv_rn"
