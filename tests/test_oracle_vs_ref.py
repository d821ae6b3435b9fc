"""Pin the C restatement (oracle/sfft_oracle.c) bit-for-bit against the UNMODIFIED
reference sources compiled into oracle/_ref (skipped where /root/reference and
its build are absent, e.g. on the GPU box if _ref did not travel)."""
import numpy as np
import pytest

from util import bits_equal, random_phase_spectrum_signal

CASES = [(1, 8192, 50), (2, 8192, 50), (1, 32768, 50), (2, 65536, 50), (1, 131072, 70), (2, 262144, 50)]


@pytest.mark.parametrize("version,n,k", CASES)
def test_every_intermediate_is_bit_identical(oracle_mod, ref_mod, version, n, k):
    x, _ = ref_mod.generate_input(n, k, 555)
    x2, _ = oracle_mod.generate_input(n, k, 555)
    assert bits_equal(x, x2)
    rp = ref_mod.RefPlan(n, k, version)
    op = oracle_mod.Plan(n, k, version)
    assert bits_equal(rp.filter_time(False), op.arr("time_loc"))
    assert bits_equal(rp.filter_freq(False), op.arr("freq_loc"))
    assert bits_equal(rp.filter_time(True), op.arr("time_est"))
    assert bits_equal(rp.filter_freq(True), op.arr("freq_est"))
    for rep in range(2):     # reseed before every exec
        rp.seed(17, 100 + rep)
        ro = rp.exec(x)
        oracle_mod.seed(17, 100 + rep)
        oo = op.exec(x)
        assert bits_equal(rp.permute(), op.arr("ai"))
        assert bits_equal(rp.x_sampt(), op.arr("x_sampt"))
        assert bits_equal(rp.x_samp(), op.arr("x_samp"))
        assert bits_equal(rp.score(), op.arr("score"))
        assert bits_equal(rp.hits()[: op.hits_found], op.arr("hits"))
        if version == 2:
            assert bits_equal(rp.comb_approved()[: op.num_comb], op.arr("comb_approved"))
        assert bits_equal(ro, oo)
    rp.free()
    op.free()


def test_complex_amplitudes_expose_the_conjugate_in_estimate_values(oracle_mod, ref_mod):
    """The reference's estimate (cf12.cc:388-392) returns the CONJUGATE of bucket/filter;
    invisible with its own all-ones test spectra.  Both the compiled reference and the
    restatement must show it, identically."""
    n, k = 16384, 20
    x, xf = random_phase_spectrum_signal(oracle_mod, n, k, 3)
    rp = ref_mod.RefPlan(n, k, 1)
    op = oracle_mod.Plan(n, k, 1)
    rp.seed(17, 1)
    ro = rp.exec(x)
    oracle_mod.seed(17, 1)
    oo = op.exec(x)
    assert bits_equal(ro, oo)
    loc = np.flatnonzero(xf)
    assert np.abs(ro[loc] - np.conj(xf[loc])).max() < 1e-5
    assert np.abs(ro[loc] - xf[loc]).max() > 0.1
    rp.free()
    op.free()


def test_noisy_input_bit_identical(oracle_mod, ref_mod):
    n, k = 65536, 50
    x, _ = oracle_mod.generate_input(n, k, 9)
    oracle_mod.seed(17, 77)
    xn, snr = oracle_mod.awgn(x, np.sqrt(k / (2 * 100.0)))
    assert 50 < snr < 200      # ~20 dB
    rp = ref_mod.RefPlan(n, k, 1)
    op = oracle_mod.Plan(n, k, 1)
    rp.seed(17, 2)
    ro = rp.exec(xn)
    oracle_mod.seed(17, 2)
    oo = op.exec(xn)
    assert bits_equal(ro, oo)
    rp.free()
    op.free()
