"""Oracle (C restatement) against the committed golden vectors of the compiled
reference (tests/golden/*.npz, made by tests/golden/make_golden.py)."""
import numpy as np
import pytest

from util import bits_equal, golden_input, load_golden, sha

V12_CASES = [(1, 16384, 50), (2, 16384, 50), (1, 65536, 50), (1, 262144, 100), (2, 131072, 60)]


@pytest.mark.parametrize("version,n,k", V12_CASES)
def test_oracle_matches_reference_golden_v12(oracle_mod, version, n, k):
    g = load_golden(version, n, k)
    x, xf = oracle_mod.generate_input(n, k, int(g["srand48_input"]))
    assert sha(x) == str(g["sha_x"]), "input synthesis differs from the reference harness"
    p = oracle_mod.Plan(n, k, version)
    for key in ("B_loc", "B_est", "B_thresh", "W_Comb", "Comb_loops", "loops_loc", "loops_thresh",
                "loops_est", "w_loc", "w_est", "x_samp_size"):
        assert int(g["param_" + key]) == getattr(p, key), key
    assert sha(p.arr("time_loc")) == str(g["sha_time_loc"])
    assert sha(p.arr("time_est")) == str(g["sha_time_est"])
    assert sha(p.arr("freq_loc")) == str(g["sha_freq_loc"])
    assert sha(p.arr("freq_est")) == str(g["sha_freq_est"])
    assert bits_equal(p.arr("time_loc")[:16], g["time_loc_head"])
    oracle_mod.seed(int(g["srand"]), int(g["srand48_exec"]))
    out = p.exec(x)
    assert np.array_equal(p.arr("ai"), g["permute_ai"])
    assert sha(p.arr("x_samp")) == str(g["sha_x_samp"])
    assert sha(p.arr("score")) == str(g["sha_score"])
    loc = np.flatnonzero(out).astype(np.int32)
    assert np.array_equal(loc, g["loc"]), "recovered locations differ"
    assert bits_equal(out[loc], g["val"]), "recovered values differ"
    assert sha(out) == str(g["sha_out"])
    # every planted frequency is recovered to the reference's own acceptance bar
    # (src/verification.cc:39-56: |ans - f| <= 0.1)
    true = g["true_loc"]
    assert np.all(np.abs(out[true] - xf[true]) <= 0.1)
    p.free()


def test_oracle_matches_reference_golden_noisy(oracle_mod):
    """config 4's noise level (20 dB AWGN, std = sqrt(k/200), src/utils.cc:250-275) at
    n = 2^22, k = 500: the restatement against the compiled reference, bit for bit."""
    n, k = 1 << 22, 500
    g = load_golden(1, n, k, noisy=True)
    x, xf = golden_input(oracle_mod, g)
    assert abs(10 * np.log10(float(g["awgn_snr"])) - 20.0) < 0.05
    p = oracle_mod.Plan(n, k, 1)
    assert sha(p.arr("time_loc")) == str(g["sha_time_loc"])
    oracle_mod.seed(int(g["srand"]), int(g["srand48_exec"]))
    out = p.exec(x)
    assert np.array_equal(p.arr("ai"), g["permute_ai"])
    assert sha(p.arr("x_samp")) == str(g["sha_x_samp"])
    assert sha(p.arr("score")) == str(g["sha_score"])
    loc = np.flatnonzero(out).astype(np.int32)
    assert np.array_equal(loc, g["loc"]) and bits_equal(out[loc], g["val"])
    assert np.all(np.abs(out[g["true_loc"]] - 1.0) <= 0.1)
    p.free()


def test_fft_ref_against_numpy(oracle_mod):
    rng = np.random.default_rng(1)
    for n in (2, 4, 8, 64, 1024, 8192, 15, 1000, 7507):
        x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        f = np.fft.fft(x)
        assert np.abs(oracle_mod.fft(x, -1) - f).max() <= 5e-15 * np.abs(f).max() * max(1, np.log2(n))
        b = np.fft.ifft(x) * n
        assert np.abs(oracle_mod.fft(x, +1) - b).max() <= 5e-15 * np.abs(b).max() * max(1, np.log2(n))


def test_twiddle_table_exact_points(oracle_mod):
    t = oracle_mod.twiddle_table(1024)
    assert t[0] == 1.0 and t[256] == -1j
    assert np.abs(t - np.exp(-2j * np.pi * np.arange(512) / 1024)).max() < 1e-15


def test_find_largest_indices_tie_rule(oracle_mod):
    # cutoff ties admitted in index order, output ascending (src/utils.cc:137-156)
    s = np.array([5, 1, 3, 3, 3, 9, 3, 0], dtype=np.float64)
    assert oracle_mod.find_largest_indices(s, 3).tolist() == [0, 2, 5]
    assert oracle_mod.find_largest_indices(s, 4).tolist() == [0, 2, 3, 5]
    assert oracle_mod.find_largest_indices(np.zeros(8), 3).tolist() == [0, 1, 2]


def test_integer_helpers(oracle_mod):
    L = oracle_mod.lib()
    assert L.orc_floor_to_pow2(1000.0) == 512 and L.orc_floor_to_pow2(1024.0) == 1024
    for a, n in ((3, 16), (12345, 1 << 20), (2**22 - 1, 1 << 22)):
        ai = L.orc_mod_inverse(a, n)
        assert (a * ai) % n == 1
    assert L.orc_gcd(0, 16) == 16 and L.orc_gcd(6, 16) == 2


def test_oracle_rejects_unknown_version(oracle_mod):
    assert not oracle_mod.lib().orc_make_plan(16384, 50, 7)


@pytest.mark.parametrize("n,k", [(16384, 50), (262144, 100)])
def test_oracle_matches_reference_golden_v3(oracle_mod, n, k):
    g = load_golden(3, n, k)
    x, xf = oracle_mod.generate_input(n, k, int(g["srand48_input"]))
    assert sha(x) == str(g["sha_x"])
    p = oracle_mod.Plan(n, k, 3)
    for key in ("B_g1", "w_g1", "B_g2", "w_g2", "W_Man"):
        assert int(g["param_" + key]) == getattr(p, key), key
    for nm in ("filtert1", "filterf1", "filtert2", "filterf2"):
        assert sha(p.arr(nm)) == str(g["sha_" + nm]), nm
    oracle_mod.seed(int(g["srand"]), int(g["srand48_exec"]))
    out = p.exec(x)
    loc = np.flatnonzero(out).astype(np.int32)
    assert np.array_equal(loc, g["loc"])
    assert bits_equal(out[loc], g["val"])
    assert sha(out) == str(g["sha_out"])
    p.free()
