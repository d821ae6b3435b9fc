"""End-to-end parity of the CUDA v1/v2 transform with the oracle, through the C ABI
(sfft_make_plan / sfftb_exec_device / sfft_exec / sfft_exec_many).

Two regimes:
  * oracle filters injected (sfftb_set_filter): every stage must be BIT-IDENTICAL;
  * filters built on the device: locations identical, values within 1e-9 relative
    L2 (north_star tolerance), filters themselves within 1e-11 of the oracle's.
"""
import ctypes as C

import numpy as np
import pytest

from util import bits_equal, load_golden, random_phase_spectrum_signal, rel_l2

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

VALUE_TOL = 1e-9      # relative L2, north_star


def make_plan(n, k, version):
    import sfft_b200.sfft as m
    return m.sfft(n, k, version, strict_parameters=False)


def fwin_from_full(freq, half):
    n = freq.size
    idx = (np.arange(-half, half + 1) + n) % n
    return np.ascontiguousarray(freq[idx])


def inject_oracle_filters(plan, op):
    for which, (t, f, B) in enumerate((("time_loc", "freq_loc", op.B_loc), ("time_est", "freq_est", op.B_est))):
        half = (op.n // B) // 2
        plan.set_filter(which, op.arr(t), fwin_from_full(op.arr(f), half))


def run_gpu(plan, oracle_mod, x, seed48):
    oracle_mod.seed(17, seed48)
    d = plan.draw()
    xd = torch.from_numpy(x).cuda()
    cnt = plan.execute_device(xd, d)
    loc, val = plan.result()
    assert loc.size == cnt
    return d, loc, val


def run_oracle(op, oracle_mod, x, seed48):
    oracle_mod.seed(17, seed48)
    out = op.exec(x)
    return out


def sorted_result(loc, val):
    o = np.argsort(loc, kind="stable")
    return loc[o], val[o]


CASES = [(1, 16384, 50), (2, 16384, 50), (1, 65536, 50), (1, 262144, 100), (2, 131072, 60),
         (1, 1 << 20, 100)]


@pytest.mark.parametrize("version,n,k", CASES)
def test_plan_parameters_match_the_reference_derivation(oracle_mod, version, n, k):
    p = make_plan(n, k, version)
    op = oracle_mod.Plan(n, k, version)
    info = p.info()
    for key in ("B_loc", "B_est", "B_thresh", "W_Comb", "Comb_loops", "loops_loc", "loops_thresh",
                "loops_est", "w_loc", "w_est", "b_loc", "b_est", "x_samp_size"):
        assert info[key] == getattr(op, key), key
    p.close()
    op.free()


@pytest.mark.parametrize("version,n,k", [(1, 16384, 50), (1, 65536, 50), (1, 1 << 18, 100), (1, 1 << 20, 100),
                                         (1, 1 << 22, 50)])
def test_device_built_filters_agree_with_the_oracle(oracle_mod, version, n, k):
    p = make_plan(n, k, version)
    op = oracle_mod.Plan(n, k, version)
    for which, (t, f, B) in enumerate((("time_loc", "freq_loc", op.B_loc), ("time_est", "freq_est", op.B_est))):
        gt, gf = p.get_filter(which)
        assert gt.size == op.arr(t).size
        # the device plan builder reproduces the reference's filter arithmetic exactly
        assert bits_equal(gt, op.arr(t)), np.abs(gt - op.arr(t)).max()
        assert bits_equal(gf, fwin_from_full(op.arr(f), (op.n // B) // 2))
    p.close()
    op.free()


@pytest.mark.parametrize("n,k", [(16384, 50), (1 << 18, 100), (1 << 22, 50)])
def test_host_thread_chains_of_the_plan_builder(oracle_mod, monkeypatch, n, k):
    """For n >= 2^24 the plan builder runs its two history-dependent recurrences (boxcar running
    sum, phase-ramp running product: src/filters.cc:119-140) on host threads instead of
    single GPU threads.  Forced on at sizes the oracle builds in seconds: same bits."""
    monkeypatch.setenv("SFFTB_HOST_CHAINS", "1")
    p = make_plan(n, k, 1)
    monkeypatch.delenv("SFFTB_HOST_CHAINS")
    op = oracle_mod.Plan(n, k, 1)
    for which, (t, f, B) in enumerate((("time_loc", "freq_loc", op.B_loc), ("time_est", "freq_est", op.B_est))):
        gt, gf = p.get_filter(which)
        assert bits_equal(gt, op.arr(t)), np.abs(gt - op.arr(t)).max()
        assert bits_equal(gf, fwin_from_full(op.arr(f), (op.n // B) // 2))
    p.close()
    op.free()


@pytest.mark.parametrize("version,n,k", CASES)
def test_every_stage_bit_identical_with_injected_filters(oracle_mod, version, n, k):
    x, xf = oracle_mod.generate_input(n, k, 4242)
    p = make_plan(n, k, version)
    op = oracle_mod.Plan(n, k, version)
    inject_oracle_filters(p, op)
    d, loc, val = run_gpu(p, oracle_mod, x, 31)
    out = run_oracle(op, oracle_mod, x, 31)
    loops = op.loops
    assert np.array_equal(np.array(d.ai[:loops]), op.arr("ai"))
    assert np.array_equal(np.array(d.a[:loops]), op.arr("a"))
    # bucket spectra (the oracle keeps natural order; so does the device after the DIT passes)
    xs = p.debug_fetch("x_samp", np.complex128, op.x_samp_size)
    assert bits_equal(xs, op.arr("x_samp")), np.abs(xs - op.arr("x_samp")).max()
    # selected buckets of the location loops
    J = p.debug_fetch("J", np.int32, op.loops_loc * op.B_thresh)
    assert np.array_equal(J, op.arr("J")[: op.loops_loc * op.B_thresh])
    if version == 2:
        appr = p.debug_fetch("comb_approved", np.int32, op.W_Comb)
        assert np.array_equal(appr, op.arr("comb_approved"))
    # voted set == {loc : score >= threshold}
    voted = np.sort(p.debug_fetch("voted", np.int32, op.n))
    assert np.array_equal(voted, np.flatnonzero(op.arr("score") >= op.loops_thresh))
    # result: same location set, same values bit for bit
    want_loc = np.flatnonzero(out).astype(np.int32)
    gl, gv = sorted_result(loc, val)
    assert np.array_equal(gl, want_loc)
    assert bits_equal(gv, out[want_loc])
    p.close()
    op.free()


@pytest.mark.parametrize("version,n,k", CASES)
def test_device_built_plan_end_to_end(oracle_mod, version, n, k):
    x, xf = oracle_mod.generate_input(n, k, 99)
    p = make_plan(n, k, version)
    op = oracle_mod.Plan(n, k, version)
    _, loc, val = run_gpu(p, oracle_mod, x, 5)
    out = run_oracle(op, oracle_mod, x, 5)
    want_loc = np.flatnonzero(out).astype(np.int32)
    gl, gv = sorted_result(loc, val)
    assert np.array_equal(gl, want_loc), "locations must be bit-exact"
    assert rel_l2(gv, out[want_loc]) < VALUE_TOL
    assert bits_equal(gv, out[want_loc]), "identical filters => identical values"
    true = np.flatnonzero(xf)
    dense = np.zeros(n, dtype=np.complex128)
    dense[gl] = gv
    assert np.abs(dense[true] - xf[true]).max() < 1e-4     # dense-FFT ground truth
    p.close()
    op.free()


@pytest.mark.parametrize("version,n,k", [(1, 16384, 50), (2, 16384, 50), (1, 65536, 50), (1, 262144, 100),
                                         (2, 131072, 60)])
def test_against_reference_golden_vectors(oracle_mod, version, n, k):
    g = load_golden(version, n, k)
    x, _ = oracle_mod.generate_input(n, k, int(g["srand48_input"]))
    p = make_plan(n, k, version)
    _, loc, val = run_gpu(p, oracle_mod, x, int(g["srand48_exec"]))
    gl, gv = sorted_result(loc, val)
    assert np.array_equal(gl, g["loc"])
    assert rel_l2(gv, g["val"]) < VALUE_TOL
    p.close()


def test_legacy_host_api_dense_output(oracle_mod):
    n, k = 65536, 50
    x, _ = oracle_mod.generate_input(n, k, 11)
    p = make_plan(n, k, 1)
    op = oracle_mod.Plan(n, k, 1)
    inject_oracle_filters(p, op)
    oracle_mod.seed(17, 3)
    got = p.execute(x)            # sfft_exec: host in, dense host out
    out = run_oracle(op, oracle_mod, x, 3)
    assert bits_equal(got, out)
    p.close()
    op.free()


def test_exec_many_draws_in_signal_order(oracle_mod):
    from sfft_b200 import _lib
    n, k, num = 32768, 50, 5
    L = _lib.load()
    p = make_plan(n, k, 1)
    op = oracle_mod.Plan(n, k, 1)
    inject_oracle_filters(p, op)
    xs = [oracle_mod.generate_input(n, k, 100 + i)[0] for i in range(num)]
    outs = [np.empty(n, dtype=np.complex128) for _ in range(num)]
    ip = (C.c_void_p * num)(*[a.ctypes.data for a in xs])
    opp = (C.c_void_p * num)(*[a.ctypes.data for a in outs])
    oracle_mod.seed(17, 8)
    L.sfft_exec_many(p.sfft_plan, num, ip, opp)
    oracle_mod.seed(17, 8)
    for i in range(num):
        want = op.exec(xs[i])      # sequential reference order: one draw per signal
        assert bits_equal(outs[i], want), i
    p.close()
    op.free()


def test_complex_amplitudes_and_noise(oracle_mod):
    n, k = 65536, 40
    x, xf = random_phase_spectrum_signal(oracle_mod, n, k, 5)
    oracle_mod.seed(17, 21)
    xn, _ = oracle_mod.awgn(x, 0.05)
    p = make_plan(n, k, 1)
    op = oracle_mod.Plan(n, k, 1)
    inject_oracle_filters(p, op)
    for sig in (x, xn):
        _, loc, val = run_gpu(p, oracle_mod, sig, 77)
        out = run_oracle(op, oracle_mod, sig, 77)
        want_loc = np.flatnonzero(out).astype(np.int32)
        gl, gv = sorted_result(loc, val)
        assert np.array_equal(gl, want_loc)
        assert bits_equal(gv, out[want_loc])
    p.close()
    op.free()


def test_unsupported_shapes_return_null(oracle_mod):
    from sfft_b200 import _lib
    L = _lib.load()
    assert not L.sfft_make_plan(1000, 50, 0, 64)           # not a power of two
    assert not L.sfft_make_plan(4194304, 4000, 1, 64)      # W_Comb < 2k+1 (reference: utils.cc:134)
    assert b"" != L.sfftb_last_error()
