"""Edge cases of the transform path, GPU vs oracle through the C ABI: degenerate inputs
(all zeros -> every magnitude ties at the cutoff; dense noise -> nothing is sparse),
the smallest and some larger documented shapes, repeated transforms on one plan
(reseeded and not), and plans of different versions living side by side (the
reference's process-global mode flags, src/common.cc:22-23, forbid that)."""
import numpy as np
import pytest

from util import bits_equal, rel_l2

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def make_plan(n, k, version):
    import sfft_b200.sfft as m
    return m.sfft(n, k, version, strict_parameters=False)


def fwin_from_full(freq, half):
    n = freq.size
    return np.ascontiguousarray(freq[(np.arange(-half, half + 1) + n) % n])


def inject(plan, op):
    for which, (t, f, B) in enumerate((("time_loc", "freq_loc", op.B_loc), ("time_est", "freq_est", op.B_est))):
        plan.set_filter(which, op.arr(t), fwin_from_full(op.arr(f), (op.n // B) // 2))


def gpu_vs_oracle(plan, op, oracle_mod, x, seed48):
    oracle_mod.seed(17, seed48)
    cnt = plan.execute_device(torch.from_numpy(np.ascontiguousarray(x)).cuda(), None)
    loc, val = plan.result()
    oracle_mod.seed(17, seed48)
    out = op.exec(x)
    o = np.argsort(loc, kind="stable")
    return loc[o], val[o], out


@pytest.mark.parametrize("version", [1, 2])
def test_all_zero_signal_exercises_the_tie_rule(oracle_mod, version):
    n, k = 16384, 50
    p, op = make_plan(n, k, version), oracle_mod.Plan(n, k, version)
    inject(p, op)
    x = np.zeros(n, dtype=np.complex128)
    loc, val, out = gpu_vs_oracle(p, op, oracle_mod, x, 1)
    # every bucket magnitude is 0: the top-2k are the first 2k indices (src/utils.cc:145-155)
    J = p.debug_fetch("J", np.int32, op.loops_loc * op.B_thresh).reshape(op.loops_loc, op.B_thresh)
    assert np.array_equal(J[0], np.arange(op.B_thresh))
    assert np.array_equal(J.ravel(), op.arr("J")[: J.size])
    voted = np.sort(p.debug_fetch("voted", np.int32, n))
    assert np.array_equal(voted, np.flatnonzero(op.arr("score") >= op.loops_thresh))
    assert np.all(val == 0) and np.all(out == 0)
    p.close(); op.free()


def test_v3_all_zero_signal(oracle_mod):
    n, k = 16384, 50
    p, op = make_plan(n, k, 3), oracle_mod.Plan(n, k, 3)
    x = np.zeros(n, dtype=np.complex128)
    loc, val, out = gpu_vs_oracle(p, op, oracle_mod, x, 1)
    assert loc.size == 0 and np.all(out == 0)
    p.close(); op.free()


@pytest.mark.parametrize("version", [1, 2])
def test_dense_noise_input_still_matches(oracle_mod, version):
    n, k = 32768, 50
    rng = np.random.default_rng(5)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    p, op = make_plan(n, k, version), oracle_mod.Plan(n, k, version)
    inject(p, op)
    loc, val, out = gpu_vs_oracle(p, op, oracle_mod, x, 2)
    want = np.flatnonzero(out).astype(np.int32)
    assert np.array_equal(loc, want)
    assert bits_equal(val, out[want])
    p.close(); op.free()


@pytest.mark.parametrize("version,n,k", [(1, 8192, 50), (2, 8192, 50), (3, 8192, 50), (1, 1 << 22, 500),
                                         (2, 1 << 19, 50), (1, 1 << 21, 2000), (3, 1 << 18, 16)])
def test_more_shapes(oracle_mod, version, n, k):
    x, xf = oracle_mod.generate_input(n, k, 404)
    p, op = make_plan(n, k, version), oracle_mod.Plan(n, k, version)
    loc, val, out = gpu_vs_oracle(p, op, oracle_mod, x, 6)
    nz = val != 0
    want = np.flatnonzero(out).astype(np.int32)
    assert np.array_equal(loc[nz], want)
    assert rel_l2(val[nz], out[want]) < 1e-9
    p.close(); op.free()


def test_repeated_transforms_and_coexisting_plans(oracle_mod):
    n, k = 65536, 50
    x, _ = oracle_mod.generate_input(n, k, 9)
    plans = {v: make_plan(n, k, v) for v in (1, 2, 3)}       # all alive at once
    oracles = {v: oracle_mod.Plan(n, k, v) for v in (1, 2, 3)}
    for rep in range(3):                                      # graph replay kicks in from the 2nd call
        for v in (2, 1, 3, 1, 2):
            loc, val, out = gpu_vs_oracle(plans[v], oracles[v], oracle_mod, x, 100 + rep)
            nz = val != 0
            want = np.flatnonzero(out).astype(np.int32)
            assert np.array_equal(loc[nz], want), (rep, v)
            assert rel_l2(val[nz], out[want]) < 1e-9
    # without reseeding the libc state advances exactly as the reference's would
    oracle_mod.seed(17, 1)
    a = [plans[1].execute_device(torch.from_numpy(x).cuda(), None) for _ in range(3)]
    oracle_mod.seed(17, 1)
    b = [int(np.count_nonzero(oracles[1].exec(x))) for _ in range(3)]
    assert a == b
    for v in (1, 2, 3):
        plans[v].close(); oracles[v].free()


def test_exec_many_device_batch_matches_single(oracle_mod):
    n, k, num = 1 << 17, 50, 7
    p, op = make_plan(n, k, 1), oracle_mod.Plan(n, k, 1)
    inject(p, op)
    xs = np.stack([oracle_mod.generate_input(n, k, 300 + i)[0] for i in range(num)])
    oracle_mod.seed(17, 2)
    counts = p.execute_many_device(torch.from_numpy(xs).cuda(), None)
    oracle_mod.seed(17, 2)
    for i in range(num):
        out = op.exec(xs[i])
        loc, val = p.result(i)
        o = np.argsort(loc, kind="stable")
        want = np.flatnonzero(out).astype(np.int32)
        assert counts[i] == want.size and np.array_equal(loc[o], want)
        assert bits_equal(val[o], out[want])
    p.close(); op.free()


def test_plan_cache_round_trip(tmp_path, oracle_mod):
    """A plan re-created from its file is the same plan: same filters, same results, bit for
    bit, for every version (include/sfft.h sfftb_save_plan / sfftb_load_plan)."""
    import sfft_b200.sfft as m
    for version, n, k in ((1, 1 << 16, 50), (2, 1 << 16, 50), (3, 1 << 16, 50)):
        x, _ = oracle_mod.generate_input(n, k, 99)
        p = m.sfft(n, k, version)
        path = tmp_path / ("plan_v%d.bin" % version)
        p.save(path)
        q = m.sfft.load(path)
        assert (q.length, q.sparsity, q.version) == (n, k, version)
        for which in (0, 1):
            ta, fa = p.get_filter(which)
            tb, fb = q.get_filter(which)
            assert bits_equal(ta, tb) and bits_equal(fa, fb)
        oracle_mod.seed(17, 5)
        a = p.execute(x)
        oracle_mod.seed(17, 5)
        b = q.execute(x)
        assert bits_equal(a, b)
        p.close(); q.close()
    # a file that is not a plan is refused loudly
    bad = tmp_path / "bad.bin"
    bad.write_bytes(b"not a plan")
    with pytest.raises(RuntimeError):
        m.sfft.load(bad)


def test_tuned_by_k_is_opt_in(oracle_mod):
    """k > 50 plans run on the reference's defaults (its by-K lookup never matches,
    src/sfft.cc:316-321) unless SFFTB_PLAN_TUNED_BY_K asks for the by-k row
    (src/parameters.cc:282-513); either way the planted coefficients come back."""
    import sfft_b200.sfft as m
    n, k = 1 << 22, 100
    x, xf = oracle_mod.generate_input(n, k, 7)
    true = np.flatnonzero(xf)
    p = m.sfft(n, k, 1)
    q = m.sfft(n, k, 1, tuned_by_k=True)
    ip, iq = p.info(), q.info()
    assert (ip["loops_loc"], ip["loops_est"]) == (4, 16)            # the defaults, sfft.cc:306-314
    assert (iq["loops_loc"], iq["loops_est"]) == (3, 12)            # parameters.cc by-K row for k = 100
    for plan in (p, q):
        oracle_mod.seed(17, 2)
        out = plan.execute(x)
        assert np.abs(out[true] - xf[true]).max() < 0.1             # verification.cc:39-56
        plan.close()


@pytest.mark.parametrize("scale", [1e-60, 1e60])
def test_v2_fused_estimation_outside_the_fast_division_band(oracle_mod, scale):
    """Spectra outside [2^-160, 2^190) make the fused v2 estimation kernel redo its tiles
    with real divisions (v12_kernels.cu: run flags -> `unsafe`); the result must still be
    the oracle's, bit for bit."""
    import sfft_b200.sfft as m
    n, k = 1 << 20, 100
    x, _ = oracle_mod.generate_input(n, k, 31)
    x = x * scale
    p = m.sfft(n, k, 2, strict_parameters=False)
    op = oracle_mod.Plan(n, k, 2)
    oracle_mod.seed(17, 9)
    cnt = p.execute_device(torch.from_numpy(x).cuda(), None)
    loc, val = p.result()
    oracle_mod.seed(17, 9)
    out = op.exec(x)
    want = np.flatnonzero(out)
    o = np.argsort(loc, kind="stable")
    keep = val[o] != 0
    assert np.array_equal(loc[o][keep], want)
    assert bits_equal(val[o][keep], out[want])
    p.close(); op.free()


def test_v2_fused_and_generic_estimation_agree(oracle_mod, monkeypatch):
    """The fused v2 estimation kernel and the generic per-hit kernel (the fallback for shapes
    the fused one does not cover; SFFTB_NO_V2_STRUCT=1 forces it) are two implementations of
    cf12.cc:341-419 over the same pre-filled list: identical results, different order."""
    import sfft_b200.sfft as m
    n, k = 1 << 20, 100
    x, _ = oracle_mod.generate_input(n, k, 5)
    xd = torch.from_numpy(x).cuda()
    fused = m.sfft(n, k, 2, strict_parameters=False)
    monkeypatch.setenv("SFFTB_NO_V2_STRUCT", "1")
    generic = m.sfft(n, k, 2, strict_parameters=False)
    monkeypatch.delenv("SFFTB_NO_V2_STRUCT")
    oracle_mod.seed(17, 11)
    d = fused.draw()
    res = []
    for p in (fused, generic):
        cnt = p.execute_device(xd, d)
        loc, val = p.result()
        assert cnt == loc.size
        o = np.argsort(loc, kind="stable")
        res.append((loc[o], val[o]))
        p.close()
    assert np.array_equal(res[0][0], res[1][0]) and bits_equal(res[0][1], res[1][1])
