"""BASELINE.json's configurations at FULL size.

Where the oracle finishes in about a minute (C1 n=2^22, C2 n=2^24, C5 n=2^20) the GPU result
is compared with it directly (bit-identical).  Where it does not (C3 n=2^26, C4 n=2^27: the
reference's own plan builder needs minutes on a CPU) parity rests on size-independent
properties: every planted coefficient is recovered at the right place with the right value,
the transform is a pure function of (signal, draw), the dense output equals the scattered
sparse one, and the batched path equals the single-signal path."""
import numpy as np
import pytest

from util import bits_equal

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def make_plan(n, k, version):
    import sfft_b200.sfft as m
    return m.sfft(n, k, version, strict_parameters=False)


def planted_signal(n, k, seed, snr_db=None):
    """k unit spikes -> time signal (torch FFT: only used to SYNTHESISE the input)."""
    g = torch.Generator().manual_seed(seed)
    loc = torch.unique(torch.randint(0, n, (k,), generator=g))
    xf = torch.zeros(n, dtype=torch.complex128, device="cuda")
    xf[loc.cuda()] = 1.0
    x = torch.fft.ifft(xf) * n
    if snr_db is not None:
        std = (loc.numel() / (2.0 * 10 ** (snr_db / 10.0))) ** 0.5
        gd = torch.Generator(device="cuda").manual_seed(seed + 1)
        u = torch.rand(n, generator=gd, device="cuda", dtype=torch.float64).clamp_min(1e-300)
        v = torch.rand(n, generator=gd, device="cuda", dtype=torch.float64)
        x = x + std * torch.sqrt(-2 * torch.log(u)) * torch.exp(2j * torch.pi * v)
    return x.contiguous(), loc.numpy()


@pytest.mark.parametrize("name,version,n,k", [("C1", 1, 1 << 22, 50), ("C5", 1, 1 << 20, 100),
                                              ("C2", 2, 1 << 24, 1000)])
def test_full_size_bit_identical_to_the_oracle(oracle_mod, name, version, n, k):
    x, xf = oracle_mod.generate_input(n, k, 12345)
    op = oracle_mod.Plan(n, k, version)
    p = make_plan(n, k, version)
    oracle_mod.seed(17, 1)
    cnt = p.execute_device(torch.from_numpy(x).cuda(), None)
    loc, val = p.result()
    oracle_mod.seed(17, 1)
    out = op.exec(x)
    want = np.flatnonzero(out).astype(np.int32)
    o = np.argsort(loc, kind="stable")
    assert cnt == want.size and np.array_equal(loc[o], want)
    assert bits_equal(val[o], out[want])
    true = np.flatnonzero(xf)
    assert np.abs(out[true] - xf[true]).max() < 1e-4
    p.close(); op.free()


@pytest.mark.parametrize("name,version,n,k,snr,tol", [("C4", 1, 1 << 27, 500, 20.0, 0.1),
                                                      ("C3", 3, 1 << 26, 2000, None, 0.1),
                                                      ("C4-exact", 1, 1 << 27, 500, None, 1e-5)])
def test_full_size_properties(oracle_mod, name, version, n, k, snr, tol):
    x, true = planted_signal(n, k, 7, snr)
    p = make_plan(n, k, version)
    oracle_mod.seed(17, 3)
    d = p.draw()
    cnt = p.execute_device(x, d)
    loc, val = p.result()
    assert cnt == loc.size and cnt >= 1
    # 1. every planted coefficient is there, with value 1 (the reference's acceptance bar is 0.1)
    dense = np.zeros(n, dtype=np.complex128)
    dense[loc] = val
    found = np.abs(dense[true] - 1.0) < tol
    if version == 3:
        # the reference itself is not exact at this size (1993-2000 of 2000 within
        # its own 0.1 acceptance bar, plus a few dozen spurious outputs)
        assert found.mean() > 0.98
    else:
        assert found.all()
    # 2. pure function of (signal, draw): same bits again, also after other transforms in between
    cnt2 = p.execute_device(x, p.draw())
    cnt3 = p.execute_device(x, d)
    loc3, val3 = p.result()
    o, o3 = np.argsort(loc, kind="stable"), np.argsort(loc3, kind="stable")
    assert cnt3 == cnt and np.array_equal(loc[o], loc3[o3]) and bits_equal(val[o], val3[o3])
    # 3. dense output == scattered sparse output
    dd = torch.empty(n, dtype=torch.complex128, device="cuda")
    p.densify(dd)
    nz = torch.nonzero(dd).flatten().cpu().numpy()
    assert np.array_equal(nz, np.sort(loc[val != 0]))
    assert bits_equal(dd[torch.from_numpy(nz).cuda()].cpu().numpy(), dense[nz])
    del dd
    # 4. no location outside [0, n), no duplicates
    assert loc.min() >= 0 and loc.max() < n and np.unique(loc).size == loc.size
    p.close()


def test_batch_equals_single_at_c5_shape(oracle_mod):
    n, k, num = 1 << 20, 100, 16
    p = make_plan(n, k, 1)
    xs = torch.stack([planted_signal(n, k, 50 + i)[0] for i in range(num)])
    oracle_mod.seed(17, 4)
    draws = [p.draw() for _ in range(num)]
    counts = p.execute_many_device(xs, draws)
    batch = []
    for i in range(num):
        l, v = p.result(i)
        o = np.argsort(l, kind="stable")
        batch.append((l[o], v[o]))
    for i in range(num):
        c = p.execute_device(xs[i], draws[i])
        l, v = p.result()
        o = np.argsort(l, kind="stable")
        assert c == counts[i] and np.array_equal(l[o], batch[i][0]) and bits_equal(v[o], batch[i][1])
    p.close()


def test_v2_batch_equals_single_at_c2_shape(oracle_mod):
    """The fused v2 estimation (persistent CTAs, tiles claimed from a per-signal counter)
    must give every signal of a batch exactly what it gets alone."""
    n, k, num = 1 << 24, 1000, 3
    p = make_plan(n, k, 2)
    xs = torch.stack([planted_signal(n, k, 80 + i)[0] for i in range(num)])
    oracle_mod.seed(17, 6)
    draws = [p.draw() for _ in range(num)]
    counts = p.execute_many_device(xs, draws)
    batch = []
    for i in range(num):
        l, v = p.result(i)
        o = np.argsort(l, kind="stable")
        batch.append((l[o], v[o]))
    for i in range(num):
        c = p.execute_device(xs[i], draws[i])
        l, v = p.result()
        o = np.argsort(l, kind="stable")
        assert c == counts[i] == l.size and np.unique(l).size == l.size
        assert np.array_equal(l[o], batch[i][0]) and bits_equal(v[o], batch[i][1])
    p.close()
