"""Shared helpers for the parity tests (test infrastructure)."""
import hashlib
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def load_golden(version, n, k, noisy=False):
    path = os.path.join(GOLDEN_DIR, f"ref_v{version}_n{n}_k{k}{'_noisy' if noisy else ''}.npz")
    return dict(np.load(path))


def golden_input(oracle, g):
    """The exact input a fixture was generated from: the reference's generator
    (src/simulation.cc:104-111) and, for noisy fixtures, AWGN continuing the same drand48
    stream (src/utils.cc:250-275).  Checked against the fixture's digest of x."""
    n, k = int(g["n"]), int(g["k"])
    x, xf = oracle.generate_input(n, k, int(g["srand48_input"]))
    if "awgn_std" in g:
        oracle.lib().orc_awgn(x.ctypes.data, n, float(g["awgn_std"]))
    assert sha(x) == str(g["sha_x"]), "oracle-generated input differs from the fixture's"
    return x, xf


def bits_equal(a, b):
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and a.tobytes() == b.tobytes()


def rel_l2(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    den = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / (den if den else 1.0))


def random_phase_spectrum_signal(oracle, n, k, seed):
    """k-sparse spectrum with random complex amplitudes (beyond the reference's own
    all-ones generator) -> time signal by the oracle's inverse DFT."""
    rng = np.random.default_rng(seed)
    loc = rng.choice(n, size=k, replace=False)
    xf = np.zeros(n, dtype=np.complex128)
    xf[loc] = (0.5 + rng.random(k)) * np.exp(2j * np.pi * rng.random(k))
    x = oracle.fft(xf, +1)
    return x, xf
