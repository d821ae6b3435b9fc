"""oracle/oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes wrapper for oracle/liboracle.so (the C restatement in sfft_oracle.c).
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


class OCplx(C.Structure):
    _fields_ = [("re", C.c_double), ("im", C.c_double)]


_P = C.POINTER
_cp = C.c_void_p


class OrcPlan(C.Structure):
    _fields_ = [
        ("version", C.c_int), ("n_requested", C.c_int), ("n", C.c_int), ("k", C.c_int),
        ("with_comb", C.c_int),
        ("B_loc", C.c_int), ("B_est", C.c_int), ("B_thresh", C.c_int), ("W_Comb", C.c_int),
        ("Comb_loops", C.c_int),
        ("loops_loc", C.c_int), ("loops_thresh", C.c_int), ("loops_est", C.c_int),
        ("w_loc", C.c_int), ("w_est", C.c_int), ("b_loc", C.c_int), ("b_est", C.c_int),
        ("tolerance_loc", C.c_double), ("tolerance_est", C.c_double),
        ("lobefrac_loc", C.c_double), ("lobefrac_est", C.c_double),
        ("time_loc", _cp), ("freq_loc", _cp), ("time_est", _cp), ("freq_est", _cp),
        ("x_samp_size", C.c_long),
        ("a", _cp), ("ai", _cp),
        ("x_sampt", _cp), ("x_samp", _cp),
        ("mag", _cp),
        ("J", _cp),
        ("score", _cp),
        ("hits", _cp), ("hits_found", C.c_long),
        ("hits_prefill", C.c_long),
        ("comb_approved", _cp), ("num_comb", C.c_int), ("comb_offsets", _cp),
        ("comb_spec", _cp),
        ("B_g1", C.c_int), ("w_g1", C.c_int), ("B_g2", C.c_int), ("w_g2", C.c_int),
        ("W_Man", C.c_int),
        ("filtert1", _cp), ("filterf1", _cp), ("filtert2", _cp), ("filterf2", _cp),
        ("man_samp", _cp), ("gauss_samp", _cp), ("gauss_perm_samp", _cp), ("perm_x", _cp),
        ("v3_a", C.c_int), ("v3_ai", C.c_int), ("v3_b", C.c_int), ("v3_shift", C.c_int),
        ("v3_init_offset", C.c_int), ("v3_init_G_offset", C.c_int),
        ("v3_keys", _cp), ("v3_vals", _cp), ("v3_count", C.c_int), ("v3_cap", C.c_int),
        ("v3_rounds", C.c_int),
        ("tw", _cp), ("tw_n", C.c_long),
    ]


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])


def lib():
    global _lib
    if _lib is not None:
        return _lib
    path = os.path.join(_HERE, "liboracle.so")
    if not os.path.exists(path):
        build()
    L = C.CDLL(path, mode=C.RTLD_LOCAL)
    L.orc_make_plan.restype = _P(OrcPlan)
    L.orc_make_plan.argtypes = [C.c_int, C.c_int, C.c_int]
    L.orc_free_plan.argtypes = [_P(OrcPlan)]
    L.orc_exec.argtypes = [_P(OrcPlan), _cp, _cp]
    for nm in ("orc_draw_permutations", "orc_bucket_ffts", "orc_select_and_vote"):
        getattr(L, nm).argtypes = [_P(OrcPlan)]
    for nm in ("orc_comb_stage", "orc_bucketize", "orc_estimate"):
        getattr(L, nm).argtypes = [_P(OrcPlan), _cp]
    L.orc_floor_to_pow2.restype = C.c_int
    L.orc_floor_to_pow2.argtypes = [C.c_double]
    L.orc_mod_inverse.restype = C.c_int
    L.orc_mod_inverse.argtypes = [C.c_int, C.c_int]
    L.orc_gcd.restype = C.c_int
    L.orc_gcd.argtypes = [C.c_int, C.c_int]
    L.orc_find_largest_indices.argtypes = [_cp, C.c_int, _cp, C.c_int]
    L.orc_dolph_chebyshev.restype = _cp
    L.orc_dolph_chebyshev.argtypes = [C.c_double, C.c_double, _P(C.c_int)]
    L.orc_dolph_chebyshev_samples.argtypes = [C.c_double, C.c_double, C.c_int, _cp]
    L.orc_make_multiple.restype = _cp
    L.orc_make_multiple.argtypes = [_cp, C.c_int, C.c_int, C.c_int]
    L.orc_generate_input.argtypes = [C.c_int, C.c_int, _cp, _cp]
    L.orc_awgn.restype = C.c_double
    L.orc_awgn.argtypes = [_cp, C.c_int, C.c_double]
    L.orc_fft_any.argtypes = [_cp, C.c_long, C.c_int]
    L.orc_fft_pow2.argtypes = [_cp, C.c_long, C.c_int, _cp, C.c_long]
    L.orc_twiddle_table.restype = _cp
    L.orc_twiddle_table.argtypes = [C.c_long]
    _lib = L
    return L


_libc = C.CDLL(None)
_libc.srand.argtypes = [C.c_uint]
_libc.srand48.argtypes = [C.c_long]
_libc.free.argtypes = [C.c_void_p]


def seed(s=17, s48=12345):
    """srand(s) (== srandom) and srand48(s48) on the process-wide libc state."""
    _libc.srand(s)
    _libc.srand48(s48)


def _view(ptr, count, dtype):
    if not ptr or count <= 0:
        return np.empty(0, dtype=dtype)
    nbytes = int(count) * np.dtype(dtype).itemsize
    buf = (C.c_char * nbytes).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=int(count))


class Plan:
    def __init__(self, n, k, version):
        self.L = lib()
        self.pp = self.L.orc_make_plan(n, k, version)
        if not self.pp:
            raise RuntimeError("orc_make_plan returned NULL")
        self.c = self.pp.contents

    def __getattr__(self, name):
        c = object.__getattribute__(self, "c")
        return getattr(c, name)

    @property
    def loops(self):
        return self.c.loops_loc + self.c.loops_est

    def arr(self, name):
        c = self.c
        sizes = {
            "time_loc": (c.w_loc, np.complex128), "freq_loc": (c.n, np.complex128),
            "time_est": (c.w_est, np.complex128), "freq_est": (c.n, np.complex128),
            "a": (self.loops, np.int32), "ai": (self.loops, np.int32),
            "x_sampt": (c.x_samp_size, np.complex128), "x_samp": (c.x_samp_size, np.complex128),
            "mag": (c.x_samp_size, np.float64),
            "J": (self.loops * c.B_thresh, np.int32),
            "score": (c.n, np.int32), "hits": (c.hits_found, np.int32),
            "comb_approved": (c.num_comb, np.int32), "comb_offsets": (c.Comb_loops, np.int32),
            "comb_spec": (c.Comb_loops * c.W_Comb, np.complex128),
            "filtert1": (c.w_g1, np.complex128), "filterf1": (c.n, np.complex128),
            "filtert2": (c.w_g2, np.complex128), "filterf2": (c.n, np.complex128),
            "man_samp": (2 * c.W_Man, np.complex128), "gauss_samp": (2 * c.B_g1, np.complex128),
            "gauss_perm_samp": (2 * c.B_g2, np.complex128),
            "v3_keys": (c.v3_count, np.int32), "v3_vals": (c.v3_count, np.complex128),
            "tw": (c.tw_n // 2, np.complex128),
        }
        cnt, dt = sizes[name]
        return _view(getattr(c, name), cnt, dt)

    def exec(self, x):
        x = np.ascontiguousarray(x, dtype=np.complex128)
        assert x.size == self.c.n_requested
        out = np.empty(self.c.n_requested, dtype=np.complex128)
        self.L.orc_exec(self.pp, x.ctypes.data, out.ctypes.data)
        return out

    def stage(self, name, arr=None):
        fn = getattr(self.L, "orc_" + name)
        if arr is None:
            fn(self.pp)
        else:
            fn(self.pp, arr.ctypes.data)

    def free(self):
        if self.pp:
            self.L.orc_free_plan(self.pp)
            self.pp = None


def generate_input(n, k, seed48=12345):
    """(x, x_f) as src/simulation.cc:95-112 with srand(17), srand48(seed48)."""
    L = lib()
    seed(17, seed48)
    x = np.empty(n, dtype=np.complex128)
    xf = np.empty(n, dtype=np.complex128)
    L.orc_generate_input(n, k, x.ctypes.data, xf.ctypes.data)
    return x, xf


def awgn(x, std_noise):
    L = lib()
    x = np.ascontiguousarray(x, dtype=np.complex128).copy()
    snr = L.orc_awgn(x.ctypes.data, x.size, float(std_noise))
    return x, snr


def fft(x, sign=-1):
    L = lib()
    y = np.ascontiguousarray(x, dtype=np.complex128).copy()
    L.orc_fft_any(y.ctypes.data, y.size, sign)
    return y


def twiddle_table(n):
    L = lib()
    p = L.orc_twiddle_table(n)
    out = _view(p, max(n // 2, 1), np.complex128).copy()
    _libc.free(p)
    return out


def find_largest_indices(samples, num):
    L = lib()
    s = np.ascontiguousarray(samples, dtype=np.float64)
    out = np.empty(num, dtype=np.int32)
    L.orc_find_largest_indices(out.ctypes.data, num, s.ctypes.data, s.size)
    return out
