"""oracle/ref.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes window onto oracle/_ref/libsfft_ref_{parity,fast}.so: the UNMODIFIED
reference sources (/root/reference/src) compiled by oracle/Makefile against the
FFTW-API shim.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs may import this module.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_cache = {}


def available(kind="parity"):
    return os.path.exists(os.path.join(_HERE, "_ref", f"libsfft_ref_{kind}.so"))


def mkl_provider():
    """The shared object that exports MKL's DFTI in this image: PyTorch's libtorch_cpu.so."""
    import importlib.util
    spec = importlib.util.find_spec("torch")
    if not spec or not spec.origin:
        return None
    p = os.path.join(os.path.dirname(spec.origin), "lib", "libtorch_cpu.so")
    return p if os.path.exists(p) else None


def lib(kind="parity"):
    """kind: "parity" (IEEE flags, oracle FFT), "fast" (reference flags, oracle FFT), "mkl"
    (reference flags over MKL DFTI -- timing baseline only)."""
    if kind in _cache:
        return _cache[kind]
    if kind == "mkl":
        prov = mkl_provider()
        if not prov:
            raise OSError("no libtorch_cpu.so to take MKL DFTI from")
        os.environ["SFFT_REF_MKL_LIB"] = prov
    path = os.path.join(_HERE, "_ref", f"libsfft_ref_{kind}.so")
    L = C.CDLL(path, mode=C.RTLD_LOCAL)
    vp, ci, cd = C.c_void_p, C.c_int, C.c_double
    L.ref_set_threads.argtypes = [ci]
    L.ref_get_max_threads.restype = ci
    L.ref_seed.argtypes = [C.c_uint, C.c_long]
    L.ref_make_plan.restype = vp
    L.ref_make_plan.argtypes = [ci, ci, ci]
    L.ref_exec.argtypes = [vp, vp, vp]
    L.ref_exec_many.argtypes = [vp, ci, vp, vp]
    L.ref_free_plan.argtypes = [vp]
    L.ref_v12_params.argtypes = [vp, vp]
    L.ref_v3_params.argtypes = [vp, vp]
    for nm in ("ref_v12_filter_time", "ref_v12_filter_freq"):
        getattr(L, nm).restype = vp
        getattr(L, nm).argtypes = [vp, ci]
    for nm in ("ref_v12_x_samp", "ref_v12_x_sampt", "ref_v12_score", "ref_v12_hits",
               "ref_v12_permute", "ref_v12_comb_approved", "ref_v12_J"):
        getattr(L, nm).restype = vp
        getattr(L, nm).argtypes = [vp]
    L.ref_v3_filter.restype = vp
    L.ref_v3_filter.argtypes = [vp, ci]
    L.ref_v3_samples.restype = vp
    L.ref_v3_samples.argtypes = [vp, ci]
    L.ref_floor_to_pow2.restype = ci
    L.ref_floor_to_pow2.argtypes = [cd]
    L.ref_mod_inverse.restype = ci
    L.ref_mod_inverse.argtypes = [ci, ci]
    L.ref_fftw_dft.argtypes = [vp, ci, vp, ci]
    L.ref_find_largest_indices.argtypes = [vp, ci, vp, ci, vp]
    L.ref_awgn.restype = cd
    L.ref_awgn.argtypes = [vp, ci, cd]
    L.ref_generate_input.argtypes = [ci, ci, vp, vp]
    _cache[kind] = L
    return L


def _view(ptr, count, dtype):
    if not ptr:
        return None
    nbytes = count * np.dtype(dtype).itemsize
    buf = (C.c_char * nbytes).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=count)


V12_KEYS = ("B_loc", "B_est", "B_thresh", "W_Comb", "Comb_loops", "loops_loc",
            "loops_thresh", "loops_est", "w_loc", "w_est", "x_samp_size", "threads")
V3_KEYS = ("B_g1", "w_g1", "B_g2", "w_g2", "W_Man", "Gauss_loops", "Gauss2_loops", "Man_loops")


class RefPlan:
    """sfft_make_plan / sfft_exec of the compiled reference (src/sfft.h:158-167)."""

    def __init__(self, n, k, version, kind="parity", threads=1):
        self.L = lib(kind)
        self.L.ref_set_threads(threads)
        self.n_req, self.k, self.version = n, k, version
        self.p = self.L.ref_make_plan(n, k, version - 1)
        if not self.p:
            raise RuntimeError("reference sfft_make_plan returned NULL")
        self.n = 1 << (int(n).bit_length() - 1)
        out = (C.c_int * 12)()
        if version in (1, 2):
            self.L.ref_v12_params(self.p, out)
            self.params = dict(zip(V12_KEYS, list(out)))
        else:
            self.L.ref_v3_params(self.p, out)
            self.params = dict(zip(V3_KEYS, list(out)[:8]))

    def seed(self, s=17, s48=12345):
        self.L.ref_seed(s, s48)

    def exec(self, x):
        x = np.ascontiguousarray(x, dtype=np.complex128)
        assert x.size == self.n_req
        out = np.empty(self.n_req, dtype=np.complex128)
        self.L.ref_exec(self.p, x.ctypes.data, out.ctypes.data)
        return out

    def exec_many(self, xs):
        xs = [np.ascontiguousarray(x, dtype=np.complex128) for x in xs]
        outs = [np.empty(self.n_req, dtype=np.complex128) for _ in xs]
        ip = (C.c_void_p * len(xs))(*[x.ctypes.data for x in xs])
        op = (C.c_void_p * len(xs))(*[o.ctypes.data for o in outs])
        self.L.ref_exec_many(self.p, len(xs), ip, op)
        return outs

    # --- v1/v2 internals -------------------------------------------------
    def filter_time(self, est=False):
        w = self.params["w_est" if est else "w_loc"]
        return _view(self.L.ref_v12_filter_time(self.p, int(est)), w, np.complex128)

    def filter_freq(self, est=False):
        return _view(self.L.ref_v12_filter_freq(self.p, int(est)), self.n, np.complex128)

    def x_samp(self):
        return _view(self.L.ref_v12_x_samp(self.p), self.params["x_samp_size"], np.complex128)

    def x_sampt(self):
        return _view(self.L.ref_v12_x_sampt(self.p), self.params["x_samp_size"], np.complex128)

    def score(self):
        return _view(self.L.ref_v12_score(self.p), self.n, np.int32)

    def hits(self):
        return _view(self.L.ref_v12_hits(self.p), self.n, np.int32)

    def permute(self):
        loops = self.params["loops_loc"] + self.params["loops_est"]
        return _view(self.L.ref_v12_permute(self.p), loops, np.int32)

    def comb_approved(self):
        cnt = self.params["Comb_loops"] * self.params["B_thresh"]
        return _view(self.L.ref_v12_comb_approved(self.p), cnt, np.int32)

    # --- v3 internals ----------------------------------------------------
    def v3_filter(self, which):
        sizes = {0: self.params["w_g1"], 1: self.n, 2: self.params["w_g2"], 3: self.n}
        return _view(self.L.ref_v3_filter(self.p, which), sizes[which], np.complex128)

    def v3_samples(self, which):
        sizes = {0: 2 * self.params["W_Man"], 1: 2 * self.params["B_g1"], 2: 2 * self.params["B_g2"]}
        return _view(self.L.ref_v3_samples(self.p, which), sizes[which], np.complex128)

    def free(self):
        if self.p:
            self.L.ref_free_plan(self.p)
            self.p = None


def generate_input(n, k, seed48=12345, kind="parity"):
    """k unit spikes at floor(drand48()*n); x = unnormalised inverse DFT
    (src/simulation.cc:95-112 with a fixed seed). Returns (x, x_f)."""
    L = lib(kind)
    L.ref_seed(17, seed48)
    x = np.empty(n, dtype=np.complex128)
    xf = np.empty(n, dtype=np.complex128)
    L.ref_generate_input(n, k, x.ctypes.data, xf.ctypes.data)
    return x, xf


def fftw_dft(x, backwards=False, kind="parity"):
    L = lib(kind)
    x = np.ascontiguousarray(x, dtype=np.complex128)
    out = np.empty_like(x)
    L.ref_fftw_dft(out.ctypes.data, x.size, x.ctypes.data, int(backwards))
    return out
