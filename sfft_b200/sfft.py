"""Python binding with the reference's entry points (reference python/sfft/sfft.py).

Same module-level constants, same class surface `sfft(length, sparsity, version,
optimization)` / `.execute(a)`, same TypeError/ValueError validation
(python/sfft/sfft.py:52-67), ported to Python 3 / NumPy 2 and bound to the
B200-native libsfft.so.  On top of `execute` (host ndarray in, dense host ndarray
out, as in the reference) the class exposes the device-resident path:
`execute_device` / `execute_many_device` (CUDA tensor in, sparse result left on the
device; `result()` copies it to the host, `result_device()` views it as CUDA tensors).
"""
import ctypes as C

import numpy as np

from . import _lib

# flags copied from fftw3.h (python/sfft/sfft.py:7-8)
FFTW_MEASURE = 0
FFTW_ESTIMATE = 1 << 6
PLAN_TUNED_BY_K = 1 << 24      # include/sfft.h SFFTB_PLAN_TUNED_BY_K

# python/sfft/sfft.py:10-31
V1_V2_INPUT_PARAMETERS = [
    {"size": 8192, "sparsity": 50},
    {"size": 16384, "sparsity": 50},
    {"size": 32768, "sparsity": 50},
    {"size": 65536, "sparsity": 50},
    {"size": 131072, "sparsity": 50},
    {"size": 262144, "sparsity": 50},
    {"size": 524288, "sparsity": 50},
    {"size": 1048576, "sparsity": 50},
    {"size": 2097152, "sparsity": 50},
    {"size": 4194304, "sparsity": 50},
    {"size": 8388608, "sparsity": 50},
    {"size": 16777216, "sparsity": 50},
    {"size": 4194304, "sparsity": 50},
    {"size": 4194304, "sparsity": 100},
    {"size": 4194304, "sparsity": 200},
    {"size": 4194304, "sparsity": 500},
    {"size": 4194304, "sparsity": 1000},
    {"size": 4194304, "sparsity": 2000},
    {"size": 4194304, "sparsity": 2500},
    {"size": 4194304, "sparsity": 4000},
]


class sfft:
    """Sparse FFT plan.  `strict_parameters=True` (default) keeps the reference's
    whitelist of (n, k) for versions 1 and 2 (python/sfft/sfft.py:64-65); the C API
    itself accepts any power-of-two n, so pass False to reach e.g. n=2^27."""

    def __init__(self, length=16384, sparsity=50, version=1, optimization=FFTW_ESTIMATE,
                 strict_parameters=True, tuned_by_k=False):
        if not isinstance(length, (int, np.integer)) or isinstance(length, bool):
            raise TypeError("length is not an integer")
        if not isinstance(sparsity, (int, np.integer)) or isinstance(sparsity, bool):
            raise TypeError("sparsity is not an integer")
        if not isinstance(version, (int, np.integer)) or isinstance(version, bool):
            raise TypeError("version is not an integer")
        if not isinstance(optimization, (int, np.integer)) or isinstance(optimization, bool):
            raise TypeError("optimization is not an integer")

        if version not in [1, 2, 3]:
            raise ValueError("sFFT version %d is not valid.  Try 1, 2, or 3." % (version))
        if (strict_parameters and version in [1, 2]
                and {"size": length, "sparsity": sparsity} not in V1_V2_INPUT_PARAMETERS):
            raise ValueError(
                "n = %d and k = %d is not a valid input parameter combination for sFFT version %d."
                % (length, sparsity, version))
        if optimization not in [FFTW_MEASURE, FFTW_ESTIMATE]:
            raise ValueError("FFTW optimization %d is not valid." % (optimization))

        self.length = int(length)
        self.sparsity = int(sparsity)
        self.version = int(version)
        self.optimization = int(optimization)

        self._L = _lib.load()
        # python/sfft/sfft.py:74: the binding always plans with FFTW_ESTIMATE
        # tuned_by_k: opt in to the by-k parameter lookup the reference's table was written
        # for (include/sfft.h SFFTB_PLAN_TUNED_BY_K); off = the reference's behaviour
        flags = FFTW_ESTIMATE | (PLAN_TUNED_BY_K if tuned_by_k else 0)
        self.sfft_plan = self._L.sfft_make_plan(self.length, self.sparsity, self.version - 1, flags)
        if not self.sfft_plan:
            raise RuntimeError("sfft_make_plan failed: " + _lib.last_error())

    # -- plan cache ------------------------------------------------------------
    def save(self, path):
        """Write (n, k, version, flags) and both filters to `path` (sfftb_save_plan)."""
        if self._L.sfftb_save_plan(self.sfft_plan, str(path).encode()):
            raise RuntimeError(_lib.last_error())

    @classmethod
    def load(cls, path):
        """Re-create a saved plan without running the filter builder (sfftb_load_plan)."""
        L = _lib.load()
        plan = L.sfftb_load_plan(str(path).encode())
        if not plan:
            raise RuntimeError("sfftb_load_plan failed: " + _lib.last_error())
        self = cls.__new__(cls)
        self._L = L
        self.sfft_plan = plan
        self.length = int(plan.contents.n)
        self.sparsity = int(plan.contents.k)
        self.version = int(plan.contents.version) + 1
        self.optimization = FFTW_ESTIMATE
        return self

    # -- reference surface ---------------------------------------------------
    def execute(self, a):
        """Dense spectrum of `a` (python/sfft/sfft.py:76-82)."""
        a = np.asanyarray(a)
        a = np.require(a, np.complex128, ["C", "ALIGNED"])
        if a.ndim != 1 or a.size != self.length:
            raise ValueError("input must be a 1-d array of %d complex samples" % self.length)
        b = np.empty_like(a)
        self._L.sfft_exec(self.sfft_plan, a.ctypes.data, b.ctypes.data)
        return b

    # -- device-resident extension --------------------------------------------
    def info(self):
        inf = _lib.Info()
        if self._L.sfftb_plan_info(self.sfft_plan, C.byref(inf)):
            raise RuntimeError(_lib.last_error())
        return inf.as_dict()

    def draw(self):
        """One transform's random draw from libc random()/drand48()."""
        d = _lib.Draw()
        if self._L.sfftb_draw_random(self.sfft_plan, C.byref(d)):
            raise RuntimeError(_lib.last_error())
        return d

    def set_stream(self, cuda_stream):
        if self._L.sfftb_set_stream(self.sfft_plan, C.c_void_p(cuda_stream or 0)):
            raise RuntimeError(_lib.last_error())

    def execute_device(self, x, draw=None, sync=True):
        """x: CUDA tensor complex128[n] (torch).  Returns the number of recovered
        coefficients (or None when sync=False); fetch them with `result()`."""
        res = _lib.Result()
        rc = self._L.sfftb_exec_device(self.sfft_plan, C.c_void_p(x.data_ptr()),
                                       C.byref(draw) if draw is not None else None,
                                       C.byref(res), 1 if sync else 0)
        if rc:
            raise RuntimeError(_lib.last_error())
        self._last = res
        return int(res.count) if sync else None

    def execute_many_device(self, x, draws=None, sync=True):
        """x: CUDA tensor complex128[num, n], contiguous."""
        num = int(x.shape[0])
        arr = None
        if draws is not None:
            arr = (_lib.Draw * num)(*draws)
        counts = (C.c_longlong * num)()
        res = _lib.Result()
        rc = self._L.sfftb_exec_many_device(self.sfft_plan, num, C.c_void_p(x.data_ptr()),
                                            int(x.stride(0)), arr, C.byref(res), counts,
                                            1 if sync else 0)
        if rc:
            raise RuntimeError(_lib.last_error())
        self._last = res
        return list(counts) if sync else None

    def result(self, which=0):
        """(locations int32[count], values complex128[count]) of signal `which`."""
        cap = self.info()["max_hits"]
        cnt = self._L.sfftb_fetch_result(self.sfft_plan, which, None, None, 0)
        if cnt < 0:
            raise RuntimeError(_lib.last_error())
        loc = np.empty(cnt, dtype=np.int32)
        val = np.empty(cnt, dtype=np.complex128)
        if cnt:
            got = self._L.sfftb_fetch_result(self.sfft_plan, which, loc.ctypes.data, val.ctypes.data,
                                             min(cnt, cap))
            if got < 0:
                raise RuntimeError(_lib.last_error())
        return loc, val

    def result_device(self, which=0):
        """The same list as `result()`, left on the device: (int32[count], complex128[count])
        torch tensors viewing the plan's own buffers -- valid until the next transform."""
        import torch
        cnt = self._L.sfftb_fetch_result(self.sfft_plan, which, None, None, 0)
        if cnt < 0:
            raise RuntimeError(_lib.last_error())
        cap = self.info()["max_hits"]
        res = self._last
        dev = torch.device("cuda", self.info()["device"])

        class _View:
            def __init__(self, ptr, count, typestr):
                self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": typestr,
                                                 "data": (int(ptr), False), "version": 2}
        if cnt == 0:
            return (torch.empty(0, dtype=torch.int32, device=dev),
                    torch.empty(0, dtype=torch.complex128, device=dev))
        loc = torch.as_tensor(_View(res.d_loc + 4 * which * cap, cnt, "<i4"), device=dev)
        val = torch.as_tensor(_View(res.d_val + 16 * which * cap, cnt, "<c16"), device=dev)
        return loc, val

    def densify(self, out, which=0, sync=True):
        """Zero `out` (CUDA tensor complex128[n]) and scatter the sparse result.  The work is
        queued on the plan's stream; sync=True waits for it (needed unless the caller's own
        work runs on that same stream, see set_stream)."""
        if self._L.sfftb_densify(self.sfft_plan, which, C.c_void_p(out.data_ptr())):
            raise RuntimeError(_lib.last_error())
        if sync:
            self.synchronize()
        return out

    def synchronize(self):
        if self._L.sfftb_synchronize(self.sfft_plan):
            raise RuntimeError(_lib.last_error())

    def debug_fetch(self, what, dtype, count):
        buf = np.empty(count, dtype=dtype)
        got = self._L.sfftb_debug_fetch(self.sfft_plan, what.encode(), buf.ctypes.data, buf.nbytes)
        if got < 0:
            raise RuntimeError(_lib.last_error())
        return buf[: got // buf.itemsize]

    def get_filter(self, which):
        w, fl = C.c_int(), C.c_int()
        if self._L.sfftb_filter_sizes(self.sfft_plan, which, C.byref(w), C.byref(fl)):
            raise RuntimeError(_lib.last_error())
        t = np.empty(w.value, dtype=np.complex128)
        f = np.empty(fl.value, dtype=np.complex128)
        if self._L.sfftb_get_filter(self.sfft_plan, which, t.ctypes.data, f.ctypes.data):
            raise RuntimeError(_lib.last_error())
        return t, f

    def set_filter(self, which, time=None, freq_window=None):
        t = None if time is None else np.ascontiguousarray(time, dtype=np.complex128)
        f = None if freq_window is None else np.ascontiguousarray(freq_window, dtype=np.complex128)
        if self._L.sfftb_set_filter(self.sfft_plan, which,
                                    None if t is None else t.ctypes.data,
                                    None if f is None else f.ctypes.data):
            raise RuntimeError(_lib.last_error())

    def stage_timing(self, on=True):
        self._L.sfftb_enable_stage_timing(self.sfft_plan, 1 if on else 0)

    def stage_times(self):
        ms = (C.c_float * 16)()
        names = (C.c_char_p * 16)()
        cnt = self._L.sfftb_stage_times(self.sfft_plan, ms, names, 16)
        return {names[i].decode(): float(ms[i]) for i in range(max(cnt, 0))}

    def close(self):
        if getattr(self, "sfft_plan", None):
            self._L.sfft_free_plan(self.sfft_plan)
            self.sfft_plan = None

    def __del__(self):
        # python/sfft/sfft.py:84-85
        try:
            self.close()
        except Exception:
            pass
