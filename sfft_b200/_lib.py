"""ctypes loader for the in-tree libsfft.so (C ABI in include/sfft.h).

Fails loudly when the library is missing: there is no Python or CPU fallback.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsfft.so")

SFFTB_MAX_LOOPS = 64
SFFTB_MAX_COMB_LOOPS = 16


class SfftPlan(C.Structure):
    """struct sfft_plan (reference src/sfft.h:44-50)."""
    _fields_ = [("version", C.c_int), ("n", C.c_uint), ("k", C.c_uint), ("data", C.c_void_p)]


class Info(C.Structure):
    _fields_ = [
        ("version", C.c_int), ("n", C.c_int), ("k", C.c_int), ("device", C.c_int),
        ("B_loc", C.c_int), ("B_est", C.c_int), ("B_thresh", C.c_int), ("W_Comb", C.c_int),
        ("Comb_loops", C.c_int),
        ("loops_loc", C.c_int), ("loops_thresh", C.c_int), ("loops_est", C.c_int),
        ("w_loc", C.c_int), ("w_est", C.c_int), ("b_loc", C.c_int), ("b_est", C.c_int),
        ("x_samp_size", C.c_longlong),
        ("B_g1", C.c_int), ("w_g1", C.c_int), ("B_g2", C.c_int), ("w_g2", C.c_int), ("W_Man", C.c_int),
        ("max_hits", C.c_longlong),
        ("gather_samples", C.c_longlong), ("gather_tap_bytes", C.c_longlong),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class Draw(C.Structure):
    _fields_ = [
        ("loops", C.c_int),
        ("a", C.c_int * SFFTB_MAX_LOOPS),
        ("ai", C.c_int * SFFTB_MAX_LOOPS),
        ("comb_offset", C.c_int * SFFTB_MAX_COMB_LOOPS),
        ("v3_a", C.c_int), ("v3_ai", C.c_int), ("v3_b", C.c_int),
        ("v3_init_offset", C.c_int), ("v3_init_G_offset", C.c_int),
    ]


class PeerHandle(C.Structure):
    """sfftb_peer_handle: two CUDA IPC memory handles (plain bytes)."""
    _fields_ = [("spectra", C.c_ubyte * 64), ("flags", C.c_ubyte * 64)]


class Result(C.Structure):
    _fields_ = [("d_loc", C.c_void_p), ("d_val", C.c_void_p), ("d_count", C.c_void_p),
                ("count", C.c_longlong)]


# every symbol include/sfft.h declares: name -> (restype, argtypes)
_vp, _ci, _ll = C.c_void_p, C.c_int, C.c_longlong
_PP = C.POINTER(SfftPlan)
SYMBOLS = {
    "sfft_malloc": (_vp, [C.c_size_t]),
    "sfft_free": (None, [_vp]),
    "sfft_make_plan": (_PP, [_ci, _ci, _ci, _ci]),
    "sfft_free_plan": (None, [_PP]),
    "sfft_exec": (None, [_PP, _vp, _vp]),
    "sfft_exec_many": (None, [_PP, _ci, _vp, _vp]),
    "sfftb_last_error": (C.c_char_p, []),
    "sfftb_device_count": (_ci, []),
    "sfftb_plan_info": (_ci, [_PP, C.POINTER(Info)]),
    "sfftb_set_stream": (_ci, [_PP, _vp]),
    "sfftb_draw_random": (_ci, [_PP, C.POINTER(Draw)]),
    "sfftb_exec_device": (_ci, [_PP, _vp, C.POINTER(Draw), C.POINTER(Result), _ci]),
    "sfftb_exec_many_device": (_ci, [_PP, _ci, _vp, _ll, C.POINTER(Draw), C.POINTER(Result),
                                     C.POINTER(_ll), _ci]),
    "sfftb_densify": (_ci, [_PP, _ci, _vp]),
    "sfftb_synchronize": (_ci, [_PP]),
    "sfftb_save_plan": (_ci, [_PP, C.c_char_p]),
    "sfftb_load_plan": (_PP, [C.c_char_p]),
    "sfftb_fetch_result": (_ll, [_PP, _ci, _vp, _vp, _ll]),
    "sfftb_shard_bucketize": (_ci, [_PP, _vp, C.POINTER(Draw), _ci, _ci]),
    "sfftb_shard_spectra": (_ci, [_PP, C.POINTER(_vp), C.POINTER(_ll)]),
    "sfftb_shard_finish": (_ci, [_PP, _ci, _ci, C.POINTER(Result), _ci]),
    "sfftb_shard_loops": (_ci, [_PP, _ci, _ci, C.POINTER(_ci), C.POINTER(_ci)]),
    "sfftb_shard_export": (_ci, [_PP, C.POINTER(PeerHandle)]),
    "sfftb_shard_attach": (_ci, [_PP, _ci, _ci, C.POINTER(PeerHandle)]),
    "sfftb_shard_detach": (_ci, [_PP]),
    "sfftb_shard_exec": (_ci, [_PP, _vp, C.POINTER(Draw), C.POINTER(Result), _ci]),
    "sfftb_shard_status": (_ci, [_PP, C.POINTER(_ll), C.POINTER(_ll)]),
    "sfftb_shard_slice": (_ci, [_PP, _ci, _ci, C.POINTER(_ll), C.POINTER(_ll)]),
    "sfftb_filter_sizes": (_ci, [_PP, _ci, C.POINTER(_ci), C.POINTER(_ci)]),
    "sfftb_get_filter": (_ci, [_PP, _ci, _vp, _vp]),
    "sfftb_set_filter": (_ci, [_PP, _ci, _vp, _vp]),
    "sfftb_debug_fetch": (_ll, [_PP, C.c_char_p, _vp, C.c_size_t]),
    "sfftb_debug_fft": (_ci, [_vp, _vp, _ci, _ci, _ci, _ci]),
    "sfftb_debug_select": (_ci, [_vp, _ci, _ci, _ci, _vp]),
    "sfftb_debug_dft_any": (_ci, [_vp, _vp, _ci]),
    "sfftb_debug_div_check": (_ll, [C.c_ulonglong, _ll]),
    "sfftb_enable_stage_timing": (_ci, [_PP, _ci]),
    "sfftb_stage_times": (_ci, [_PP, _vp, _vp, _ci]),
    "sfftb_launch_count": (_ll, []),
}

_lib = None


def load():
    """Load libsfft.so; raises if it has not been built (python -m sfft_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OSError(
            f"{LIB_PATH} is missing: build it with `python -m sfft_b200.build` "
            "(nvcc, sm_100a).  There is no fallback implementation.")
    L = C.CDLL(LIB_PATH, mode=C.RTLD_LOCAL)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(L, name)      # AttributeError if the export is missing
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def last_error():
    return load().sfftb_last_error().decode(errors="replace")
