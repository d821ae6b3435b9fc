"""CPU-only checks of the boundary: the C-ABI library loads and exports every symbol
include/sfft.h declares, fails loudly without a GPU, and the Python binding keeps
the reference's validation behaviour (python/sfft/sfft.py:52-67)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from util import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "sfft.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sfftb?_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from sfft_b200 import _lib
    L = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 28
    for name in declared:
        assert hasattr(L, name), f"libsfft.so does not export {name}"
    assert set(declared) == set(_lib.SYMBOLS), "ctypes table and header disagree"


def test_plan_struct_layout_matches_reference_header():
    from sfft_b200 import _lib
    # struct sfft_plan {sfft_version version; unsigned n; unsigned k; void *data;} (src/sfft.h:44-50)
    assert C.sizeof(_lib.SfftPlan) == 24
    assert _lib.SfftPlan.data.offset == 16


def test_no_cpu_fallback_without_a_device():
    from sfft_b200 import _lib
    L = _lib.load()
    if L.sfftb_device_count() > 0:
        pytest.skip("a GPU is present")
    assert not L.sfft_make_plan(16384, 50, 0, 64)
    assert b"no CUDA device" in L.sfftb_last_error()
    x = np.zeros(8, dtype=np.complex128)
    assert L.sfftb_debug_fft(x.ctypes.data, x.ctypes.data, 3, 1, -1, 1) == -1
    # sfft_malloc still hands out usable, aligned host memory
    p = L.sfft_malloc(1024)
    assert p and p % 16 == 0
    L.sfft_free(p)


def test_unknown_version_returns_null():
    from sfft_b200 import _lib
    L = _lib.load()
    assert not L.sfft_make_plan(16384, 50, 7, 64)      # src/sfft.cc:87-88


def test_binding_validation_matches_reference():
    import sfft_b200.sfft as m
    assert m.FFTW_MEASURE == 0 and m.FFTW_ESTIMATE == 64
    assert {"size": 4194304, "sparsity": 2500} in m.V1_V2_INPUT_PARAMETERS
    assert len(m.V1_V2_INPUT_PARAMETERS) == 20
    with pytest.raises(TypeError):
        m.sfft(16384.0, 50, 1)
    with pytest.raises(TypeError):
        m.sfft(16384, "50", 1)
    with pytest.raises(ValueError):
        m.sfft(16384, 50, 4)
    with pytest.raises(ValueError):
        m.sfft(16384, 51, 1)           # not in the whitelist
    with pytest.raises(ValueError):
        m.sfft(16384, 50, 1, optimization=3)


def test_product_does_not_touch_the_oracle():
    """The shipped path must not include, import, link, load or execute anything under
    oracle/ (comments may cite it: the oracle pins definitions the product also follows)."""
    pkg = os.path.join(ROOT, "sfft_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath.split(os.sep) or "__pycache__" in dirpath:
            continue
        for f in files:
            path = os.path.join(dirpath, f)
            if f.endswith((".cu", ".cuh", ".c", ".h", ".inc")):
                for line in open(path, errors="replace"):
                    code = line.split("//")[0]
                    if code.lstrip().startswith(("*", "/*")):
                        continue
                    assert not re.search(r'#\s*include.*oracle', code), (path, line)
                    assert not re.search(r'"[^"]*oracle[^"]*"', code), (path, line)
            elif f.endswith(".py"):
                text = open(path, errors="replace").read()
                assert not re.search(r"^\s*(import|from)\s+oracle", text, flags=re.M), path
                assert not re.search(r"""["'][^"'\n]*oracle[/_.][^"'\n]*["']""", text), path
    out = os.popen(f"ldd {os.path.join(pkg, 'libsfft.so')}").read()
    assert "oracle" not in out


def test_plan_cache_refuses_missing_and_foreign_files(tmp_path):
    """sfftb_load_plan fails loudly (NULL + message) before touching any device: no CUDA needed."""
    from sfft_b200 import _lib
    L = _lib.load()
    assert not L.sfftb_load_plan(str(tmp_path / "nope.plan").encode())
    assert "cannot open" in _lib.last_error()
    bad = tmp_path / "foreign.plan"
    bad.write_bytes(b"SFFTBPL0" + bytes(64))
    assert not L.sfftb_load_plan(str(bad).encode())
    assert "not a plan file" in _lib.last_error()


def test_generated_median_networks_are_current(tmp_path):
    """sfft_b200/csrc/median_networks.inc is what tools/gen_median_networks.py generates --
    and generating it re-verifies every network (0/1 principle, exhaustively up to L = 20)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = tmp_path / "median_networks.inc"
    subprocess.check_call([sys.executable, os.path.join(root, "tools", "gen_median_networks.py"), str(out)],
                          stdout=subprocess.DEVNULL)
    assert out.read_text() == open(os.path.join(root, "sfft_b200", "csrc", "median_networks.inc")).read()
