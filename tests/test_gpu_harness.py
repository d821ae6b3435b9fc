"""The reference's acceptance test (sfft-verification, src/verification.cc:26-62) and its
timing drivers, re-created over this library in tools/sfft_harness.cc."""
import os
import subprocess

import pytest

from util import ROOT

pytestmark = pytest.mark.gpu
TOOLS = os.path.join(ROOT, "tools")


# the last two are BASELINE configs 2 and 4 (exact variant) at full size: the driver synthesises its
# input with an inverse FFT on the device, so they take seconds
@pytest.mark.parametrize("version,n,k", [(1, 16384, 50), (2, 16384, 50), (3, 16384, 50), (1, 1 << 20, 100),
                                         (3, 1 << 22, 1000), (2, 1 << 24, 1000), (1, 1 << 27, 500)])
def test_verification_driver_says_ok(version, n, k):
    out = subprocess.run([os.path.join(TOOLS, "sfft-verification"), "-n", str(n), "-k", str(k), "-v", str(version)],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout + out.stderr


def test_timing_drivers_run():
    out = subprocess.run([os.path.join(TOOLS, "sfft-timing"), "-n", "65536", "-k", "50", "-r", "3"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.startswith("TIME"), out.stdout + out.stderr
    for extra in ([], ["-s"]):
        out = subprocess.run([os.path.join(TOOLS, "sfft-timing_many"), "-n", "65536", "-k", "50", "-i", "8"] + extra,
                             capture_output=True, text=True, timeout=300)
        assert out.returncode == 0 and out.stdout.startswith("TIME"), out.stdout + out.stderr
