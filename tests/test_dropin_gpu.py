"""The reference's own test programs (test_cfft.cpp, test_rfft.cpp), compiled UNCHANGED against include/ and
linked to libcl_fft.so / libb200fft.so (tests/test_capi.py builds them where /root/reference exists and
parks the binaries in oracle/_ref/, which travels to the GPU box), must print the known answers."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
BIN = os.path.join(ROOT, "oracle", "_ref")


def _vector(line):
    return line[line.index("[") + 1: line.rindex("]")]


def _run(prog):
    exe = os.path.join(BIN, prog)
    if not os.path.exists(exe):
        pytest.skip(f"{exe} not built (needs /root/reference at build time)")
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = os.path.join(ROOT, "opencl_fft_b200", "lib") + ":" + env.get("LD_LIBRARY_PATH", "")
    res = subprocess.run([exe], capture_output=True, text=True, env=env, timeout=120)
    assert res.returncode == 0, res.stdout + res.stderr
    return res.stdout.splitlines()


def test_reference_test_cfft_prints_known_answer():
    out = _run("test_cfft_b200")
    assert out[0].startswith("using device 0:") and "B200" in out[0]
    spec = [complex(float(a), float(b)) for a, b in re.findall(r"\(([-\d.e]+),([-\d.e]+)\)", _vector(out[2]))]
    want = np.zeros(16, complex)
    want[1], want[15] = -0.5j, 0.5j  # 16-point sine, forward scaled by 1/N
    assert np.abs(np.array(spec) - want).max() < 2e-3  # the program prints 3 decimals
    vin = np.array([float(v) for v in _vector(out[1]).split(",")])
    vout = np.array([float(v) for v in _vector(out[3]).split(",")])
    assert np.abs(vin - vout).max() < 2e-3


def test_reference_test_rfft_prints_known_answer():
    out = _run("test_rfft_b200")
    spec = [complex(float(a), float(b)) for a, b in re.findall(r"\(([-\d.e]+),([-\d.e]+)\)", _vector(out[2]))]
    want = np.zeros(8, complex)
    want[0], want[1] = 0.5 + 0.5j, -1j  # dc + fundamental + nyquist
    assert np.abs(np.array(spec) - want).max() < 2e-3
    vin = np.array([float(v) for v in _vector(out[1]).split(",")])
    vout = np.array([float(v) for v in _vector(out[3]).split(",")])
    assert np.abs(vin - vout).max() < 2e-3
