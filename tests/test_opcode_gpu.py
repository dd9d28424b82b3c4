"""SURVEY 8f rank 1, the Csound opcode layer: the reference's csound/opcode.cpp -- compiled UNCHANGED, hosted by the
stand-in Csound framework in tests/csound_host (Csound 7 is not installed in this image) -- runs its `clconv` and
`cltvconv` opcodes init()/aperf() cycle by cycle, once over this repository's class library on the B200
(oracle/_ref/opcode_host_b200) and once over the reference's own implementation on the host-CPU OpenCL runtime
(oracle/_ref/opcode_host_ref). Same audio in, same audio out within the parity tolerance: ksmps -> partition buffering
with one partition of latency (opcode.cpp:241-249), IR x 0dbfs (189-191), parts == 1 -> Cldconv(size, ksmps)
(184-187), input / 0dbfs and output x 0dbfs with the freeze flags of cltvconv (317-340). Both binaries are built by
__graft_entry__.build() where /root/reference exists and travel to the GPU box in oracle/_ref/."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, TOL, rel_l2

pytestmark = pytest.mark.gpu
BIN = os.path.join(ROOT, "oracle", "_ref")


def _run(which, tmp_path, args):
    exe = os.path.join(BIN, "opcode_host_" + which)
    if not os.path.exists(exe):
        pytest.skip(f"{exe} not built (needs /root/reference at build time)")
    res = subprocess.run([exe, *map(str, args), str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout + res.stderr
    return np.fromfile(tmp_path / "out.f32", np.float32)


def _both(tmp_path, args):
    return _run("b200", tmp_path, args), _run("ref", tmp_path, args)


@pytest.mark.parametrize("parts,ksmps,taps,dbfs", [(64, 16, 1000, 32768.0), (512, 64, 4096, 1.0), (256, 256, 2048, 32768.0)])
def test_clconv_partitioned(tmp_path, parts, ksmps, taps, dbfs):
    rng = np.random.default_rng(parts + ksmps)
    ncycles = (3 * (taps // parts) + 4) * (parts // ksmps)
    (rng.standard_normal(taps) * 0.05).astype(np.float32).tofile(tmp_path / "ir.f32")
    (rng.uniform(-1, 1, ksmps * ncycles) * dbfs).astype(np.float32).tofile(tmp_path / "in.f32")
    got, want = _both(tmp_path, ["conv", parts, ksmps, ncycles, dbfs])
    assert np.all(got[:parts] == 0) and np.all(want[:parts] == 0)  # one partition of latency (Q7)
    assert np.abs(want).max() > 0
    assert rel_l2(got, want) < TOL


def test_clconv_direct_when_parts_is_one(tmp_path):
    """parts == 1 selects Cldconv(size, ksmps): sample-by-sample direct convolution, no buffering latency beyond the
    reference's one-sample delay."""
    rng = np.random.default_rng(5)
    ksmps, taps, ncycles, dbfs = 32, 256, 40, 32768.0
    (rng.standard_normal(taps) / 16).astype(np.float32).tofile(tmp_path / "ir.f32")
    (rng.uniform(-1, 1, ksmps * ncycles) * dbfs).astype(np.float32).tofile(tmp_path / "in.f32")
    got, want = _both(tmp_path, ["conv", 1, ksmps, ncycles, dbfs])
    assert got[0] == 0 and np.abs(got[1:ksmps]).max() > 0
    assert rel_l2(got, want) < TOL


@pytest.mark.parametrize("parts,ksmps,size", [(64, 16, 640), (1, 32, 256)])
def test_cltvconv_with_freeze(tmp_path, parts, ksmps, size):
    rng = np.random.default_rng(parts)
    dbfs = 32768.0
    ncycles = 6 * max(1, size // max(parts, ksmps)) * max(1, parts // ksmps)
    (rng.uniform(-1, 1, ksmps * ncycles) * dbfs).astype(np.float32).tofile(tmp_path / "in.f32")
    (rng.uniform(-1, 1, ksmps * ncycles) * dbfs * 0.1).astype(np.float32).tofile(tmp_path / "in2.f32")
    frz = np.ones(ncycles, np.float32)
    frz[ncycles // 3: ncycles // 2] = 0  # hold the buffers for a while, then resume
    frz.tofile(tmp_path / "frz.f32")
    got, want = _both(tmp_path, ["tvconv", parts, ksmps, ncycles, dbfs, size])
    assert np.abs(want).max() > 0
    assert rel_l2(got, want) < TOL
