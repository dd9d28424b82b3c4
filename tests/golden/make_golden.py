"""Generate tests/golden/reference_vectors.npz by running the UNMODIFIED reference.

The reference C++ sources in /root/reference are compiled (oracle/Makefile, `make ref`) against
oracle/minicl, a host-CPU OpenCL runtime, and driven through ctypes (oracle.ref()). This script can
only run where /root/reference exists (the build container); the vectors it writes are committed so
that the parity tests on the GPU box -- where the reference is absent -- still compare against outputs
of the real reference, not only against the restatement.

    python tests/golden/make_golden.py

Inputs follow SURVEY.md section 8(d) (numpy default_rng(seed), float32).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import oracle  # noqa: E402


def decay_ir(rng, n, full):
    h = rng.standard_normal(n) * np.exp(-6.9078 * np.arange(n) / full)
    return (h / np.linalg.norm(h)).astype(np.float32)


def main():
    oracle.build()
    R = oracle.ref()
    assert R is not None, "oracle/_ref/libclfft_ref.so missing: /root/reference not available?"
    out = {}

    # --- known answers implied by the reference's own test programs -----------------------------
    N = 16
    sig = np.sin(np.arange(N) * 2 * oracle_pi() / N)
    x = np.zeros(N, np.complex64)
    x.real = sig.astype(np.float32)  # test_cfft.cpp:54-56 (double sin -> float real part)
    out["kat_cfft_in"] = x
    out["kat_cfft_spec"] = R.cfft(x, True)
    out["kat_cfft_back"] = R.cfft(out["kat_cfft_spec"], False)
    i = np.arange(N)
    r = (0.5 + np.sin(i * 2 * oracle_pi() / N) + 0.5 * np.cos(i * oracle_pi())).astype(np.float32)  # test_rfft.cpp:54-57
    out["kat_rfft_in"] = r
    out["kat_rfft_spec"] = R.rfft_fwd(r, out_of_place=True)
    out["kat_rfft_back"] = R.rfft_inv(out["kat_rfft_spec"], out_of_place=True)

    # --- S1: cfft 1024 (BASELINE config 1) ------------------------------------------------------
    rng = np.random.default_rng(1001)
    x = (rng.uniform(-1, 1, 1024) + 1j * rng.uniform(-1, 1, 1024)).astype(np.complex64)
    out["cfft1024_in"] = x
    out["cfft1024_fwd"] = R.cfft(x, True)
    out["cfft1024_inv"] = R.cfft(x, False)

    # --- S2: rfft 4096 round trip (config 2) ----------------------------------------------------
    rng = np.random.default_rng(1002)
    r = rng.uniform(-1, 1, 4096).astype(np.float32)
    out["rfft4096_in"] = r
    out["rfft4096_fwd"] = R.rfft_fwd(r)
    out["rfft4096_back"] = R.rfft_inv(out["rfft4096_fwd"])

    # --- a large transform: 65536-point real (config 5a, one channel) ----------------------------
    rng = np.random.default_rng(6000)
    r = rng.uniform(-1, 1, 65536).astype(np.float32)
    out["rfft65536_in"] = r
    out["rfft65536_fwd"] = R.rfft_fwd(r)

    # --- S3: partitioned convolution 96000 taps / 512 (config 3), 12 blocks -----------------------
    rng_h, rng_x = np.random.default_rng(2000), np.random.default_rng(3000)
    ir = decay_ir(rng_h, 96000, 96000)
    xin = rng_x.uniform(-1, 1, (12, 512)).astype(np.float32)
    pc = R.pconv(96000, 512)
    pc.push_ir(ir)
    out["pconv_cfg3_ir"] = ir
    out["pconv_cfg3_in"] = xin
    out["pconv_cfg3_out"] = np.stack([pc.convolution(b) for b in xin])

    # --- small partitioned convolution that wraps its rings twice, static and time-varying --------
    rng = np.random.default_rng(77)
    cvs, pts, nb = 1000, 64, 40  # nparts = 15 (truncating, Q4)
    ir = rng.standard_normal(cvs).astype(np.float32)
    xin = rng.uniform(-1, 1, (nb, pts)).astype(np.float32)
    x2 = rng.uniform(-1, 1, (nb, pts)).astype(np.float32)
    pc = R.pconv(cvs, pts)
    pc.push_ir(ir)
    out["pconv_small_ir"], out["pconv_small_in"], out["pconv_small_in2"] = ir, xin, x2
    out["pconv_small_out"] = np.stack([pc.convolution(b) for b in xin])
    pc = R.pconv(cvs, pts)
    out["pconv_small_tv_out"] = np.stack([pc.convolution(a, b) for a, b in zip(xin, x2)])

    # --- S4: direct convolution 4096 taps / 256 (config 4), one channel, 6 blocks + time-varying ---
    rng_h, rng_x = np.random.default_rng(4000), np.random.default_rng(5000)
    h = (rng_h.standard_normal(4096) / 64).astype(np.float32)
    xin = rng_x.uniform(-1, 1, (6, 256)).astype(np.float32)
    dc = R.dconv(4096, 256)
    dc.push_ir(h)
    out["dconv_cfg4_ir"], out["dconv_cfg4_in"] = h, xin
    out["dconv_cfg4_out"] = np.stack([dc.convolution(b) for b in xin])
    rng = np.random.default_rng(78)
    xin = rng.uniform(-1, 1, (30, 16)).astype(np.float32)
    x2 = rng.uniform(-1, 1, (30, 16)).astype(np.float32)
    dc = R.dconv(64, 16)
    out["dconv_tv_in"], out["dconv_tv_in2"] = xin, x2
    out["dconv_tv_out"] = np.stack([dc.convolution(a, b) for a, b in zip(xin, x2)])

    path = os.path.join(HERE, "reference_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(out), "arrays")


def oracle_pi():
    return 3.141592653589793


if __name__ == "__main__":
    main()
