"""CPU tests of the parity oracle itself (no GPU): the C restatement (oracle/ref_cpu.c) against
  (a) the known answers implied by the reference's own test programs,
  (b) the unmodified reference run through oracle/minicl, bit for bit, when that library is present,
  (c) the committed golden vectors the reference generated (tests/golden/make_golden.py), bit for bit,
  (d) float64 ground truth, with the reference's documented quirks made explicit."""
import numpy as np
import pytest

from conftest import rel_l2


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


# ---- (a) known answers -------------------------------------------------------------------------------
def test_kat_cfft_sine(port):
    # reference test_cfft.cpp:54-56: 16-point sine -> bins 1 and 15 = (0, -/+0.5); round trip = input
    N = 16
    x = np.zeros(N, np.complex64)
    x.real = np.sin(np.arange(N) * 2 * np.pi / N).astype(np.float32)
    s = port.cfft(x, True)
    want = np.zeros(N, np.complex64)
    want[1], want[15] = -0.5j, 0.5j
    assert np.abs(s - want).max() < 1e-6
    assert np.abs(port.cfft(s, False) - x).max() < 1e-6


def test_kat_rfft_dc_fund_nyq(port):
    # reference test_rfft.cpp:54-57: dc + fundamental + nyquist -> [(0.5,0.5), (0,-1), 0, ...]
    N = 16
    i = np.arange(N)
    r = (0.5 + np.sin(i * 2 * np.pi / N) + 0.5 * np.cos(i * np.pi)).astype(np.float32)
    s = port.rfft_fwd(r)
    want = np.zeros(N // 2, np.complex64)
    want[0], want[1] = 0.5 + 0.5j, -1j
    assert np.abs(s - want).max() < 1e-6
    assert np.abs(port.rfft_inv(s) - r).max() < 1e-6


# ---- (b) restatement == unmodified reference, bit for bit ----------------------------------------------
@pytest.mark.parametrize("N", [2, 4, 8, 16, 64, 512, 1024, 4096])
@pytest.mark.parametrize("fwd", [True, False])
def test_port_equals_reference_cfft(port, ref, N, fwd):
    rng = np.random.default_rng(N + fwd)
    x = (rng.uniform(-1, 1, N) + 1j * rng.uniform(-1, 1, N)).astype(np.complex64)
    assert np.array_equal(bits(port.cfft(x, fwd)), bits(ref.cfft(x, fwd)))


@pytest.mark.parametrize("size", [4, 8, 16, 256, 4096, 65536])
def test_port_equals_reference_rfft(port, ref, size):
    rng = np.random.default_rng(size)
    r = rng.uniform(-1, 1, size).astype(np.float32)
    a, b = port.rfft_fwd(r), ref.rfft_fwd(r)
    assert np.array_equal(bits(a), bits(b))
    assert np.array_equal(bits(ref.rfft_fwd(r, out_of_place=True)), bits(b))
    assert np.array_equal(bits(port.rfft_inv(a)), bits(ref.rfft_inv(b)))


@pytest.mark.parametrize("cvs,pts,nb", [(64, 16, 12), (100, 16, 15), (2048, 512, 10), (7000, 64, 8)])
def test_port_equals_reference_pconv(port, ref, cvs, pts, nb):
    rng = np.random.default_rng(cvs)
    ir = rng.standard_normal(cvs).astype(np.float32)
    x = rng.uniform(-1, 1, (nb, pts)).astype(np.float32)
    x2 = rng.uniform(-1, 1, (nb, pts)).astype(np.float32)
    a, b = port.pconv(cvs, pts), ref.pconv(cvs, pts)
    a.push_ir(ir)
    b.push_ir(ir)
    for i in range(nb):
        assert np.array_equal(bits(a.convolution(x[i])), bits(b.convolution(x[i])))
    a, b = port.pconv(cvs, pts), ref.pconv(cvs, pts)
    for i in range(nb):  # time-varying, no push_ir: the IR is recorded from the second input
        assert np.array_equal(bits(a.convolution(x[i], x2[i])), bits(b.convolution(x[i], x2[i])))


@pytest.mark.parametrize("irsize,vsize,nb", [(64, 16, 12), (4096, 256, 3), (100, 16, 20)])
def test_port_equals_reference_dconv(port, ref, irsize, vsize, nb):
    # (100, 16) has irsize % vsize != 0: the reference's ring write is broken there (Q10) and the
    # restatement reproduces the breakage literally
    rng = np.random.default_rng(irsize)
    ir = rng.standard_normal(irsize).astype(np.float32)
    x = rng.uniform(-1, 1, (nb, vsize)).astype(np.float32)
    x2 = rng.uniform(-1, 1, (nb, vsize)).astype(np.float32)
    a, b = port.dconv(irsize, vsize), ref.dconv(irsize, vsize)
    a.push_ir(ir)
    b.push_ir(ir)
    for i in range(nb):
        assert np.array_equal(bits(a.convolution(x[i])), bits(b.convolution(x[i])))
    a, b = port.dconv(irsize, vsize), ref.dconv(irsize, vsize)
    for i in range(nb):
        assert np.array_equal(bits(a.convolution(x[i], x2[i])), bits(b.convolution(x[i], x2[i])))


# ---- (c) restatement == golden vectors produced by the reference ----------------------------------------
def test_port_equals_golden(port, golden):
    g = golden
    assert np.array_equal(bits(port.cfft(g["kat_cfft_in"], True)), bits(g["kat_cfft_spec"]))
    assert np.array_equal(bits(port.cfft(g["kat_cfft_spec"], False)), bits(g["kat_cfft_back"]))
    assert np.array_equal(bits(port.rfft_fwd(g["kat_rfft_in"])), bits(g["kat_rfft_spec"]))
    assert np.array_equal(bits(port.rfft_inv(g["kat_rfft_spec"])), bits(g["kat_rfft_back"]))
    assert np.array_equal(bits(port.cfft(g["cfft1024_in"], True)), bits(g["cfft1024_fwd"]))
    assert np.array_equal(bits(port.cfft(g["cfft1024_in"], False)), bits(g["cfft1024_inv"]))
    assert np.array_equal(bits(port.rfft_fwd(g["rfft4096_in"])), bits(g["rfft4096_fwd"]))
    assert np.array_equal(bits(port.rfft_inv(g["rfft4096_fwd"])), bits(g["rfft4096_back"]))
    assert np.array_equal(bits(port.rfft_fwd(g["rfft65536_in"])), bits(g["rfft65536_fwd"]))


def test_port_equals_golden_convolvers(port, golden):
    g = golden
    pc = port.pconv(96000, 512)
    assert pc.nparts == 187  # truncating division, Q4
    pc.push_ir(g["pconv_cfg3_ir"])
    got = np.stack([pc.convolution(b) for b in g["pconv_cfg3_in"]])
    assert np.array_equal(bits(got), bits(g["pconv_cfg3_out"]))
    pc = port.pconv(1000, 64)
    pc.push_ir(g["pconv_small_ir"])
    got = np.stack([pc.convolution(b) for b in g["pconv_small_in"]])
    assert np.array_equal(bits(got), bits(g["pconv_small_out"]))
    pc = port.pconv(1000, 64)
    got = np.stack([pc.convolution(a, b) for a, b in zip(g["pconv_small_in"], g["pconv_small_in2"])])
    assert np.array_equal(bits(got), bits(g["pconv_small_tv_out"]))
    dc = port.dconv(4096, 256)
    dc.push_ir(g["dconv_cfg4_ir"])
    got = np.stack([dc.convolution(b) for b in g["dconv_cfg4_in"]])
    assert np.array_equal(bits(got), bits(g["dconv_cfg4_out"]))
    dc = port.dconv(64, 16)
    got = np.stack([dc.convolution(a, b) for a, b in zip(g["dconv_tv_in"], g["dconv_tv_in2"])])
    assert np.array_equal(bits(got), bits(g["dconv_tv_out"]))


# ---- (d) against float64 ground truth, quirks explicit -------------------------------------------------
@pytest.mark.parametrize("N", [16, 1024, 8192])
def test_cfft_vs_float64(port, N):
    rng = np.random.default_rng(N)
    x = (rng.uniform(-1, 1, N) + 1j * rng.uniform(-1, 1, N)).astype(np.complex64)
    truth = np.fft.fft(x.astype(np.complex128))
    assert rel_l2(port.cfft(x, True), truth / N) < 1e-6  # Q1: forward scaled by 1/N
    assert rel_l2(port.cfft(x, False), np.conj(np.fft.fft(np.conj(x.astype(np.complex128))))) < 1e-6
    if N <= 1024:
        assert rel_l2(port.dft64(x, -1), truth) < 1e-12  # the oracle's own naive double DFT


def test_rfft_convention_and_quirk_q3(port):
    size = 4096
    rng = np.random.default_rng(3)
    r = rng.uniform(-1, 1, size).astype(np.float32)
    s = port.rfft_fwd(r).astype(np.complex128)
    X = np.fft.rfft(r.astype(np.float64))
    want = 2 * X[: size // 2] / size                       # Q2: bins scaled 2/size
    want[0] = (X[0].real + 1j * X[size // 2].real) / size  # packed (DC, Nyquist) scaled 1/size
    want[size // 4] = np.conj(want[size // 4])             # Q3: bin size/4 is never visited by the split
    assert rel_l2(s, want) < 1e-6


def test_pconv_half_weight_dc_quirk_q5(port):
    cvs, pts, nb = 2048, 256, 24
    rng = np.random.default_rng(5)
    ir = rng.standard_normal(cvs).astype(np.float32)
    x = rng.uniform(-1, 1, (nb, pts)).astype(np.float32)
    pc = port.pconv(cvs, pts)
    pc.push_ir(ir)
    y = np.concatenate([pc.convolution(b) for b in x])
    exact = np.convolve(x.ravel().astype(np.float64), ir.astype(np.float64))[: nb * pts]
    # the reference is NOT the exact linear convolution: the DC and Nyquist bins of every 2*pts frame
    # product arrive at half weight. Model that in float64 and it matches to float32 rounding.
    H = [np.fft.rfft(np.r_[ir[i * pts:(i + 1) * pts].astype(np.float64), np.zeros(pts)]) for i in range(cvs // pts)]
    X = [np.fft.rfft(np.r_[b.astype(np.float64), np.zeros(pts)]) for b in x]
    out, tail = [], np.zeros(pts)
    for t in range(nb):
        Y = sum(X[t - a] * H[a] for a in range(len(H)) if t - a >= 0)
        Y[0] *= 0.5
        Y[-1] *= 0.5
        yy = np.fft.irfft(Y)
        out.append(yy[:pts] + tail)
        tail = yy[pts:]
    model = np.concatenate(out)
    assert rel_l2(y, model) < 2e-6
    assert rel_l2(y, exact) > 1e-3  # and it is measurably not the exact convolution


def test_dconv_one_sample_delay_quirk_q9(port):
    irsize, vsize, nb = 256, 32, 20
    rng = np.random.default_rng(9)
    ir = rng.standard_normal(irsize).astype(np.float32)
    x = rng.uniform(-1, 1, (nb, vsize)).astype(np.float32)
    dc = port.dconv(irsize, vsize)
    dc.push_ir(ir)
    y = np.concatenate([dc.convolution(b) for b in x])
    exact = np.convolve(x.ravel().astype(np.float64), ir.astype(np.float64))[: nb * vsize]
    delayed = np.r_[0.0, exact[:-1]]
    assert y[0] == 0.0
    assert np.abs(y - delayed).max() < 1e-4
