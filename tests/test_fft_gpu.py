"""GPU parity of the FFT path, through the C ABI (opencl_fft_b200 -> libb200fft.so), against the oracle
(oracle/ref_cpu.c), the golden vectors produced by the real reference, and float64 ground truth.
Tolerance: relative L2 <= 1e-5 (BASELINE.json north_star); measured values are ~1e-7."""
import numpy as np
import pytest

from conftest import TOL, rel_l2

pytestmark = pytest.mark.gpu


def crand(rng, *shape):
    return (rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1, 1, shape)).astype(np.complex64)


def test_kat_cfft(eng, golden):
    f, i = eng.Clcfft(0, 16, True), eng.Clcfft(0, 16, False)
    assert f.get_error() == 0 and i.get_error() == 0
    x = golden["kat_cfft_in"].copy()
    assert f.transform(x) == 0
    want = np.zeros(16, np.complex64)
    want[1], want[15] = -0.5j, 0.5j
    assert np.abs(x - want).max() < 1e-6
    assert rel_l2(x, golden["kat_cfft_spec"]) < TOL
    assert i.transform(x) == 0
    assert np.abs(x - golden["kat_cfft_in"]).max() < 1e-6


def test_kat_rfft(eng, golden):
    f, i = eng.Clrfft(0, 16, True), eng.Clrfft(0, 16, False)
    r = golden["kat_rfft_in"].copy()
    spec = np.zeros(8, np.complex64)
    assert f.transform(spec, r) == 0  # out-of-place form, as test_rfft.cpp:64 uses it
    want = np.zeros(8, np.complex64)
    want[0], want[1] = 0.5 + 0.5j, -1j
    assert np.abs(spec - want).max() < 1e-6
    back = np.zeros(16, np.float32)
    assert i.transform(spec, back) == 0
    assert np.abs(back - golden["kat_rfft_in"]).max() < 1e-6
    assert np.abs(spec.view(np.float32) - back).max() == 0  # c is overwritten too (cl_fft.cpp:290-293)


@pytest.mark.parametrize("logn", range(1, 17))
@pytest.mark.parametrize("fwd", [True, False])
def test_cfft_vs_oracle_all_sizes(eng, port, logn, fwd):
    N = 1 << logn
    rng = np.random.default_rng(1001 + logn)
    x = crand(rng, N)
    p = eng.Clcfft(0, N, fwd)
    assert p.get_error() == 0
    y = x.copy()
    assert p.transform(y) == 0
    assert rel_l2(y, port.cfft(x, fwd)) < TOL
    truth = np.fft.fft(x.astype(np.complex128)) / N if fwd else np.fft.ifft(x.astype(np.complex128)) * N
    assert rel_l2(y, truth) < 2e-6


@pytest.mark.parametrize("logn,batch", [(1, 700), (4, 1000), (5, 333), (6, 257), (8, 129), (10, 67), (12, 9), (13, 5), (14, 3), (15, 5), (16, 3)])
def test_cfft_batched(eng, port, logn, batch):
    N = 1 << logn
    rng = np.random.default_rng(logn)
    x = crand(rng, batch, N)
    p = eng.Clcfft(0, N, True, max_batch=batch)
    y = x.copy()
    assert p.transform(y.reshape(-1)) == 0
    truth = np.fft.fft(x.astype(np.complex128), axis=1) / N
    assert rel_l2(y, truth) < 2e-6
    for b in (0, batch // 2, batch - 1):
        assert rel_l2(y[b], port.cfft(x[b], True)) < TOL
    # round trip through the inverse plan
    q = eng.Clcfft(0, N, False, max_batch=batch)
    assert q.transform(y.reshape(-1)) == 0
    assert rel_l2(y, x) < 2e-6
    assert p.transform(np.zeros((batch + 1) * N, np.complex64)) == 6  # over max_batch


@pytest.mark.parametrize("logn", [1, 2, 3, 4, 5])
@pytest.mark.parametrize("batch", [1, 127, 128, 129, 1023, 1025, 2500])
def test_one_thread_per_transform_kernels_every_transform(eng, port, logn, batch):
    """N <= 32 run one thread per transform (fft_thread_kernel): batches that end inside a CTA, on a CTA
    boundary and one past it; complex both ways, real forward and inverse, every transform checked."""
    N = 1 << logn
    rng = np.random.default_rng(77 * logn + batch)
    x = crand(rng, batch, N)
    for fwd in (True, False):
        y = x.copy()
        assert eng.Clcfft(0, N, fwd, max_batch=batch).transform(y.reshape(-1)) == 0
        x64 = x.astype(np.complex128)
        truth = np.fft.fft(x64, axis=1) / N if fwd else np.fft.ifft(x64, axis=1) * N
        assert max(rel_l2(y[b], truth[b]) for b in range(batch)) < 2e-6
        for b in (0, batch // 2, batch - 1):
            assert rel_l2(y[b], port.cfft(x[b], fwd)) < TOL
    size = 2 * N
    r = rng.uniform(-1, 1, (batch, size)).astype(np.float32)
    c = np.zeros((batch, N), np.complex64)
    assert eng.Clrfft(0, size, True, max_batch=batch).transform(c.reshape(-1), r.reshape(-1).copy()) == 0
    X = np.fft.rfft(r.astype(np.float64), axis=1)
    want = 2 * X[:, :N] / size
    want[:, 0] = (X[:, 0].real + 1j * X[:, N].real) / size  # packed (DC, Nyquist), Q2
    want[:, N // 2] = np.conj(want[:, N // 2])  # Q3: bin size/4 conjugated
    assert max(rel_l2(c[b], want[b]) for b in range(batch)) < 2e-6
    for b in (0, batch // 2, batch - 1):
        assert rel_l2(c[b], port.rfft_fwd(r[b])) < TOL
    back = np.zeros((batch, size), np.float32)
    assert eng.Clrfft(0, size, False, max_batch=batch).transform(c.reshape(-1), back.reshape(-1)) == 0
    assert np.abs(back - r).max() < 2e-5
    for b in (0, batch - 1):
        assert rel_l2(back[b], port.rfft_inv(port.rfft_fwd(r[b]))) < TOL


def test_golden_cfft_rfft(eng, golden):
    g = golden
    for fwd, key in ((True, "cfft1024_fwd"), (False, "cfft1024_inv")):
        y = g["cfft1024_in"].copy()
        assert eng.Clcfft(0, 1024, fwd).transform(y) == 0
        assert rel_l2(y, g[key]) < TOL
    c = g["rfft4096_in"].copy().view(np.complex64)
    assert eng.Clrfft(0, 4096, True).transform(c) == 0  # in-place form
    assert rel_l2(c, g["rfft4096_fwd"]) < TOL
    assert eng.Clrfft(0, 4096, False).transform(c) == 0
    assert rel_l2(c.view(np.float32), g["rfft4096_back"]) < TOL
    assert np.abs(c.view(np.float32) - g["rfft4096_in"]).max() < 1e-5
    c = g["rfft65536_in"].copy().view(np.complex64)
    assert eng.Clrfft(0, 65536, True).transform(c) == 0
    assert rel_l2(c, g["rfft65536_fwd"]) < TOL


@pytest.mark.parametrize("logs", range(2, 18))
def test_rfft_vs_oracle_all_sizes(eng, port, logs):
    size = 1 << logs
    rng = np.random.default_rng(1002 + logs)
    r = rng.uniform(-1, 1, size).astype(np.float32)
    f, i = eng.Clrfft(0, size, True), eng.Clrfft(0, size, False)
    assert f.get_error() == 0 and i.get_error() == 0
    c = np.zeros(size // 2, np.complex64)
    assert f.transform(c, r.copy()) == 0
    want = port.rfft_fwd(r)
    assert rel_l2(c, want) < TOL
    # quirk Q3: bin size/4 is the conjugate of the true value
    X = np.fft.rfft(r.astype(np.float64))
    if size >= 8:
        assert abs(c[size // 4] - np.conj(2 * X[size // 4] / size)) < 1e-5
    back = np.zeros(size, np.float32)
    assert i.transform(c, back) == 0
    assert rel_l2(back, port.rfft_inv(want)) < TOL
    assert np.abs(back - r).max() < 2e-5


def test_rfft_batched_cfg5_shape_properties(eng):
    """BASELINE config 5a shape (65536-point real FFT), a 64-channel slice: properties that hold at any
    size -- round trip, linearity, Parseval in the reference's scaling."""
    size, ch = 65536, 64
    rng = np.random.default_rng(6000)
    a = rng.uniform(-1, 1, (ch, size)).astype(np.float32)
    b = rng.uniform(-1, 1, (ch, size)).astype(np.float32)
    f, i = eng.Clrfft(0, size, True, max_batch=ch), eng.Clrfft(0, size, False, max_batch=ch)
    A, B, AB = (np.zeros((ch, size // 2), np.complex64) for _ in range(3))
    assert f.transform(A.reshape(-1), a.reshape(-1).copy()) == 0
    assert f.transform(B.reshape(-1), b.reshape(-1).copy()) == 0
    assert f.transform(AB.reshape(-1), (a + 2 * b).reshape(-1).copy()) == 0
    assert rel_l2(AB, A + 2 * B) < 2e-6  # linearity
    # Parseval: sum x^2 / size = DC^2 + Nyq^2 + sum_{k>=1} |S_k|^2 / 2 in the reference's scaling
    lhs = (a.astype(np.float64) ** 2).sum(axis=1) / size
    S = A.astype(np.complex128)
    rhs = S[:, 0].real ** 2 + S[:, 0].imag ** 2 + 0.5 * (np.abs(S[:, 1:]) ** 2).sum(axis=1)
    assert np.abs(lhs - rhs).max() / lhs.max() < 1e-5
    back = np.zeros((ch, size), np.float32)
    assert i.transform(A.reshape(-1), back.reshape(-1)) == 0
    assert rel_l2(back, a) < 2e-6  # round trip
    truth = np.fft.rfft(a[7].astype(np.float64)) * 2 / size
    got = np.zeros(size // 2, np.complex64)
    f1 = eng.Clrfft(0, size, True)
    assert f1.transform(got, a[7].copy()) == 0
    k = np.r_[1:size // 4, size // 4 + 1:size // 2]
    assert rel_l2(got[k], truth[k]) < 2e-6


def test_device_pointer_api_matches_host_api(eng):
    import torch

    N, batch = 1024, 4096
    rng = np.random.default_rng(1)
    x = crand(rng, batch, N)
    p = eng.Clcfft(0, N, True, max_batch=batch)
    want = x.copy()
    assert p.transform(want.reshape(-1)) == 0
    d_in = torch.from_numpy(x.view(np.float32)).cuda()
    d_out = torch.empty_like(d_in)
    assert p.transform_dev(d_in, d_out, batch) == 0
    torch.cuda.synchronize()
    assert np.array_equal(d_out.cpu().numpy().view(np.complex64), want)  # same kernel, same bits
    assert p.transform_dev(d_in, d_in, batch) == 0  # in place
    torch.cuda.synchronize()
    assert np.array_equal(d_in.cpu().numpy().view(np.complex64), want)
    # determinism
    d2 = torch.from_numpy(x.view(np.float32)).cuda()
    assert p.transform_dev(d2, d2, batch) == 0
    torch.cuda.synchronize()
    assert torch.equal(d2, d_in)


@pytest.mark.parametrize("path", ["default", "one_sm", "four_step", "four_step_unfused"])
@pytest.mark.parametrize("batch", [1, 5, 71, 149, 300])
def test_large_fft_every_transform_of_a_batch(eng, batch, path, options):
    """The 65536-point real / 32768-point complex transforms. From ~100 transforms per call: the one-SM kernel
    (fft_sm.cuh: one pass over HBM, TMA-staged input, rows parked in tensor memory, persistent CTAs looping over the
    batch -- 149 and 300 make some CTAs take two or three transforms); below: the four-step launch pair
    (fft_large.cuh). Both are forced for every batch here. Check EVERY transform of the batch (a race shows up as a
    few wrong ones), twice, and that the two runs agree bit for bit; forward real, inverse real, complex both ways."""
    if path == "one_sm":
        options("fft_sm_min_batch", 1)
    if path.startswith("four_step"):
        options("fft_sm_min_batch", 0)
    if path == "four_step_unfused":
        options("separate_split", 1)
    size = 65536
    rng = np.random.default_rng(batch)
    x = rng.uniform(-1, 1, (batch, size)).astype(np.float32)
    f = eng.Clrfft(0, size, True, max_batch=batch)
    outs = []
    for _ in range(2):
        c = np.zeros((batch, size // 2), np.complex64)
        assert f.transform(c.reshape(-1), x.reshape(-1).copy()) == 0
        outs.append(c)
    assert np.array_equal(outs[0], outs[1])
    X = np.fft.rfft(x.astype(np.float64), axis=1)
    want = 2 * X[:, : size // 2] / size
    want[:, 0] = (X[:, 0].real + 1j * X[:, size // 2].real) / size
    want[:, size // 4] = np.conj(want[:, size // 4])  # quirk Q3
    err = np.linalg.norm(outs[0] - want, axis=1) / np.linalg.norm(want, axis=1)
    assert err.max() < 2e-6, (int(err.argmax()), float(err.max()))
    # inverse real transform of every spectrum gives the signal back
    inv = eng.Clrfft(0, size, False, max_batch=batch)
    spec, back = outs[0].copy(), np.zeros((batch, size), np.float32)
    assert inv.transform(spec.reshape(-1), back.reshape(-1)) == 0
    err = np.linalg.norm(back - x, axis=1) / np.linalg.norm(x, axis=1)
    assert err.max() < 2e-6, (int(err.argmax()), float(err.max()))
    # complex path, forward and inverse, same kernel family
    z = (rng.uniform(-1, 1, (batch, size // 2)) + 1j * rng.uniform(-1, 1, (batch, size // 2))).astype(np.complex64)
    for fwd in (True, False):
        p = eng.Clcfft(0, size // 2, fwd, max_batch=batch)
        y = z.copy()
        assert p.transform(y.reshape(-1)) == 0
        zz = z.astype(np.complex128)
        truth = np.fft.fft(zz, axis=1) / (size // 2) if fwd else np.fft.ifft(zz, axis=1) * (size // 2)
        err = np.linalg.norm(y - truth, axis=1) / np.linalg.norm(truth, axis=1)
        assert err.max() < 2e-6, (fwd, int(err.argmax()), float(err.max()))


def test_empty_batch_and_null_arguments(eng):
    p = eng.Clcfft(0, 64, True, max_batch=4)
    L = eng.lib()
    assert L.b2f_cfft_exec_host(p._h, np.zeros(64, np.complex64).ctypes.data, 0) == 0  # empty batch: nothing to do
    assert L.b2f_cfft_exec_host(p._h, None, 1) == 2
    assert L.b2f_cfft_exec_host(None, None, 1) == 2
    assert L.b2f_cfft_exec_dev(p._h, None, None, 1, None) == 2
    r = eng.Clrfft(0, 64, True)
    assert L.b2f_rfft_exec_host(r._h, None, None, 1) == 2
    assert eng.Clcfft(7, 64, True).get_error() == 1  # no such device


def test_second_device_if_present(eng):
    if eng.device_count() < 2:
        pytest.skip("single-GPU box")
    rng = np.random.default_rng(2)
    x = crand(rng, 1024)
    a, b = x.copy(), x.copy()
    assert eng.Clcfft(0, 1024, True).transform(a) == 0
    assert eng.Clcfft(1, 1024, True).transform(b) == 0
    assert np.array_equal(a, b)


@pytest.mark.parametrize("batch", [1, 2, 3, 149, 297, 300])
def test_one_sm_kernel_16384_point_complex(eng, options, batch):
    """N = 2^14 on the one-SM kernel: two transforms per unit of work (the second parked in tensor memory while the
    first is processed), so odd batches end in a half-empty unit and 297+ make CTAs loop. Every transform, forward and
    inverse, against float64; then the default selection (small batches on cfft_kernel<14>) must agree to rounding."""
    N = 16384
    rng = np.random.default_rng(batch)
    z = crand(rng, batch, N)
    zz = z.astype(np.complex128)
    got = {}
    for forced in (1, 0):
        options("fft_sm_min_batch", forced)
        for fwd in (True, False):
            p = eng.Clcfft(0, N, fwd, max_batch=batch)
            y = z.copy()
            assert p.transform(y.reshape(-1)) == 0
            truth = np.fft.fft(zz, axis=1) / N if fwd else np.fft.ifft(zz, axis=1) * N
            err = np.linalg.norm(y - truth, axis=1) / np.linalg.norm(truth, axis=1)
            assert err.max() < 2e-6, (forced, fwd, int(err.argmax()), float(err.max()))
            got[forced, fwd] = y
    assert rel_l2(got[1, True], got[0, True]) < 1e-6 and rel_l2(got[1, False], got[0, False]) < 1e-6


@pytest.mark.parametrize("batch", [1, 3, 4, 5, 590, 593, 600])
def test_one_sm_kernel_8192_point_complex(eng, options, batch):
    """N = 2^13 on the one-SM kernel: FOUR transforms per unit of work (two per job, rows 0..7 each; the second job
    parked in tensor memory), so batches that are not multiples of four end in a partly empty unit and 593+ make CTAs
    loop. Every transform, forward and inverse, against float64; in place equals out of place; and the default selection
    for small batches (cfft_kernel<13>) agrees to rounding."""
    import torch

    N = 8192
    rng = np.random.default_rng(batch)
    z = crand(rng, batch, N)
    zz = z.astype(np.complex128)
    got = {}
    for forced in (1, 0):
        options("fft_sm_min_batch", forced)
        for fwd in (True, False):
            p = eng.Clcfft(0, N, fwd, max_batch=batch)
            y = z.copy()
            assert p.transform(y.reshape(-1)) == 0
            truth = np.fft.fft(zz, axis=1) / N if fwd else np.fft.ifft(zz, axis=1) * N
            err = np.linalg.norm(y - truth, axis=1) / np.linalg.norm(truth, axis=1)
            assert err.max() < 2e-6, (forced, fwd, int(err.argmax()), float(err.max()))
            got[forced, fwd] = y
            if forced and fwd:
                d = torch.from_numpy(z.view(np.float32).copy()).cuda()
                o = torch.empty_like(d)
                assert p.transform_dev(d, o, batch) == 0 and p.transform_dev(d, d, batch) == 0
                torch.cuda.synchronize()
                assert torch.equal(d, o) and np.array_equal(o.cpu().numpy().view(np.complex64), y)
    assert rel_l2(got[1, True], got[0, True]) < 1e-6 and rel_l2(got[1, False], got[0, False]) < 1e-6


@pytest.mark.parametrize("batch", [4, 5, 7, 593, 600])
def test_one_sm_kernel_16384_point_inverse_real(eng, port, options, batch):
    """The inverse 16384-point real transform (N = 2^13 complex) on the one-SM kernel, four transforms per unit: the
    unsplit pairs staged runs j1 <-> 7 - j1 of the same transform. Every transform against the one-CTA kernel (to
    rounding), the oracle for two of them, and the round trip through the forward transform."""
    size = 16384
    rng = np.random.default_rng(batch)
    x = rng.uniform(-1, 1, (batch, size)).astype(np.float32)
    options("fft_sm_min_batch", 0)
    spec = np.zeros((batch, size // 2), np.complex64)
    assert eng.Clrfft(0, size, True, max_batch=batch).transform(spec.reshape(-1), x.reshape(-1).copy()) == 0
    got = {}
    for forced in (1, 0):
        options("fft_sm_min_batch", forced)
        iv = eng.Clrfft(0, size, False, max_batch=batch)
        c, r = spec.copy(), np.zeros((batch, size), np.float32)
        assert iv.transform(c.reshape(-1), r.reshape(-1)) == 0
        got[forced] = r
    err = np.linalg.norm(got[1] - got[0], axis=1) / np.linalg.norm(got[0], axis=1)
    assert err.max() < 1e-6, (int(err.argmax()), float(err.max()))
    assert not np.array_equal(got[1], got[0])  # (the forced plan really took the other kernel)
    back = np.linalg.norm(got[1] - x, axis=1) / np.linalg.norm(x, axis=1)
    assert back.max() < 2e-6
    for k in (0, batch - 1):
        assert rel_l2(got[1][k], port.rfft_inv(spec[k])) < 1e-6


@pytest.mark.parametrize("batch", [1, 2, 3, 149, 297, 300])
def test_one_sm_kernel_32768_point_real(eng, port, options, batch):
    """The 32768-point real transform (N = 2^14 complex) on the one-SM kernel: the split pairs lanes of one warp inside a
    job and row 0 across warps, the unsplit pairs staged runs j1 <-> 15 - j1 of the same transform. Every transform,
    forward and inverse, against float64 with the reference's quirks (Q2, Q3), a few against the oracle; run twice
    (bit-identical); the default selection for small batches (one CTA per transform) must agree to rounding."""
    size = 32768
    rng = np.random.default_rng(1000 + batch)
    x = rng.uniform(-1, 1, (batch, size)).astype(np.float32)
    X = np.fft.rfft(x.astype(np.float64), axis=1)
    want = 2 * X[:, : size // 2] / size
    want[:, 0] = (X[:, 0].real + 1j * X[:, size // 2].real) / size
    want[:, size // 4] = np.conj(want[:, size // 4])
    got = {}
    for forced in (1, 0):
        options("fft_sm_min_batch", forced)
        f, inv = eng.Clrfft(0, size, True, max_batch=batch), eng.Clrfft(0, size, False, max_batch=batch)
        outs = []
        for _ in range(2):
            c = np.zeros((batch, size // 2), np.complex64)
            assert f.transform(c.reshape(-1), x.reshape(-1).copy()) == 0
            outs.append(c)
        assert np.array_equal(outs[0], outs[1])
        err = np.linalg.norm(outs[0] - want, axis=1) / np.linalg.norm(want, axis=1)
        assert err.max() < 2e-6, (forced, int(err.argmax()), float(err.max()))
        for b in (0, batch - 1):
            assert rel_l2(outs[0][b], port.rfft_fwd(x[b])) < TOL
        back = np.zeros((batch, size), np.float32)
        assert inv.transform(outs[0].copy().reshape(-1), back.reshape(-1)) == 0
        err = np.linalg.norm(back - x, axis=1) / np.linalg.norm(x, axis=1)
        assert err.max() < 2e-6, (forced, int(err.argmax()), float(err.max()))
        assert rel_l2(back[batch - 1], port.rfft_inv(port.rfft_fwd(x[batch - 1]))) < TOL
        got[forced] = (outs[0], back)
    assert rel_l2(got[1][0], got[0][0]) < 1e-6 and rel_l2(got[1][1], got[0][1]) < 1e-6


@pytest.mark.parametrize("size,batch", [(4096, 1031), (65536, 800)])
def test_pipelined_host_call_equals_single_stream(eng, options, size, batch):
    """Host calls above 1 MB cut the batch into chunks alternating between two streams (upload, transform and download
    of different chunks overlap). Same bits as the single-stream call, real and complex, odd batch included; 65536 x 800
    puts 100 transforms per chunk on the one-SM kernel."""
    rng = np.random.default_rng(size)
    r = rng.uniform(-1, 1, (batch, size)).astype(np.float32)
    z = crand(rng, batch // 2, size // 2)
    res = {}
    for pipe in (1, 0):
        options("pconv_pipeline", pipe)
        f = eng.Clrfft(0, size, True, max_batch=batch)
        c = np.zeros((batch, size // 2), np.complex64)
        assert f.transform(c.reshape(-1), r.reshape(-1).copy()) == 0
        i = eng.Clrfft(0, size, False, max_batch=batch)
        back = np.zeros((batch, size), np.float32)
        assert i.transform(c.copy().reshape(-1), back.reshape(-1)) == 0
        p = eng.Clcfft(0, size // 2, True, max_batch=batch // 2)
        y = z.copy()
        assert p.transform(y.reshape(-1)) == 0
        res[pipe] = (c, back, y)
    for a, b in zip(res[0], res[1]):
        assert np.array_equal(a, b)
    assert np.abs(res[1][1] - r).max() < 2e-5
