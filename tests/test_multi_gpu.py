"""One handle, several GPUs (include/b200fft.h `*_multi`): contiguous channel / transform ranges per device,
one worker thread and stream per device, no communication. The result must be bit-identical to the single-device
handle's -- the same kernels run on the same data, only elsewhere. The two-device cases need a box with >= 2 GPUs
(gpurun --gpus 2); the one-device case runs the same fan-out machinery on any GPU box."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _devsets(eng):
    n = eng.device_count()
    return [[0]] + ([[0, 1], [1, 0]] if n >= 2 else []) + ([list(range(n))] if n > 2 else [])


def test_pconv_multi_equals_single_device(eng):
    pts, nparts, channels, nb = 512, 7, 37, 2 * 7 + 2
    cvs = pts * nparts
    rng = np.random.default_rng(3)
    ir = (rng.standard_normal((channels, cvs)) * 0.05).astype(np.float32)
    x = rng.uniform(-1, 1, (nb, channels, pts)).astype(np.float32)
    x2 = (rng.uniform(-1, 1, (nb, channels, pts)) * 0.05).astype(np.float32)

    def run(**kw):
        c = eng.Clpconv(0, cvs, pts, channels=channels, **kw)
        assert c.get_cl_err() == 0 and c.nparts == nparts and c.push_ir(ir) == 0
        y = np.zeros_like(x)
        for t in range(nb):
            assert c.convolution(y[t], x[t]) == 0
        assert c.reset() == 0
        ytv = np.zeros_like(x)
        for t in range(nb):
            assert c.convolution(ytv[t], x[t], x2[t]) == 0
        return y, ytv

    want = run()
    for devs in _devsets(eng):
        got = run(devices=devs)
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]), devs


def test_dconv_and_fft_multi_equal_single_device(eng):
    rng = np.random.default_rng(4)
    irsize, vsize, channels, nb = 256, 32, 11, 5
    h = (rng.standard_normal((channels, irsize)) / 16).astype(np.float32)
    x = rng.uniform(-1, 1, (channels, nb * vsize)).astype(np.float32)

    def drun(**kw):
        d = eng.Cldconv(0, irsize, vsize, channels=channels, max_blocks=nb, **kw)
        assert d.get_cl_err() == 0 and d.push_ir(h) == 0
        y = np.zeros_like(x)
        assert d.convolution(y, x, nblocks=nb) == 0
        return y

    size, batch = 4096, 13
    r = rng.uniform(-1, 1, (batch, size)).astype(np.float32)
    z = (rng.uniform(-1, 1, (batch, 1024)) + 1j * rng.uniform(-1, 1, (batch, 1024))).astype(np.complex64)

    def frun(**kw):
        f = eng.Clrfft(0, size, True, max_batch=batch, **kw)
        c = np.zeros((batch, size // 2), np.complex64)
        assert f.get_error() == 0 and f.transform(c.reshape(-1), r.reshape(-1).copy()) == 0
        p = eng.Clcfft(0, 1024, True, max_batch=batch, **kw)
        y = z.copy()
        assert p.transform(y.reshape(-1)) == 0
        assert p.transform(np.zeros((batch + 1) * 1024, np.complex64)) == 6  # over max_batch
        return c, y

    dwant, fwant = drun(), frun()
    for devs in _devsets(eng):
        assert np.array_equal(drun(devices=devs), dwant), devs
        got = frun(devices=devs)
        assert np.array_equal(got[0], fwant[0]) and np.array_equal(got[1], fwant[1]), devs


def test_push_ir_shard_by_shard(eng):
    pts, nparts, channels = 64, 3, 10
    cvs = pts * nparts
    rng = np.random.default_rng(8)
    ir = rng.standard_normal((channels, cvs)).astype(np.float32)
    x = rng.uniform(-1, 1, (channels, pts)).astype(np.float32)
    for devs in _devsets(eng):
        a = eng.Clpconv(0, cvs, pts, channels=channels, devices=devs)
        b = eng.Clpconv(0, cvs, pts, channels=channels, devices=devs)
        assert a.push_ir(ir) == 0
        n = len(devs)
        for g in range(n):
            lo, hi = g * channels // n, (g + 1) * channels // n
            assert b.push_ir_shard(g, ir[lo:hi]) == 0
        assert b.push_ir_shard(n, ir) == 2 and b.push_ir_shard(0, ir[:0]) == 2
        ya, yb = np.zeros_like(x), np.zeros_like(x)
        for _ in range(nparts + 1):
            assert a.convolution(ya, x) == 0 and b.convolution(yb, x) == 0
        assert np.array_equal(ya, yb)


def test_multi_argument_errors(eng):
    assert eng.Clpconv(0, 4096, 512, channels=4, devices=[0, 0], uData=1).get_cl_err() == 2  # the same device twice
    assert eng.Clpconv(0, 4096, 512, channels=1, devices=[0, 99], uData=1).get_cl_err() == 2  # fewer channels than devices
    assert eng.Clpconv(0, 4096, 512, channels=2, devices=[0, 99], uData=1).get_cl_err() == 1  # no such device
    c = eng.Clpconv(0, 4096, 512, channels=2, devices=[0])
    with pytest.raises(TypeError):
        c.convolution_dev(0, 0)
