"""GPU parity of the direct convolution (Cldconv) through the C ABI against the oracle, the golden
vectors of the real reference and float64."""
import numpy as np
import pytest

from conftest import TOL, rel_l2

pytestmark = pytest.mark.gpu


def test_golden_cfg4_and_time_varying(eng, golden):
    g = golden
    c = eng.Cldconv(0, 4096, 256)
    assert c.get_cl_err() == 0
    assert c.push_ir(g["dconv_cfg4_ir"]) == 0
    y = np.zeros_like(g["dconv_cfg4_in"])
    for t in range(y.shape[0]):
        assert c.convolution(y[t], g["dconv_cfg4_in"][t]) == 0
    assert rel_l2(y, g["dconv_cfg4_out"]) < TOL
    assert y[0, 0] == 0.0  # one-sample delay, Q9
    c = eng.Cldconv(0, 64, 16)
    y = np.zeros_like(g["dconv_tv_in"])
    for t in range(y.shape[0]):
        assert c.convolution(y[t], g["dconv_tv_in"][t], g["dconv_tv_in2"][t]) == 0
    assert rel_l2(y, g["dconv_tv_out"]) < TOL


@pytest.mark.parametrize("irsize,vsize", [(16, 16), (64, 16), (256, 64), (1024, 128), (4096, 256), (8192, 64), (20, 1), (4096, 1)])
def test_vs_oracle(eng, port, irsize, vsize):
    nb = 2 * (irsize // vsize) + 5 if irsize // vsize < 40 else 50
    rng = np.random.default_rng(irsize + vsize)
    ir = (rng.standard_normal(irsize) / np.sqrt(irsize)).astype(np.float32)
    x = rng.uniform(-1, 1, (nb, vsize)).astype(np.float32)
    c, o = eng.Cldconv(0, irsize, vsize), port.dconv(irsize, vsize)
    assert c.push_ir(ir) == 0
    o.push_ir(ir)
    y = np.zeros_like(x)
    for t in range(nb):
        assert c.convolution(y[t], x[t]) == 0
    want = np.stack([o.convolution(b) for b in x])
    assert rel_l2(y, want) < TOL
    exact = np.convolve(x.ravel().astype(np.float64), ir.astype(np.float64))[: nb * vsize]
    assert rel_l2(y.ravel(), np.r_[0.0, exact[:-1]]) < 2e-6


@pytest.mark.parametrize("irsize,vsize", [(64, 16), (512, 64), (4096, 256)])
def test_time_varying_vs_oracle(eng, port, irsize, vsize):
    nb = 3 * (irsize // vsize + 1) + 2
    rng = np.random.default_rng(irsize)
    ir = (rng.standard_normal(irsize) / np.sqrt(irsize)).astype(np.float32)
    x = rng.uniform(-1, 1, (nb, vsize)).astype(np.float32)
    x2 = (rng.uniform(-1, 1, (nb, vsize)) / np.sqrt(irsize)).astype(np.float32)
    c, o = eng.Cldconv(0, irsize, vsize), port.dconv(irsize, vsize)
    c.push_ir(ir)
    o.push_ir(ir)
    y = np.zeros_like(x)
    for t in range(nb):
        assert c.convolution(y[t], x[t], x2[t]) == 0
    want = np.stack([o.convolution(a, b) for a, b in zip(x, x2)])
    assert rel_l2(y, want) < TOL


def test_ragged_irsize_is_defined_behaviour(eng):
    """irsize % vsize != 0: the reference's ring write is broken (Q10). Here the stream semantics simply hold."""
    irsize, vsize, nb = 100, 16, 40
    rng = np.random.default_rng(10)
    ir = rng.standard_normal(irsize).astype(np.float32)
    x = rng.uniform(-1, 1, (nb, vsize)).astype(np.float32)
    c = eng.Cldconv(0, irsize, vsize)
    c.push_ir(ir)
    y = np.zeros_like(x)
    for t in range(nb):
        assert c.convolution(y[t], x[t]) == 0
    exact = np.convolve(x.ravel().astype(np.float64), ir.astype(np.float64))[: nb * vsize]
    assert rel_l2(y.ravel(), np.r_[0.0, exact[:-1]]) < 2e-6


def test_cfg4_64_channels_multiblock_equals_block_by_block(eng, port):
    """BASELINE config 4: 4096 taps, 256-sample blocks, 64 channels. Many blocks in one launch must equal
    the block-by-block stream, and both must match the oracle per channel."""
    irsize, vsize, ch, nb = 4096, 256, 64, 24
    rng = np.random.default_rng(4000)
    ir = (rng.standard_normal((ch, irsize)) / 64).astype(np.float32)
    x = rng.uniform(-1, 1, (ch, nb * vsize)).astype(np.float32)
    a = eng.Cldconv(0, irsize, vsize, channels=ch, max_blocks=nb)
    b = eng.Cldconv(0, irsize, vsize, channels=ch, max_blocks=nb)
    a.push_ir(ir)
    b.push_ir(ir)
    ya = np.zeros_like(x)
    assert a.convolution(ya, x, nblocks=nb) == 0
    yb = np.zeros_like(x)
    for t in range(nb):
        blk = np.ascontiguousarray(x[:, t * vsize:(t + 1) * vsize])
        out = np.zeros_like(blk)
        assert b.convolution(out, blk) == 0
        yb[:, t * vsize:(t + 1) * vsize] = out
    assert rel_l2(ya, yb) < 2e-6
    for c in (0, 31, 63):
        o = port.dconv(irsize, vsize)
        o.push_ir(ir[c])
        want = np.concatenate([o.convolution(x[c, t * vsize:(t + 1) * vsize]) for t in range(nb)])
        assert rel_l2(ya[c], want) < TOL
    # a second multi-block call continues the stream
    x2 = rng.uniform(-1, 1, (ch, nb * vsize)).astype(np.float32)
    ya2 = np.zeros_like(x2)
    assert a.convolution(ya2, x2, nblocks=nb) == 0
    full = np.convolve(np.r_[x[5], x2[5]].astype(np.float64), ir[5].astype(np.float64))[: 2 * nb * vsize]
    assert rel_l2(np.r_[ya[5], ya2[5]], np.r_[0.0, full[:-1]]) < 2e-6


def test_time_varying_multichannel_and_device_api(eng, port):
    import torch

    irsize, vsize, ch, nb = 256, 32, 5, 40
    rng = np.random.default_rng(21)
    ir = (rng.standard_normal((ch, irsize)) / 16).astype(np.float32)
    x = rng.uniform(-1, 1, (nb, ch, vsize)).astype(np.float32)
    x2 = (rng.uniform(-1, 1, (nb, ch, vsize)) / 16).astype(np.float32)
    c = eng.Cldconv(0, irsize, vsize, channels=ch)
    d = eng.Cldconv(0, irsize, vsize, channels=ch)
    assert c.push_ir(ir) == 0
    assert d.push_ir_dev(torch.from_numpy(ir).cuda(), irsize) == 0
    y = np.zeros_like(x)
    dx, dx2 = torch.from_numpy(x).cuda(), torch.from_numpy(x2).cuda()
    dy = torch.zeros_like(dx)
    for t in range(nb):
        assert c.convolution(y[t], x[t], x2[t]) == 0
        assert d.convolution_dev(dy[t], dx[t], dx2[t]) == 0
    torch.cuda.synchronize()
    assert np.array_equal(dy.cpu().numpy(), y)  # host and device entry points: same kernels, same bits
    for k in range(ch):
        o = port.dconv(irsize, vsize)
        o.push_ir(ir[k])
        want = np.stack([o.convolution(x[t, k], x2[t, k]) for t in range(nb)])
        assert rel_l2(y[:, k], want) < TOL
    assert c.reset() == 0
    y2 = np.zeros((ch, vsize), np.float32)
    assert c.convolution(y2, x[0]) == 0
    assert np.all(y2[:, 0] == 0)  # after reset the delay line is empty again (one-sample delay: first output is 0)


def test_argument_errors(eng):
    c = eng.Cldconv(0, 64, 16, channels=2, max_blocks=2)
    x = np.zeros((2, 48), np.float32)
    assert c.convolution(x.copy(), x, nblocks=3) == 6  # more blocks than max_blocks
    assert eng.lib().b2f_dconv_process_host(None, None, None, 1) == 2
    assert eng.Cldconv(0, 0, 16, errs=lambda s, d: None).get_cl_err() == 2
