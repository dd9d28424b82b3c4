"""Out-of-bounds write detection without compute-sanitizer (closed on this pool, in round 2 too): every device-pointer entry
point is run on buffers carved out of a larger allocation whose surroundings hold a sentinel; the sentinel
must be intact afterwards, for odd batch / channel counts that leave partially filled CTAs and tiles."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
SENT = 12345.678


class Guarded:
    """`n` float32 elements with `pad` sentinel elements on both sides (256-byte aligned payload)."""

    def __init__(self, n, pad=4096):
        self.n, self.pad = n, pad
        self.buf = torch.full((n + 2 * pad,), SENT, device="cuda", dtype=torch.float32)
        self.view = self.buf[pad:pad + n]

    def intact(self):
        return bool((self.buf[: self.pad] == SENT).all() and (self.buf[self.pad + self.n:] == SENT).all())


@pytest.mark.parametrize("logn,batch", [(4, 7), (6, 37), (9, 5), (10, 13), (12, 3), (13, 3), (14, 1), (15, 3), (16, 1)])
def test_fft_device_entry_points_stay_in_bounds(eng, logn, batch):
    n = 1 << logn
    for real in (False, True):
        for fwd in (True, False):
            plan = eng.Clrfft(0, 2 * n, fwd, max_batch=batch) if real else eng.Clcfft(0, n, fwd, max_batch=batch)
            assert plan.get_error() == 0
            src, dst = Guarded(batch * n * 2), Guarded(batch * n * 2)
            src.view.uniform_(-1, 1)
            assert plan.transform_dev(src.view, dst.view, batch) == 0
            torch.cuda.synchronize()
            assert src.intact() and dst.intact()
            assert torch.isfinite(dst.view).all() and not (dst.view == SENT).any()
            assert plan.transform_dev(src.view, src.view, batch) == 0  # in place
            torch.cuda.synchronize()
            assert src.intact()


@pytest.mark.parametrize("batch", [3, 149])
def test_one_sm_fft_stays_in_bounds(eng, options, batch):
    """The one-SM kernel (TMA-staged input, persistent CTAs): forced for a batch smaller than the grid and for one that
    makes a single CTA take a second transform; complex both ways, real forward and inverse, in and out of place."""
    options("fft_sm_min_batch", 1)
    n = 32768
    for real in (False, True):
        for fwd in (True, False):
            plan = eng.Clrfft(0, 2 * n, fwd, max_batch=batch) if real else eng.Clcfft(0, n, fwd, max_batch=batch)
            assert plan.get_error() == 0
            src, dst = Guarded(batch * n * 2), Guarded(batch * n * 2)
            src.view.uniform_(-1, 1)
            assert plan.transform_dev(src.view, dst.view, batch) == 0
            torch.cuda.synchronize()
            assert src.intact() and dst.intact()
            assert torch.isfinite(dst.view).all() and not (dst.view == SENT).any()
            assert plan.transform_dev(src.view, src.view, batch) == 0  # in place
            torch.cuda.synchronize()
            assert src.intact()


@pytest.mark.parametrize("pts,nparts,channels", [(16, 3, 5), (512, 7, 3), (512, 5, 301), (1024, 4, 65), (2048, 3, 2), (4096, 3, 3), (8192, 2, 3)])
def test_pconv_device_entry_points_stay_in_bounds(eng, pts, nparts, channels):
    cvs = pts * nparts + 3
    c = eng.Clpconv(0, cvs, pts, channels=channels)
    ir = Guarded(channels * cvs)
    ir.view.normal_(0, 0.1)
    assert c.push_ir_dev(ir.view, cvs) == 0
    x, x2, y = Guarded(channels * pts), Guarded(channels * pts), Guarded(channels * pts)
    x.view.uniform_(-1, 1)
    x2.view.uniform_(-0.1, 0.1)
    for t in range(nparts + 2):
        assert c.convolution_dev(y.view, x.view) == 0
        assert c.convolution_dev(y.view, x.view, x2.view) == 0
    torch.cuda.synchronize()
    assert ir.intact() and x.intact() and x2.intact() and y.intact()
    assert torch.isfinite(y.view).all() and not (y.view == SENT).any()


@pytest.mark.parametrize("opts,pts,nparts,channels", [({"pconv_cluster": 16}, 512, 40, 1), ({"pconv_cluster": 2}, 2048, 200, 2),
                                                      ({"pconv_cluster": 2}, 4096, 200, 1), ({}, 8192, 40, 1), ({}, 16384, 5, 3),
                                                      ({"pconv_general_fused": 0}, 8192, 5, 3), ({}, 32768, 3, 2)])
def test_pconv_few_channel_paths_stay_in_bounds(eng, options, opts, pts, nparts, channels):
    """The launch shapes added for a few channels with long IRs: cluster of 16, 32 KB TMA stages, general path with the
    partitions split over CTAs, fused frames / inverse + overlap-add, both time-varying inputs in one batch."""
    for k, v in opts.items():
        options(k, v)
    cvs = pts * nparts + 5
    c = eng.Clpconv(0, cvs, pts, channels=channels)
    assert c.get_cl_err() == 0
    ir = Guarded(channels * cvs)
    ir.view.normal_(0, 0.1)
    assert c.push_ir_dev(ir.view, cvs) == 0
    x, x2, y = Guarded(channels * pts), Guarded(channels * pts), Guarded(channels * pts)
    x.view.uniform_(-1, 1)
    x2.view.uniform_(-0.1, 0.1)
    for t in range(4):
        assert c.convolution_dev(y.view, x.view) == 0
        assert c.convolution_dev(y.view, x.view, x2.view) == 0
    torch.cuda.synchronize()
    assert ir.intact() and x.intact() and x2.intact() and y.intact()
    assert torch.isfinite(y.view).all() and not (y.view == SENT).any()


@pytest.mark.parametrize("batch", [5, 593])
def test_one_sm_fft_8192_stays_in_bounds(eng, options, batch):
    """Complex N = 8192 on the one-SM kernel (four transforms per unit): a last unit with one transform in it."""
    options("fft_sm_min_batch", 1)
    n = 8192
    for fwd in (True, False, None):  # None: the inverse 16384-point real transform, on the same kernel
        plan = eng.Clcfft(0, n, fwd, max_batch=batch) if fwd is not None else eng.Clrfft(0, 2 * n, False, max_batch=batch)
        assert plan.get_error() == 0
        src, dst = Guarded(batch * n * 2), Guarded(batch * n * 2)
        src.view.uniform_(-1, 1)
        assert plan.transform_dev(src.view, dst.view, batch) == 0
        torch.cuda.synchronize()
        assert src.intact() and dst.intact()
        assert torch.isfinite(dst.view).all() and not (dst.view == SENT).any()
        assert plan.transform_dev(src.view, src.view, batch) == 0  # in place
        torch.cuda.synchronize()
        assert src.intact()


@pytest.mark.parametrize("irsize,vsize,channels,nblocks", [(4096, 256, 3, 1), (4096, 256, 3, 7), (100, 16, 5, 3), (33, 1, 2, 5), (512, 64, 300, 1)])
def test_dconv_device_entry_points_stay_in_bounds(eng, irsize, vsize, channels, nblocks):
    d = eng.Cldconv(0, irsize, vsize, channels=channels)
    ir = Guarded(channels * irsize)
    ir.view.normal_(0, 0.1)
    assert d.push_ir_dev(ir.view, irsize) == 0
    x, y = Guarded(channels * nblocks * vsize), Guarded(channels * nblocks * vsize)
    x.view.uniform_(-1, 1)
    for _ in range(3):
        assert d.convolution_dev(y.view, x.view, nblocks=nblocks) == 0
    xb, x2, yb = Guarded(channels * vsize), Guarded(channels * vsize), Guarded(channels * vsize)
    xb.view.uniform_(-1, 1)
    x2.view.uniform_(-0.1, 0.1)
    assert d.convolution_dev(yb.view, xb.view, x2.view) == 0
    torch.cuda.synchronize()
    for g in (ir, x, y, xb, x2, yb):
        assert g.intact()
    assert torch.isfinite(y.view).all() and not (y.view == SENT).any()


def test_misaligned_device_pointers_are_rejected_not_faulted(eng):
    """FFT / partitioned-convolution device pointers must be 16-byte aligned (include/b200fft.h): a float-aligned
    view is refused with a status code instead of a misaligned-address fault that would poison the context."""
    buf = torch.zeros(4 * 1024 + 8, device="cuda", dtype=torch.float32)
    plan = eng.Clcfft(0, 1024, True, max_batch=1)
    assert plan.transform_dev(buf[1:2049], buf[2052:4100], 1) != 0
    assert plan.transform_dev(buf[0:2048], buf[2048:4096], 1) == 0
    c = eng.Clpconv(0, 64, 16, channels=2)
    assert c.convolution_dev(buf[1:33], buf[64:96]) != 0
    assert c.convolution_dev(buf[0:32], buf[64:96]) == 0
    torch.cuda.synchronize()
