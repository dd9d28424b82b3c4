"""GPU parity of the partitioned convolution (Clpconv) through the C ABI against the oracle, the golden
vectors of the real reference, and float64 models. Streaming over enough blocks to wrap both rings."""
import numpy as np
import pytest

from conftest import TOL, rel_l2

pytestmark = pytest.mark.gpu


def run_stream(conv, x, x2=None):
    """x: [nb][channels][pts] -> y same shape"""
    y = np.zeros_like(x)
    for t in range(x.shape[0]):
        rc = conv.convolution(y[t], x[t]) if x2 is None else conv.convolution(y[t], x[t], x2[t])
        assert rc == 0
    return y


def test_golden_cfg3(eng, golden):
    g = golden
    c = eng.Clpconv(0, 96000, 512)
    assert c.get_cl_err() == 0 and c.nparts == 187
    assert c.push_ir(g["pconv_cfg3_ir"]) == 0
    y = run_stream(c, g["pconv_cfg3_in"][:, None, :])[:, 0]
    assert rel_l2(y, g["pconv_cfg3_out"]) < TOL


def test_golden_small_static_and_time_varying(eng, golden):
    g = golden
    c = eng.Clpconv(0, 1000, 64)
    assert c.nparts == 15
    assert c.push_ir(g["pconv_small_ir"]) == 0
    y = run_stream(c, g["pconv_small_in"][:, None, :])[:, 0]
    assert rel_l2(y, g["pconv_small_out"]) < TOL
    c = eng.Clpconv(0, 1000, 64)
    y = run_stream(c, g["pconv_small_in"][:, None, :], g["pconv_small_in2"][:, None, :])[:, 0]
    assert rel_l2(y, g["pconv_small_tv_out"]) < TOL


@pytest.mark.parametrize("pts", [2, 4, 16, 32, 64, 256, 512, 1024, 2048, 4096])
def test_static_vs_oracle_all_partition_sizes(eng, port, pts):
    nparts = 5
    cvs = nparts * pts + pts // 2  # truncating division must drop the half partition
    nb = 2 * nparts + 3
    rng = np.random.default_rng(pts)
    ir = rng.standard_normal(cvs).astype(np.float32)
    x = rng.uniform(-1, 1, (nb, 1, pts)).astype(np.float32)
    c = eng.Clpconv(0, cvs, pts)
    assert c.get_cl_err() == 0 and c.nparts == nparts
    o = port.pconv(cvs, pts)
    assert c.push_ir(ir) == 0
    o.push_ir(ir)
    # white box: the IR spectra ring has the reference's frame order and values
    assert rel_l2(c.read_spectra(2), o.spec2()) < TOL
    y = run_stream(c, x)[:, 0]
    want = np.stack([o.convolution(b[0]) for b in x])
    assert rel_l2(y, want) < TOL
    assert rel_l2(c.read_spectra(1), o.spec1()) < TOL  # frequency-domain delay line


@pytest.mark.parametrize("pts,nparts", [(16, 1), (16, 2), (64, 7), (512, 9)])
def test_time_varying_vs_oracle(eng, port, pts, nparts):
    cvs = nparts * pts
    nb = 3 * nparts + 2
    rng = np.random.default_rng(pts + nparts)
    x = rng.uniform(-1, 1, (nb, 1, pts)).astype(np.float32)
    x2 = rng.uniform(-1, 1, (nb, 1, pts)).astype(np.float32)
    c = eng.Clpconv(0, cvs, pts)
    o = port.pconv(cvs, pts)
    y = run_stream(c, x, x2)[:, 0]
    want = np.stack([o.convolution(a[0], b[0]) for a, b in zip(x, x2)])
    assert rel_l2(y, want) < TOL
    assert rel_l2(c.read_spectra(2), o.spec2()) < TOL


def test_mixed_static_then_time_varying(eng, port):
    """push_ir, a few static blocks, then time-varying blocks re-recording the IR ring (Q12)."""
    cvs, pts = 640, 64
    rng = np.random.default_rng(12)
    ir = rng.standard_normal(cvs).astype(np.float32)
    x = rng.uniform(-1, 1, (30, 1, pts)).astype(np.float32)
    x2 = rng.uniform(-1, 1, (30, 1, pts)).astype(np.float32)
    c, o = eng.Clpconv(0, cvs, pts), port.pconv(cvs, pts)
    c.push_ir(ir)
    o.push_ir(ir)
    got, want = [], []
    for t in range(30):
        y = np.zeros((1, pts), np.float32)
        if t < 8 or t >= 20:
            assert c.convolution(y, x[t]) == 0
            want.append(o.convolution(x[t, 0]))
        else:
            assert c.convolution(y, x[t], x2[t]) == 0
            want.append(o.convolution(x[t, 0], x2[t, 0]))
        got.append(y[0])
    assert rel_l2(np.stack(got), np.stack(want)) < TOL


@pytest.mark.parametrize("channels", [1, 3, 40, 300])
def test_multichannel_matches_per_channel_oracle(eng, port, channels):
    """channels selects the cluster split: 1 -> 8 CTAs per channel, 40 -> 8, 300 -> 1."""
    cvs, pts, nb = 3200, 128, 30
    rng = np.random.default_rng(channels)
    ir = rng.standard_normal((channels, cvs)).astype(np.float32)
    x = rng.uniform(-1, 1, (nb, channels, pts)).astype(np.float32)
    c = eng.Clpconv(0, cvs, pts, channels=channels)
    assert c.push_ir(ir) == 0
    y = run_stream(c, x)
    for ch in sorted({0, channels // 2, channels - 1}):
        o = port.pconv(cvs, pts)
        o.push_ir(ir[ch])
        want = np.stack([o.convolution(x[t, ch]) for t in range(nb)])
        assert rel_l2(y[:, ch], want) < TOL
    # determinism: a second object fed the same stream gives the same bits
    c2 = eng.Clpconv(0, cvs, pts, channels=channels)
    c2.push_ir(ir)
    assert np.array_equal(run_stream(c2, x), y)


def test_float64_model_with_quirk_q5(eng):
    """Against float64: exact linear convolution except half-weight DC/Nyquist per frame product."""
    cvs, pts, nb = 4096, 512, 20
    rng = np.random.default_rng(5)
    ir = rng.standard_normal(cvs).astype(np.float32)
    x = rng.uniform(-1, 1, (nb, 1, pts)).astype(np.float32)
    c = eng.Clpconv(0, cvs, pts)
    c.push_ir(ir)
    y = run_stream(c, x)[:, 0].ravel()
    H = [np.fft.rfft(np.r_[ir[i * pts:(i + 1) * pts].astype(np.float64), np.zeros(pts)]) for i in range(cvs // pts)]
    X = [np.fft.rfft(np.r_[b[0].astype(np.float64), np.zeros(pts)]) for b in x]
    out, tail = [], np.zeros(pts)
    for t in range(nb):
        Y = sum(X[t - a] * H[a] for a in range(len(H)) if t - a >= 0)
        Y[0] *= 0.5
        Y[-1] *= 0.5
        yy = np.fft.irfft(Y)
        out.append(yy[:pts] + tail)
        tail = yy[pts:]
    assert rel_l2(y, np.concatenate(out)) < 2e-6


def test_cfg5_shape_properties(eng, port):
    """BASELINE config 5b shape (480000-tap IR -> 937 partitions of 512), a 16-channel slice: linearity and
    time invariance hold at full size; one channel is also checked block by block against the oracle."""
    cvs, pts, ch, nb = 480000, 512, 16, 6
    rng = np.random.default_rng(7000)
    n = np.arange(cvs)
    ir = (rng.standard_normal((ch, cvs)) * np.exp(-6.9078 * n / cvs)).astype(np.float32)
    ir /= np.linalg.norm(ir, axis=1, keepdims=True)
    xa = rng.uniform(-1, 1, (nb, ch, pts)).astype(np.float32)
    xb = rng.uniform(-1, 1, (nb, ch, pts)).astype(np.float32)

    def run(x):
        c = eng.Clpconv(0, cvs, pts, channels=ch)
        assert c.get_cl_err() == 0 and c.nparts == 937
        assert c.push_ir(ir) == 0
        return run_stream(c, x)

    ya, yb, yab = run(xa), run(xb), run(xa + 0.5 * xb)
    assert rel_l2(yab, ya + 0.5 * yb) < 2e-6
    # time invariance: delaying the input by one block delays the output by one block
    xd = np.concatenate([np.zeros((1, ch, pts), np.float32), xa[:-1]])
    yd = run(xd)
    assert rel_l2(yd[1:], ya[:-1]) < 2e-6
    o = port.pconv(cvs, pts)
    o.push_ir(ir[3])
    want = np.stack([o.convolution(xa[t, 3]) for t in range(nb)])
    assert rel_l2(ya[:, 3], want) < TOL


def test_reset_and_device_api(eng):
    import torch

    cvs, pts, ch, nb = 2048, 256, 8, 12
    rng = np.random.default_rng(3)
    ir = rng.standard_normal((ch, cvs)).astype(np.float32)
    x = rng.uniform(-1, 1, (nb, ch, pts)).astype(np.float32)
    c = eng.Clpconv(0, cvs, pts, channels=ch)
    c.push_ir(ir)
    y = run_stream(c, x)
    assert c.reset() == 0
    d = eng.Clpconv(0, cvs, pts, channels=ch)
    d_ir = torch.from_numpy(ir).cuda()
    assert d.push_ir_dev(d_ir, cvs) == 0
    d_x = torch.from_numpy(x).cuda()
    d_y = torch.zeros_like(d_x)
    for t in range(nb):
        assert c.convolution_dev(d_y[t], d_x[t]) == 0
    torch.cuda.synchronize()
    assert np.array_equal(d_y.cpu().numpy(), y)  # after reset, same stream -> same bits
    d_y2 = torch.zeros_like(d_x)
    for t in range(nb):
        assert d.convolution_dev(d_y2[t], d_x[t]) == 0
    torch.cuda.synchronize()
    assert np.array_equal(d_y2.cpu().numpy(), y)  # device-side push_ir == host-side push_ir


@pytest.mark.parametrize("pts", [8192, 16384, 32768])
def test_large_partitions_general_path(eng, port, pts):
    """Partition sizes above 4096 (the upper half of the reference's csound/tests.py sweep) run on the general
    path (batched real-FFT plans + MAC / overlap-add kernels) instead of the fused kernel: same semantics."""
    nparts = 3
    cvs = nparts * pts + 100
    nb = 2 * nparts + 2
    rng = np.random.default_rng(pts)
    ir = (rng.standard_normal((2, cvs)) * 0.05).astype(np.float32)
    x = rng.uniform(-1, 1, (nb, 2, pts)).astype(np.float32)
    x2 = (rng.uniform(-1, 1, (nb, 2, pts)) * 0.05).astype(np.float32)
    c = eng.Clpconv(0, cvs, pts, channels=2)
    assert c.get_cl_err() == 0 and c.nparts == nparts
    assert c.push_ir(ir) == 0
    y = run_stream(c, x)
    for ch in range(2):
        o = port.pconv(cvs, pts)
        o.push_ir(ir[ch])
        assert rel_l2(c.read_spectra(2, ch), o.spec2()) < TOL
        want = np.stack([o.convolution(x[t, ch]) for t in range(nb)])
        assert rel_l2(y[:, ch], want) < TOL
    # time-varying on a fresh object
    c = eng.Clpconv(0, cvs, pts, channels=2)
    y = run_stream(c, x, x2)
    o = port.pconv(cvs, pts)
    want = np.stack([o.convolution(x[t, 1], x2[t, 1]) for t in range(nb)])
    assert rel_l2(y[:, 1], want) < TOL


def test_partition_size_limit(eng):
    c = eng.Clpconv(0, 1 << 18, 1 << 16, errs=lambda s, d: None)
    assert c.get_cl_err() == 3  # frames above 32768 complex points are not implemented


def test_time_varying_multichannel(eng, port):
    cvs, pts, ch, nb = 1024, 128, 6, 20  # 8 partitions, rings wrap twice
    rng = np.random.default_rng(31)
    x = rng.uniform(-1, 1, (nb, ch, pts)).astype(np.float32)
    x2 = (rng.uniform(-1, 1, (nb, ch, pts)) * 0.1).astype(np.float32)
    c = eng.Clpconv(0, cvs, pts, channels=ch)
    y = run_stream(c, x, x2)
    for k in (0, 3, 5):
        o = port.pconv(cvs, pts)
        want = np.stack([o.convolution(x[t, k], x2[t, k]) for t in range(nb)])
        assert rel_l2(y[:, k], want) < TOL


@pytest.mark.parametrize("tma", [0, 1])
@pytest.mark.parametrize("pts,channels", [(64, 3), (512, 70), (1024, 70), (2048, 2), (4096, 3), (8192, 2)])
def test_both_mac_feeds(eng, port, options, tma, pts, channels):
    """The spectral multiply-accumulate exists twice: fed by 128-bit register loads and fed by the TMA engine
    (cp.async.bulk into an mbarrier ring). The library picks by shape; option pconv_tma forces either. Both must
    match the oracle, static and time-varying, including ring wrap."""
    options("pconv_tma", tma)  # copied by the handles created below
    nparts = 5
    cvs, nb = nparts * pts, 2 * nparts + 2
    rng = np.random.default_rng(pts + channels)
    ir = (rng.standard_normal((channels, cvs)) * 0.1).astype(np.float32)
    x = rng.uniform(-1, 1, (nb, channels, pts)).astype(np.float32)
    x2 = (rng.uniform(-1, 1, (nb, channels, pts)) * 0.1).astype(np.float32)
    c = eng.Clpconv(0, cvs, pts, channels=channels)
    assert c.push_ir(ir) == 0
    y = run_stream(c, x)
    ctv = eng.Clpconv(0, cvs, pts, channels=channels)
    ytv = run_stream(ctv, x, x2)
    for k in sorted({0, channels - 1}):
        o = port.pconv(cvs, pts)
        o.push_ir(ir[k])
        assert rel_l2(y[:, k], np.stack([o.convolution(x[t, k]) for t in range(nb)])) < TOL
        o = port.pconv(cvs, pts)
        assert rel_l2(ytv[:, k], np.stack([o.convolution(x[t, k], x2[t, k]) for t in range(nb)])) < TOL


def test_pipelined_host_call_equals_device_path(eng, options):
    """Many channels with a block above 1 MB: the synchronous host call runs the two halves of the channels on two
    streams (upload of one half overlapping the kernel of the other). Same bits as the single-stream form and as
    the device-pointer entry point, over enough blocks to wrap the delay line, odd channel count included."""
    import torch

    pts, nparts, channels, nb = 512, 5, 1027, 8
    rng = np.random.default_rng(11)
    ir = (rng.standard_normal((channels, pts * nparts)) * 0.05).astype(np.float32)
    x = rng.uniform(-1, 1, (nb, channels, pts)).astype(np.float32)

    def host_run():
        c = eng.Clpconv(0, pts * nparts, pts, channels=channels)
        assert c.push_ir(ir.reshape(-1)) == 0
        return run_stream(c, x)

    y_pipe = host_run()
    options("pconv_pipeline", 0)
    y_single = host_run()
    assert np.array_equal(y_pipe, y_single)
    c = eng.Clpconv(0, pts * nparts, pts, channels=channels)
    assert c.push_ir(ir.reshape(-1)) == 0
    d_y = torch.empty(channels, pts, device="cuda")
    for t in range(nb):
        assert c.convolution_dev(d_y, torch.from_numpy(x[t]).cuda()) == 0
        torch.cuda.synchronize()
        assert np.array_equal(d_y.cpu().numpy(), y_pipe[t])


@pytest.mark.parametrize("tv", [False, True])
def test_caller_pinned_buffers_run_in_place(eng, options, tv):
    """Host calls above the bounce-buffer size on buffers the CALLER has page-locked: one launch reads the input blocks
    from and writes the output blocks to that memory directly (option pinned_direct). Same bits as the staged form
    (pageable buffers, and pinned ones with the option off), static and time-varying, delay line wrapped."""
    import torch

    pts, nparts, channels, nb = 512, 5, 600, 12
    rng = np.random.default_rng(12)
    ir = (rng.standard_normal((channels, pts * nparts)) * 0.05).astype(np.float32)
    x = rng.uniform(-1, 1, (nb, channels, pts)).astype(np.float32)
    x2 = (rng.standard_normal((nb, channels, pts)) * 0.05).astype(np.float32)

    def run(pinned):
        c = eng.Clpconv(0, pts * nparts, pts, channels=channels)
        if not tv:
            assert c.push_ir(ir.reshape(-1)) == 0
        if pinned:
            hx, hx2, hy = (torch.empty(channels, pts).pin_memory().numpy() for _ in range(3))
        else:
            hx, hx2, hy = (np.empty((channels, pts), np.float32) for _ in range(3))
        ys = []
        for t in range(nb):
            hx[:], hx2[:] = x[t], x2[t]
            assert (c.convolution(hy, hx, hx2) if tv else c.convolution(hy, hx)) == 0
            ys.append(hy.copy())
        return np.stack(ys)

    y_direct = run(True)
    y_pageable = run(False)
    options("pinned_direct", 0)
    y_staged = run(True)
    assert np.array_equal(y_direct, y_pageable) and np.array_equal(y_direct, y_staged)
    assert np.abs(y_direct).max() > 0


def test_benched_launch_configuration_vs_oracle(eng, port):
    """The exact configuration bench.py times (BASELINE configs[4]): 1024 channels x 480000 taps x 512-sample
    partitions = 937 partitions, which selects 2-CTA clusters. Four scattered channels against the oracle over three
    blocks; IRs and inputs are generated on the device as the bench does and copied back for the oracle."""
    import torch

    channels, cvs, pts, nb = 1024, 480000, 512, 3
    c = eng.Clpconv(0, cvs, pts, channels=channels)
    assert c.get_cl_err() == 0 and c.nparts == 937
    g = torch.Generator(device="cuda").manual_seed(7000)
    n = torch.arange(cvs, device="cuda", dtype=torch.float32)
    ir = torch.randn(channels, cvs, generator=g, device="cuda") * torch.exp(-6.9078 * n / cvs)
    ir /= ir.norm(dim=1, keepdim=True)
    assert c.push_ir_dev(ir, cvs) == 0
    x = torch.rand(nb, channels, pts, generator=g, device="cuda") * 2 - 1
    y = torch.empty(nb, channels, pts, device="cuda")
    for t in range(nb):
        assert c.convolution_dev(y[t], x[t]) == 0
    torch.cuda.synchronize()
    for k in (0, 341, 682, 1023):
        o = port.pconv(cvs, pts)
        o.push_ir(ir[k].cpu().numpy())
        want = np.stack([o.convolution(x[t, k].cpu().numpy()) for t in range(nb)])
        assert rel_l2(y[:, k].cpu().numpy(), want) < TOL, k


@pytest.mark.parametrize("cluster", [1, 2, 4, 8, 16])
def test_ring_wrap_for_every_cluster_size(eng, port, options, cluster):
    """40 partitions split over clusters of 1 / 2 / 4 / 8 / 16 CTAs (option pconv_cluster), 2 * nparts + 2 blocks so that the
    delay line wraps twice, static and time-varying, against the oracle."""
    options("pconv_cluster", cluster)
    pts, nparts, channels = 512, 40, 3
    cvs, nb = nparts * pts, 2 * nparts + 2
    rng = np.random.default_rng(cluster)
    ir = (rng.standard_normal((channels, cvs)) * 0.05).astype(np.float32)
    x = rng.uniform(-1, 1, (nb, channels, pts)).astype(np.float32)
    x2 = (rng.uniform(-1, 1, (nb, channels, pts)) * 0.05).astype(np.float32)
    c = eng.Clpconv(0, cvs, pts, channels=channels)
    assert c.get_cl_err() == 0 and c.push_ir(ir) == 0
    y = run_stream(c, x)
    ctv = eng.Clpconv(0, cvs, pts, channels=channels)
    ytv = run_stream(ctv, x, x2)
    for k in (0, channels - 1):
        o = port.pconv(cvs, pts)
        o.push_ir(ir[k])
        assert rel_l2(y[:, k], np.stack([o.convolution(x[t, k]) for t in range(nb)])) < TOL
        o = port.pconv(cvs, pts)
        assert rel_l2(ytv[:, k], np.stack([o.convolution(x[t, k], x2[t, k]) for t in range(nb)])) < TOL


@pytest.mark.parametrize("pts", [2048, 4096])
def test_deep_ring_for_long_streams(eng, port, options, pts):
    """Few channels, long IR: CTAs that stream >= 96 partitions each take the TMA feed with 32 KB stages (option
    pconv_deep_ring). Same partition order per bin, so the same bits as the 4 KB-slice feed; and the oracle's values,
    static and time-varying, over a full wrap of the delay line."""
    nparts, channels = 200, 2
    cvs, nb = nparts * pts, nparts + 3
    rng = np.random.default_rng(pts + 1)
    ir = (rng.standard_normal((channels, cvs)) * 0.02).astype(np.float32)
    x = rng.uniform(-1, 1, (nb, channels, pts)).astype(np.float32)
    x2 = (rng.uniform(-1, 1, (nb, channels, pts)) * 0.02).astype(np.float32)
    options("pconv_cluster", 2)
    res = {}
    for deep in (1, 0):
        options("pconv_deep_ring", deep)
        c = eng.Clpconv(0, cvs, pts, channels=channels)
        assert c.get_cl_err() == 0 and c.push_ir(ir) == 0
        ctv = eng.Clpconv(0, cvs, pts, channels=channels)
        res[deep] = (run_stream(c, x), run_stream(ctv, x, x2))
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
    o = port.pconv(cvs, pts)
    o.push_ir(ir[1])
    assert rel_l2(res[1][0][:, 1], np.stack([o.convolution(x[t, 1]) for t in range(nb)])) < TOL
    o = port.pconv(cvs, pts)
    assert rel_l2(res[1][1][:, 1], np.stack([o.convolution(x[t, 1], x2[t, 1]) for t in range(nb)])) < TOL


@pytest.mark.parametrize("nparts", [128, 1600])
def test_mono_long_ir_4096_sample_partitions(eng, port, nparts):
    """One channel, 4096-sample partitions, long IR: the measured choice is a cluster of 16 CTAs (non-portable size; the
    portable 8 where the device cannot co-schedule 16) and, from 96 partitions per CTA, 32 KB TMA stages at 198 KB of
    shared memory per CTA. The oracle's values, static and time-varying."""
    pts = 4096
    cvs, nb = nparts * pts, 5
    rng = np.random.default_rng(nparts)
    ir = (rng.standard_normal((1, cvs)) * 0.01).astype(np.float32)
    x = rng.uniform(-1, 1, (nb, 1, pts)).astype(np.float32)
    x2 = (rng.uniform(-1, 1, (nb, 1, pts)) * 0.01).astype(np.float32)
    c = eng.Clpconv(0, cvs, pts)
    assert c.get_cl_err() == 0 and c.push_ir(ir) == 0
    y = run_stream(c, x)
    ctv = eng.Clpconv(0, cvs, pts)
    ytv = run_stream(ctv, x, x2)
    o = port.pconv(cvs, pts)
    o.push_ir(ir[0])
    assert rel_l2(y[:, 0], np.stack([o.convolution(x[t, 0]) for t in range(nb)])) < TOL
    o = port.pconv(cvs, pts)
    assert rel_l2(ytv[:, 0], np.stack([o.convolution(x[t, 0], x2[t, 0]) for t in range(nb)])) < TOL


def test_invalid_cluster_option_is_rejected_at_create(eng, options):
    options("pconv_cluster", 3)
    assert eng.Clpconv(0, 4096, 512, uData=1).get_cl_err() == 2
    options("pconv_cluster", 8)
    assert eng.Clpconv(0, 4 * 512, 512, uData=1).get_cl_err() == 2  # more CTAs than partitions


@pytest.mark.parametrize("pts,channels", [(8192, 2), (32768, 1), (32768, 100)])  # 100 channels: the one-SM FFT kernel (programmatic dependent launch) inside the graph
def test_graph_replay_equals_stream_launches(eng, port, options, pts, channels):
    """General path (pts >= 8192, 7-9 launches per block): the host call replays the block as one CUDA graph over
    device-resident ring positions. Same bits as the plain stream launches (option graph = 0), static and
    time-varying, over enough blocks to wrap the rings; and the oracle's values."""
    nparts = 3
    cvs, nb = nparts * pts, 2 * nparts + 2
    rng = np.random.default_rng(pts)
    ir = (rng.standard_normal((channels, cvs)) * 0.05).astype(np.float32)
    x = rng.uniform(-1, 1, (nb, channels, pts)).astype(np.float32)
    x2 = (rng.uniform(-1, 1, (nb, channels, pts)) * 0.05).astype(np.float32)
    res = {}
    for graph in (1, 0):
        options("graph", graph)
        c = eng.Clpconv(0, cvs, pts, channels=channels)
        assert c.get_cl_err() == 0 and c.push_ir(ir) == 0
        ys = run_stream(c, x)
        c.reset()  # the graph survives a reset: the ring positions it reads are device state
        ys2 = run_stream(c, x)
        assert np.array_equal(ys, ys2)
        ctv = eng.Clpconv(0, cvs, pts, channels=channels)
        res[graph] = (ys, run_stream(ctv, x, x2))
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
    o = port.pconv(cvs, pts)
    o.push_ir(ir[0])
    assert rel_l2(res[1][0][:, 0], np.stack([o.convolution(x[t, 0]) for t in range(nb)])) < TOL
    o = port.pconv(cvs, pts)
    assert rel_l2(res[1][1][:, 0], np.stack([o.convolution(x[t, 0], x2[t, 0]) for t in range(nb)])) < TOL


@pytest.mark.parametrize("ksplit", [0, 3, -1])
def test_general_path_partition_split(eng, port, options, ksplit):
    """pts >= 8192 with few channels: the partitions of the multiply-accumulate are split over several CTAs per tile
    and the partial sums added by a second launch (option pconv_ksplit; 0 = measured choice, here 5). The oracle's
    values, static and time-varying, over more than a full wrap of the delay line."""
    options("pconv_ksplit", ksplit)
    pts, nparts, channels = 8192, 40, 1
    cvs, nb = nparts * pts, nparts + 3
    rng = np.random.default_rng(7)
    ir = (rng.standard_normal((channels, cvs)) * 0.02).astype(np.float32)
    x = rng.uniform(-1, 1, (nb, channels, pts)).astype(np.float32)
    x2 = (rng.uniform(-1, 1, (nb, channels, pts)) * 0.02).astype(np.float32)
    c = eng.Clpconv(0, cvs, pts, channels=channels)
    assert c.get_cl_err() == 0 and c.push_ir(ir) == 0
    y = run_stream(c, x)
    ctv = eng.Clpconv(0, cvs, pts, channels=channels)
    ytv = run_stream(ctv, x, x2)
    o = port.pconv(cvs, pts)
    o.push_ir(ir[0])
    assert rel_l2(y[:, 0], np.stack([o.convolution(x[t, 0]) for t in range(nb)])) < TOL
    o = port.pconv(cvs, pts)
    assert rel_l2(ytv[:, 0], np.stack([o.convolution(x[t, 0], x2[t, 0]) for t in range(nb)])) < TOL


@pytest.mark.parametrize("pts,channels", [(8192, 3), (16384, 2)])
def test_general_path_fused_launches_same_bits(eng, port, options, pts, channels):
    """pts 8192 / 16384: the block's frames (pad + rFFT + frame copy, per input) and its tail (inverse rFFT + overlap-add
    + ring advance) run as one launch each on the register-level transforms, and push_ir as one launch for all
    partitions. Same arithmetic as the separate launches (option pconv_general_fused = 0): same bits; and the oracle's
    values."""
    nparts = 5
    cvs, nb = nparts * pts, 2 * nparts + 2
    rng = np.random.default_rng(pts + channels)
    ir = (rng.standard_normal((channels, cvs)) * 0.05).astype(np.float32)
    x = rng.uniform(-1, 1, (nb, channels, pts)).astype(np.float32)
    x2 = (rng.uniform(-1, 1, (nb, channels, pts)) * 0.05).astype(np.float32)
    res = {}
    for fused in (1, 0):
        options("pconv_general_fused", fused)
        c = eng.Clpconv(0, cvs, pts, channels=channels)
        assert c.get_cl_err() == 0 and c.push_ir(ir) == 0
        spec = c.read_spectra(2, channels - 1)
        ctv = eng.Clpconv(0, cvs, pts, channels=channels)
        res[fused] = (run_stream(c, x), run_stream(ctv, x, x2), spec)
    assert np.array_equal(res[0][2], res[1][2])
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
    k = channels - 1
    o = port.pconv(cvs, pts)
    o.push_ir(ir[k])
    assert rel_l2(res[1][0][:, k], np.stack([o.convolution(x[t, k]) for t in range(nb)])) < TOL
    o = port.pconv(cvs, pts)
    assert rel_l2(res[1][1][:, k], np.stack([o.convolution(x[t, k], x2[t, k]) for t in range(nb)])) < TOL


@pytest.mark.parametrize("pts", [32, 512, 8192])
def test_push_ir_dev_with_an_odd_channel_stride(eng, pts):
    """push_ir_dev takes any ir_stride >= the IR length: with an odd stride every second channel's rows are only 4-byte
    aligned and the transform reads them with scalar loads. Same spectra, bit for bit, as the contiguous push (pts 32: the
    step kernel's frame routine; 512: the register-level transform; 8192: the general path's single-launch push)."""
    import torch

    nparts, channels = 6, 3
    cvs = nparts * pts
    rng = np.random.default_rng(pts)
    ir = (rng.standard_normal((channels, cvs)) * 0.05).astype(np.float32)
    wide = np.zeros((channels, cvs + 1), np.float32)
    wide[:, :cvs] = ir
    a, b = eng.Clpconv(0, cvs, pts, channels=channels), eng.Clpconv(0, cvs, pts, channels=channels)
    assert a.push_ir_dev(torch.from_numpy(ir).cuda(), cvs) == 0
    assert b.push_ir_dev(torch.from_numpy(wide).cuda(), cvs + 1) == 0
    torch.cuda.synchronize()
    for k in range(channels):
        assert np.array_equal(a.read_spectra(2, k), b.read_spectra(2, k)), k


def test_python_binding_rejects_short_buffers(eng):
    """The C entry points read and write channels * pts floats behind the pointers they are given: the binding checks
    the element counts first (B2F_ERR_INVALID_VALUE = 2) instead of letting the engine run off the arrays."""
    c = eng.Clpconv(0, 4 * 64, 64, channels=3)
    assert c.push_ir(np.zeros(3 * 4 * 64 - 1, np.float32)) == 2       # one float short
    assert c.push_ir(np.zeros((2, 4 * 64), np.float32)) == 2          # a row missing
    assert c.push_ir(np.zeros((3, 4 * 64), np.float32)) == 0
    good, short = np.zeros((3, 64), np.float32), np.zeros((3, 63), np.float32)
    assert c.convolution(short, good) == 2 and c.convolution(good, short) == 2
    assert c.convolution(good, good, short) == 2
    assert c.convolution(good, good) == 0 and c.convolution(good, good, good) == 0
    d = eng.Cldconv(0, 32, 8, channels=2, max_blocks=4)
    assert d.push_ir(np.zeros(63, np.float32)) == 2 and d.push_ir(np.zeros((2, 32), np.float32)) == 0
    assert d.convolution(np.zeros((2, 16), np.float32), np.zeros((2, 32), np.float32), nblocks=4) == 2
    assert d.convolution(np.zeros((2, 32), np.float32), np.zeros((2, 32), np.float32), nblocks=4) == 0
    r = eng.Clrfft(0, 64, True)
    assert r.transform(np.zeros(32, np.complex64), np.zeros(63, np.float32)) == 2
    assert r.transform(np.zeros(32, np.complex64), np.zeros(64, np.float32)) == 0
