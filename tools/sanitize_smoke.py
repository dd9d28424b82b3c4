"""Small invocation of every kernel family, for compute-sanitizer (memcheck / racecheck, one tool per run):
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py
Sizes are kept tiny: the tools slow kernels down 10-100x."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import opencl_fft_b200 as eng  # noqa: E402

rng = np.random.default_rng(0)


def crand(*shape):
    return (rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1, 1, shape)).astype(np.complex64)


# batched FFTs of every schedule class (single pass, 2, 3, 4 passes), forward and inverse, real and complex
for n, batch in ((16, 5), (64, 5), (512, 3), (1024, 3), (4096, 2), (8192, 1)):
    x = crand(batch, n)
    for fwd in (True, False):
        p = eng.Clcfft(0, n, fwd, max_batch=batch)
        y = x.copy()
        assert p.transform(y.reshape(-1)) == 0
    r = rng.uniform(-1, 1, (batch, 2 * n)).astype(np.float32)
    c = np.zeros((batch, n), np.complex64)
    assert eng.Clrfft(0, 2 * n, True, max_batch=batch).transform(c.reshape(-1), r.reshape(-1)) == 0
    assert eng.Clrfft(0, 2 * n, False, max_batch=batch).transform(c.reshape(-1), r.reshape(-1)) == 0
# N = 2^15: the four-step launch pair (small batches) and the one-SM kernel (forced here for a batch of 3)
for min_batch in (0, 1):
    eng.set_option("fft_sm_min_batch", min_batch)
    x = crand(3, 32768)
    assert eng.Clcfft(0, 32768, True, max_batch=3).transform(x.reshape(-1)) == 0
    assert eng.Clcfft(0, 32768, False, max_batch=3).transform(x.reshape(-1)) == 0
    r = rng.uniform(-1, 1, (3, 65536)).astype(np.float32)
    c = np.zeros((3, 32768), np.complex64)
    assert eng.Clrfft(0, 65536, True, max_batch=3).transform(c.reshape(-1), r.reshape(-1)) == 0
    assert eng.Clrfft(0, 65536, False, max_batch=3).transform(c.reshape(-1), r.reshape(-1)) == 0
eng.set_option("fft_sm_min_batch", 96)
# partitioned convolution: fused kernel (1 CTA, clusters of 2..8, register- and TMA-fed), general path, time-varying
for tma in (0, 1):
    eng.set_option("pconv_tma", tma)
    for pts, nparts, ch in ((64, 5, 1), (512, 9, 3), (512, 5, 200), (1024, 4, 70), (8192, 3, 2)):
        cvs = pts * nparts
        c = eng.Clpconv(0, cvs, pts, channels=ch)
        assert c.push_ir((rng.standard_normal((ch, cvs)) * 0.1).astype(np.float32)) == 0
        y = np.zeros((ch, pts), np.float32)
        for t in range(nparts + 2):
            x = rng.uniform(-1, 1, (ch, pts)).astype(np.float32)
            assert c.convolution(y, x) == 0
            assert c.convolution(y, x, x * 0.1) == 0
eng.set_option("pconv_tma", -1)
# added in the last third of round 2: cluster of 16, 32 KB TMA stages (pts 2048 / 4096, >= 96 partitions per CTA), general path with
# fused frames / inverse + overlap-add (pts 8192, 16384), partitions split over CTAs + partial sums, unfused general path
# (pts 32768, both time-varying inputs in one batch), push_ir on the register-level transform with an odd stride
for opts, pts, nparts, ch in (({"pconv_cluster": 16}, 512, 40, 1), ({"pconv_cluster": 2}, 2048, 200, 1), ({"pconv_cluster": 2}, 4096, 200, 1),
                              ({}, 8192, 40, 1), ({}, 16384, 5, 2), ({"pconv_general_fused": 0}, 8192, 5, 2), ({}, 32768, 3, 1)):
    for k, v in opts.items():
        eng.set_option(k, v)
    cvs = pts * nparts
    c = eng.Clpconv(0, cvs, pts, channels=ch)
    assert c.get_cl_err() == 0
    assert c.push_ir((rng.standard_normal((ch, cvs)) * 0.1).astype(np.float32)) == 0
    y = np.zeros((ch, pts), np.float32)
    for t in range(3):
        x = rng.uniform(-1, 1, (ch, pts)).astype(np.float32)
        assert c.convolution(y, x) == 0
        assert c.convolution(y, x, x * 0.1) == 0
    for k in opts:
        eng.set_option(k, {"pconv_cluster": 0, "pconv_general_fused": 1}[k])
import torch  # noqa: E402

cw = eng.Clpconv(0, 6 * 512, 512, channels=3)
assert cw.push_ir_dev(torch.randn(3, 6 * 512 + 1, device="cuda"), 6 * 512 + 1) == 0
torch.cuda.synchronize()
# direct convolution: single block (cluster tap split), multi-block (16 outputs per thread), ragged sizes, time-varying
for irsize, vsize, ch, nb in ((4096, 256, 2, 1), (4096, 256, 2, 6), (100, 16, 3, 1), (64, 1, 2, 3)):
    d = eng.Cldconv(0, irsize, vsize, channels=ch, max_blocks=nb)
    assert d.push_ir((rng.standard_normal((ch, irsize)) / 8).astype(np.float32)) == 0
    x = rng.uniform(-1, 1, (ch, nb * vsize)).astype(np.float32)
    y = np.zeros_like(x)
    assert d.convolution(y, x, nblocks=nb) == 0
    if nb == 1:
        assert d.convolution(y, x, x * 0.1) == 0
print("sanitize smoke done")
