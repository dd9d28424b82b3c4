"""Relative L2 error of the batched transforms against a float64 DFT with the reference's conventions
(forward 1/N; real forward 2/size with packed DC/Nyquist and the conjugated bin size/4)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import opencl_fft_b200 as eng  # noqa: E402

rng = np.random.default_rng(7)
for n in (16, 256, 1024, 4096, 8192, 16384, 32768, 65536):
    batch = 8
    z = (rng.uniform(-1, 1, (batch, n)) + 1j * rng.uniform(-1, 1, (batch, n))).astype(np.complex64)
    y = z.copy()
    assert eng.Clcfft(0, n, True, max_batch=batch).transform(y.reshape(-1)) == 0
    t = np.fft.fft(z.astype(np.complex128), axis=1) / n
    e_c = np.linalg.norm(y - t) / np.linalg.norm(t)
    size = 2 * n
    x = rng.uniform(-1, 1, (batch, size)).astype(np.float32)
    c = np.zeros((batch, n), np.complex64)
    assert eng.Clrfft(0, size, True, max_batch=batch).transform(c.reshape(-1), x.reshape(-1).copy()) == 0
    X = np.fft.rfft(x.astype(np.float64), axis=1)
    w = 2 * X[:, :n] / size
    w[:, 0] = (X[:, 0].real + 1j * X[:, n].real) / size
    w[:, n // 2] = np.conj(w[:, n // 2])
    e_r = np.linalg.norm(c - w) / np.linalg.norm(w)
    back = np.zeros((batch, size), np.float32)
    assert eng.Clrfft(0, size, False, max_batch=batch).transform(c.copy().reshape(-1), back.reshape(-1)) == 0
    e_rt = np.linalg.norm(back - x) / np.linalg.norm(x)
    print(f"N={n:6d}: c2c fwd {e_c:.2e}   r2c (size {size}) {e_r:.2e}   r2c->c2r round trip {e_rt:.2e}", flush=True)
