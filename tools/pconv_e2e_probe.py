"""Host-call time of the partitioned convolution at BASELINE configuration 5b (1024 channels x 480000 taps x 512) on
caller-pinned buffers: the in-place form (option pinned_direct = 1) against the staged two-stream pipeline (0), and
pageable buffers, in one process."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import opencl_fft_b200 as eng  # noqa: E402

ch, cvs, pts = 1024, 480000, 512
ir = torch.randn(ch, cvs, device="cuda") * 0.01
for label, direct, pinned in (("pinned, in place", 1, True), ("pinned, staged pipeline", 0, True), ("pageable", 1, False),
                              ("pinned, in place", 1, True), ("pinned, staged pipeline", 0, True)):
    eng.set_option("pinned_direct", direct)
    conv = eng.Clpconv(0, cvs, pts, channels=ch)
    conv.push_ir_dev(ir, cvs)
    if pinned:
        hx, hy = torch.rand(ch, pts).pin_memory().numpy(), torch.empty(ch, pts).pin_memory().numpy()
    else:
        hx, hy = np.random.rand(ch, pts).astype(np.float32), np.empty((ch, pts), np.float32)
    for _ in range(5):
        conv.convolution(hy, hx)
    t0 = time.perf_counter()
    n = 100
    for _ in range(n):
        conv.convolution(hy, hx)
    ms = (time.perf_counter() - t0) / n * 1e3
    print(f"{label:26s} {ms:.4f} ms/block  {ch * (pts / 48000.0) / (ms * 1e-3):9.1f} real-time channels", flush=True)
    conv.close()
eng.set_option("pinned_direct", 1)
