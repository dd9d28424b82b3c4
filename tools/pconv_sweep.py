"""Partitioned-convolution sweep: channels x IR length x partition size.
Reports per-block device time, real-time channels @48 kHz, HBM GB/s on algorithmic bytes, and the host-API
(synchronous, pinned host buffers) per-block latency for the small cases."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import opencl_fft_b200 as eng  # noqa: E402

SR = 48000.0
peak = 6544.7
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(p):
    peak = json.load(open(p))["hbm_gbs"]


def run(channels, cvs, pts, steps=100, tv=False):
    conv = eng.Clpconv(0, cvs, pts, channels=channels)
    assert conv.get_cl_err() == 0, conv.get_cl_err()
    ir = torch.randn(channels, cvs, device="cuda") * 0.01
    conv.push_ir_dev(ir, cvs)
    del ir
    x = torch.rand(4, channels, pts, device="cuda") * 2 - 1
    x2 = torch.rand(4, channels, pts, device="cuda") * 0.02
    y = torch.empty(channels, pts, device="cuda")
    f = (lambda i: conv.convolution_dev(y, x[i % 4], x2[i % 4])) if tv else (lambda i: conv.convolution_dev(y, x[i % 4]))
    for i in range(5):
        f(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        f(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    nparts = conv.nparts
    bytes_ = channels * 8 * pts * (2 * nparts + 3)
    row = {"channels": channels, "ir_taps": cvs, "pts": pts, "nparts": nparts, "tv": tv, "us_per_block": round(ms * 1e3, 2),
           "realtime_channels": round(channels * (pts / SR) / (ms * 1e-3), 1), "GBps": round(bytes_ / ms / 1e6, 1),
           "frac_of_measured_peak": round(bytes_ / ms / 1e6 / peak, 3), "state_MB": round(2 * bytes_ / 2e6 / 1, 1)}
    if channels <= 64:
        hx = torch.empty(channels, pts).pin_memory().numpy()
        hy = torch.empty(channels, pts).pin_memory().numpy()
        for _ in range(5):
            conv.convolution(hy, hx)
        t0 = time.perf_counter()
        n = 200
        for _ in range(n):
            conv.convolution(hy, hx)
        row["host_api_us_per_block"] = round((time.perf_counter() - t0) / n * 1e6, 1)
    conv.close()
    print(row, flush=True)
    return row


if len(sys.argv) == 2 and sys.argv[1] == "--cluster-sweep":
    # the strong-scaling regime (1024 channels split over 8 GPUs -> 128 per GPU) and its neighbours: the measured
    # choice of the cluster split (0) against every forced split
    for ch, cvs, pts in ((64, 96000, 512), (128, 96000, 512), (256, 96000, 512), (64, 480000, 512), (128, 480000, 512),
                         (256, 480000, 512), (512, 480000, 512), (256, 480000, 4096), (1024, 480000, 4096),
                         (16, 480000, 2048), (64, 480000, 2048), (128, 480000, 2048), (16, 480000, 4096), (64, 480000, 4096),
                         (128, 480000, 4096)):
        for S in (0, 1, 2, 4, 8):
            eng.set_option("pconv_cluster", S)
            r = run(ch, cvs, pts, steps=30)
    eng.set_option("pconv_cluster", 0)
    sys.exit(0)
if len(sys.argv) == 2 and sys.argv[1] == "--feed-sweep":
    # register-fed against TMA-fed MAC for the partitions wider than the CTA
    for pts in (1024, 2048, 4096):
        for ch in (16, 64, 256, 1024):
            for tma in (0, 1):
                eng.set_option("pconv_tma", tma)
                r = run(ch, 480000, pts, steps=20)
    eng.set_option("pconv_tma", -1)
    sys.exit(0)
if len(sys.argv) == 2 and sys.argv[1] == "--wide":
    # partitions wider than the CTA with the default feed, static and time-varying
    for pts in (2048, 4096):
        for ch in (64, 256, 1024):
            for tv in (False, True):
                run(ch, 480000, pts, steps=20, tv=tv)
    sys.exit(0)
if len(sys.argv) == 4:  # one configuration: channels ir_taps partition
    run(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), steps=30)
    sys.exit(0)
rows = []
for ch in (1, 8, 64, 512, 1024, 4096):
    rows.append(run(ch, 96000, 512))
rows.append(run(1, 96000, 512, tv=True))
rows.append(run(1024, 480000, 512, steps=30))
rows.append(run(1024, 480000, 512, steps=30, tv=True))
for pts in (2048, 4096):
    rows.append(run(256, 480000, pts, steps=30))
# the grid of the reference's csound/tests.py (partition 2^9, 2^11; IR 2^16..2^20), mono, time-varying
for lp in (9, 11):
    for ll in (16, 18, 20):
        rows.append(run(1, 1 << ll, 1 << lp, tv=True))
print(json.dumps(rows))
