"""Sweep batched FFT throughput over sizes (device resident, CUDA events): GB/s and fraction of the measured
HBM copy peak. usage: python tools/fft_sweep.py [--mb 512]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import opencl_fft_b200 as eng  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--mb", type=int, default=512, help="input megabytes per launch (>> L2)")
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--logn-min", type=int, default=1)
ap.add_argument("--logn-max", type=int, default=16)
args = ap.parse_args()
peak = 6544.7
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(p):
    peak = json.load(open(p))["hbm_gbs"]


def timeit(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


rows = []
total = args.mb << 20
x = torch.randn(2, total // 4, device="cuda")
y = torch.empty_like(x)
for logn in range(args.logn_min, args.logn_max + 1):
    N = 1 << logn
    batch = total // (8 * N)
    for kind in ("c2c", "r2c", "c2r"):
        if kind == "c2c":
            plan = eng.Clcfft(0, N, True, max_batch=batch)
        else:
            plan = eng.Clrfft(0, 2 * N, kind == "r2c", max_batch=batch)
        assert plan.get_error() == 0
        k = [0]

        def fn():
            k[0] ^= 1
            assert plan.transform_dev(x[k[0]], y[k[0]], batch) == 0

        ms = timeit(fn, args.iters)
        gbs = 2 * total / ms / 1e6
        rows.append({"kind": kind, "N_complex": N, "batch": batch, "ms": round(ms, 4), "GBps": round(gbs, 1),
                     "frac_of_measured_peak": round(gbs / peak, 3)})
        print(rows[-1], flush=True)
        plan.close()
print(json.dumps(rows))
