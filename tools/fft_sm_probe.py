"""A/B of the N = 2^15 paths (mode 0: four-step launch pair, mode 1: one-SM kernel of fft_sm.cuh, selected with the
fft_sm_min_batch option): results against float64 for every transform of a batch, device-resident timing, and with
--sweep the small-batch crossover between the two.
usage: python tools/fft_sm_probe.py [--batch 1024] [--modes 0,1,2]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import opencl_fft_b200 as eng  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--modes", default="0,1")
ap.add_argument("--check-batch", type=int, default=300)
ap.add_argument("--skip-check", action="store_true")
ap.add_argument("--sweep", action="store_true", help="time r2c at batches 1..592 on both paths")
ap.add_argument("--kinds", default="r2c,c2c,c2r")
args = ap.parse_args()
peak = 6544.7
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(p):
    peak = json.load(open(p))["hbm_gbs"]
size = 65536


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


rng = np.random.default_rng(0)
cb = args.check_batch
xr = rng.uniform(-1, 1, (cb, size)).astype(np.float32)
X = np.fft.rfft(xr.astype(np.float64), axis=1)
want_r = 2 * X[:, : size // 2] / size
want_r[:, 0] = (X[:, 0].real + 1j * X[:, size // 2].real) / size
want_r[:, size // 4] = np.conj(want_r[:, size // 4])
xc = (rng.uniform(-1, 1, (cb, size // 2)) + 1j * rng.uniform(-1, 1, (cb, size // 2))).astype(np.complex64)
want_c = np.fft.fft(xc.astype(np.complex128), axis=1) / (size // 2)
want_ci = np.fft.ifft(xc.astype(np.complex128), axis=1) * (size // 2)

total = args.batch * size * 4
buf = torch.randn(2, total // 4, device="cuda")
out = torch.empty_like(buf)
for mode in [int(m) for m in args.modes.split(",")]:
    eng.set_option("fft_sm_min_batch", 1 if mode else 0)
    res = {"mode": mode}
    if not args.skip_check:
        # correctness, every transform, device API
        f = eng.Clrfft(0, size, True, max_batch=cb)
        d = torch.from_numpy(xr).cuda()
        o = torch.empty(cb, size, device="cuda")
        assert f.transform_dev(d, o, cb) == 0
        torch.cuda.synchronize()
        got = o.cpu().numpy().view(np.complex64)
        err = np.linalg.norm(got - want_r, axis=1) / np.linalg.norm(want_r, axis=1)
        res["r2c_err_max"] = float(err.max())
        res["r2c_bad"] = int((err > 2e-6).sum())
        assert f.transform_dev(d, d, cb) == 0  # in place
        torch.cuda.synchronize()
        res["r2c_inplace_same"] = bool(np.array_equal(d.cpu().numpy().view(np.complex64), got))
        # inverse real transform of every spectrum gives the signal back
        iv = eng.Clrfft(0, size, False, max_batch=cb)
        ds = torch.from_numpy(got.view(np.float32).copy()).cuda()
        ob = torch.empty(cb, size, device="cuda")
        assert iv.transform_dev(ds, ob, cb) == 0
        torch.cuda.synchronize()
        back = ob.cpu().numpy()
        err = np.linalg.norm(back - xr, axis=1) / np.linalg.norm(xr, axis=1)
        res["c2r_err_max"] = float(err.max())
        res["c2r_bad"] = int((err > 2e-6).sum())
        iv.close()
        f.close()
        for fwd, want, key in ((True, want_c, "c2c_fwd"), (False, want_ci, "c2c_inv")):
            pl = eng.Clcfft(0, size // 2, fwd, max_batch=cb)
            d = torch.from_numpy(xc.view(np.float32)).cuda()
            o = torch.empty_like(d)
            assert pl.transform_dev(d, o, cb) == 0
            torch.cuda.synchronize()
            got = o.cpu().numpy().view(np.complex64)
            err = np.linalg.norm(got - want, axis=1) / np.linalg.norm(want, axis=1)
            res[key + "_err_max"] = float(err.max())
            res[key + "_bad"] = int((err > 2e-6).sum())
            pl.close()
    # timing
    for kind in args.kinds.split(","):
        batch = args.batch if kind == "r2c" else args.batch
        plan = eng.Clrfft(0, size, kind == "r2c", max_batch=batch) if kind != "c2c" else eng.Clcfft(0, size // 2, True, max_batch=batch)
        k = [0]

        def fn():
            k[0] ^= 1
            assert plan.transform_dev(buf[k[0]], out[k[0]], batch) == 0

        ms = timeit(fn, args.iters)
        gbs = 2 * total / ms / 1e6
        res[kind + "_ms"] = round(ms, 4)
        res[kind + "_GBps"] = round(gbs, 1)
        res[kind + "_frac"] = round(gbs / peak, 3)
        plan.close()
    print(json.dumps(res), flush=True)

if args.sweep:
    for b in (1, 8, 32, 64, 96, 128, 148, 222, 296, 444, 592):
        row = {"batch": b}
        for mode in (0, 1):
            eng.set_option("fft_sm_min_batch", 1 if mode else 0)
            plan = eng.Clrfft(0, size, True, max_batch=b)
            k = [0]

            def fn():
                k[0] ^= 1
                assert plan.transform_dev(buf[k[0]], out[k[0]], b) == 0

            row["four_step_us" if mode == 0 else "one_sm_us"] = round(timeit(fn, 50) * 1e3, 2)
            plan.close()
        print(json.dumps(row), flush=True)
