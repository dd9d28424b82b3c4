"""The benchmark grid of the reference's csound/tests.py (lines 5-36), without Csound: mono time-varying
partitioned convolution (what tests.csd:15 runs, `cltvconv`), partition M = 2^{9,11,13,15} x IR length
L = 2^{16..22}, real-time ratio = audio duration / wall time, through the SYNCHRONOUS host API (one call
per partition-sized block, host buffers in and out, as the opcode's aperf would call it). Also the CPU
reference (unmodified reference classes on oracle/minicl, one thread) on a shorter run for comparison.

usage: python tools/rt_ratio_grid.py [--seconds 20] [--cpu]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import opencl_fft_b200 as eng  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=float, default=20.0, help="audio seconds per grid point (the reference used 100)")
ap.add_argument("--cpu", action="store_true", help="also time the CPU reference (oracle) on 2 s of audio")
args = ap.parse_args()
SR = 48000
rng = np.random.default_rng(0)
rows = []
for lm in (9, 11, 13, 15):
    M = 1 << lm
    for ll in range(16, 23):
        L = 1 << ll
        if L < M:
            continue
        conv = eng.Clpconv(0, L, M)
        assert conv.get_cl_err() == 0
        nblocks = max(8, int(args.seconds * SR / M))
        x = rng.uniform(-1, 1, (4, M)).astype(np.float32)
        x2 = (rng.uniform(-1, 1, (4, M)) * 0.01).astype(np.float32)
        y = np.zeros(M, np.float32)
        for i in range(4):
            conv.convolution(y, x[i % 4], x2[i % 4])
        t0 = time.perf_counter()
        for i in range(nblocks):
            conv.convolution(y, x[i % 4], x2[i % 4])
        dt = time.perf_counter() - t0
        row = {"M": M, "L": L, "nparts": L // M, "blocks": nblocks, "us_per_block": round(dt / nblocks * 1e6, 1),
               "rt_ratio": round(nblocks * M / SR / dt, 1)}
        conv.close()
        if args.cpu:
            import oracle

            impl = oracle.best()
            o = impl.pconv(L, M)
            nb = max(2, int(2.0 * SR / M))
            nb = min(nb, 40)
            t0 = time.perf_counter()
            for i in range(nb):
                o.convolution(x[i % 4], x2[i % 4])
            dtc = time.perf_counter() - t0
            row["cpu_reference_rt_ratio"] = round(nb * M / SR / dtc, 2)
            row["cpu_kind"] = impl.kind
        rows.append(row)
        print(row, flush=True)
print(json.dumps(rows))
