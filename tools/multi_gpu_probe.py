"""Throughput of the weak pconv workload through the multi-GPU handle (one process, N worker threads) against N
single-device handles driven by N Python threads, host buffers pinned. usage: python tools/multi_gpu_probe.py [N]"""
import os
import sys
import threading
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import opencl_fft_b200 as eng  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else eng.device_count()
ch, cvs, pts, steps = 1024, 480000, 512, 30
rng = np.random.default_rng(0)
ir = (rng.standard_normal((ch, cvs), dtype=np.float32) * 1e-3)
hx = torch.empty(n * ch, pts).pin_memory()
hx.uniform_(-1, 1)
hy = torch.empty(n * ch, pts).pin_memory()
x, y = hx.numpy(), hy.numpy()


def timed(fn):
    for _ in range(3):
        fn()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    return (time.perf_counter() - t0) * 1e3 / steps


singles = []
for g in range(n):
    c = eng.Clpconv(g, cvs, pts, channels=ch)
    assert c.get_cl_err() == 0 and c.push_ir(ir) == 0
    singles.append(c)
print("one device, one handle:", round(timed(lambda: singles[0].convolution(y[:ch], x[:ch])), 3), "ms/step", flush=True)


def threads_step():
    ts = [threading.Thread(target=lambda g=g: singles[g].convolution(y[g * ch:(g + 1) * ch], x[g * ch:(g + 1) * ch])) for g in range(n)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()


print(f"{n} devices, {n} handles, {n} python threads:", round(timed(threads_step), 3), "ms/step", flush=True)
for c in singles:
    c.close()
m = eng.Clpconv(0, cvs, pts, channels=n * ch, devices=list(range(n)))
assert m.get_cl_err() == 0 and m.push_ir(np.tile(ir, (n, 1))) == 0
print(f"{n} devices, one multi handle:", round(timed(lambda: m.convolution(y, x)), 3), "ms/step", flush=True)
