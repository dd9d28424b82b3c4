"""Experiment: the batched real FFT reading its input from and writing its result to caller-pinned HOST memory in one
launch (device-pointer entry point on the UVA aliases), against the staged host call."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import opencl_fft_b200 as eng  # noqa: E402

size, batch = int(sys.argv[1]) if len(sys.argv) > 1 else 65536, int(sys.argv[2]) if len(sys.argv) > 2 else 1024
f = eng.Clrfft(0, size, True, max_batch=batch)
r = torch.rand(batch, size).pin_memory()
c = torch.zeros(batch, size // 2, dtype=torch.complex64).pin_memory()
c2 = torch.zeros(batch, size // 2, dtype=torch.complex64).pin_memory()
L = eng.lib()
st = torch.cuda.current_stream().cuda_stream


def direct():
    rc = L.b2f_rfft_exec_dev(f._h, C.c_void_p(r.data_ptr()), C.c_void_p(c.data_ptr()), batch, C.c_void_p(st))
    assert rc == 0, rc
    torch.cuda.synchronize()


def staged():
    assert f.transform(c2.numpy().reshape(-1), r.numpy().reshape(-1)) == 0


for name, fn in (("in place over PCIe", direct), ("staged host call", staged)) * 2:
    for _ in range(2):
        fn()
    t0 = time.perf_counter()
    n = 8
    for _ in range(n):
        fn()
    ms = (time.perf_counter() - t0) / n * 1e3
    print(f"{name:20s} {ms:.3f} ms  {2 * batch * size * 4 / ms / 1e6:.1f} GB/s", flush=True)
print("same bits:", bool(torch.equal(c, c2)))
