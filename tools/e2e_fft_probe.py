import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import opencl_fft_b200 as eng
size, batch = 65536, 1024
f = eng.Clrfft(0, size, True, max_batch=batch)
r = torch.rand(batch, size).pin_memory().numpy()
c = torch.empty(batch, size // 2, dtype=torch.complex64).pin_memory().numpy()
for _ in range(3): assert f.transform(c.reshape(-1), r.reshape(-1)) == 0
t0 = time.perf_counter(); n = 10
for _ in range(n): f.transform(c.reshape(-1), r.reshape(-1))
ms = (time.perf_counter() - t0) / n * 1e3
print("rfft65536 x1024 host call: %.3f ms, %.1f GB/s" % (ms, 2 * batch * size * 4 / ms / 1e6))
