import sys, os, time, numpy as np, torch
sys.path.insert(0, os.getcwd())
import opencl_fft_b200 as eng
for S in (4, 8, 16):
    eng.set_option("pconv_cluster", S)
    c = eng.Clpconv(0, 96000, 512)
    rng = np.random.default_rng(0)
    c.push_ir((rng.standard_normal(96000) * 0.01).astype(np.float32))
    x = torch.rand(1, 512, device="cuda"); y = torch.empty_like(x)
    for _ in range(10): c.convolution_dev(y, x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200): c.convolution_dev(y, x)
    e1.record(); torch.cuda.synchronize()
    xin, yout = rng.uniform(-1, 1, 512).astype(np.float32), np.zeros(512, np.float32)
    for _ in range(20): c.convolution(yout, xin)
    t0 = time.perf_counter()
    for _ in range(300): c.convolution(yout, xin)
    us = (time.perf_counter() - t0) / 300 * 1e6
    print(f"cluster {S}: device {e0.elapsed_time(e1)/200*1e3:.1f} us/block (back-to-back launches), host API {us:.1f} us")
    c.close()
