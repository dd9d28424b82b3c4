import sys, os, json, numpy as np, torch
sys.path.insert(0, os.getcwd())
import opencl_fft_b200 as eng
peak = 6544.7
size = 16384
rng = np.random.default_rng(0)
for batch in (1, 3, 5, 593):
    x = rng.uniform(-1, 1, (batch, size)).astype(np.float32)
    eng.set_option("fft_sm_min_batch", 0)
    f = eng.Clrfft(0, size, True, max_batch=batch)
    spec = np.zeros((batch, size // 2), np.complex64)
    assert f.transform(spec.reshape(-1), x.reshape(-1).copy()) == 0
    res = {}
    for forced in (0, 1):
        eng.set_option("fft_sm_min_batch", forced)
        iv = eng.Clrfft(0, size, False, max_batch=batch)
        c = spec.copy(); r = np.zeros((batch, size), np.float32)
        assert iv.transform(c.reshape(-1), r.reshape(-1)) == 0
        res[forced] = r
        err = np.linalg.norm(r - x, axis=1) / np.linalg.norm(x, axis=1)
        print("batch", batch, "forced", forced, "roundtrip err max", float(err.max()))
    print("  one-SM vs reg kernel rel", float(np.linalg.norm(res[1] - res[0]) / np.linalg.norm(res[0])))
eng.set_option("fft_sm_min_batch", 96)
batch = 8192
buf = torch.randn(2, batch * size, device="cuda"); out = torch.empty_like(buf)
for sm in (0, 1):
    eng.set_option("fft_sm_8192", sm)
    p = eng.Clrfft(0, size, False, max_batch=batch)
    k = [0]
    def fn():
        k[0] ^= 1
        assert p.transform_dev(buf[k[0]], out[k[0]], batch) == 0
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    gbs = 2 * batch * size * 4 / ms / 1e6
    print(f"c2r 16384-point real x {batch}: one_sm={sm}: {ms:.4f} ms {gbs:.0f} GB/s frac {gbs/peak:.3f}", flush=True)
