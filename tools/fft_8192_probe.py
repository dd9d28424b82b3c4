import sys, os, json, torch
sys.path.insert(0, os.getcwd())
import opencl_fft_b200 as eng
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6544.7
N, batch = 8192, 8192
buf = torch.randn(2, batch * N * 2, device="cuda"); out = torch.empty_like(buf)
for sm in (0, 1):
    eng.set_option("fft_sm_8192", sm)
    for fwd in (True, False):
        p = eng.Clcfft(0, N, fwd, max_batch=batch)
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
        gbs = 2 * batch * N * 8 / ms / 1e6
        print(f"one_sm={sm} fwd={fwd}: {ms:.4f} ms {gbs:.0f} GB/s frac {gbs/peak:.3f}", flush=True)
        p.close()
