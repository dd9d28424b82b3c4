import sys, os, torch, json
sys.path.insert(0, os.getcwd())
import opencl_fft_b200 as eng
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6544.7
for deep in (0, 1):
    eng.set_option("pconv_deep_ring", deep)
    for ch, cvs, pts in ((1, 1 << 22, 512), (1, 1 << 22, 2048), (1, 1 << 20, 1024), (4, 1 << 21, 512), (16, 480000, 512), (16, 480000, 1024), (16, 480000, 2048), (16, 480000, 4096), (64, 480000, 2048), (32, 480000, 4096)):
        c = eng.Clpconv(0, cvs, pts, channels=ch)
        x = torch.randn(ch, pts, device="cuda"); y = torch.empty_like(x)
        for _ in range(5): c.convolution_dev(y, x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 50
        e0.record()
        for _ in range(n): c.convolution_dev(y, x)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        nparts = cvs // pts
        gb = ch * 8 * pts * (2 * nparts + 3) / 1e9
        print(f"deep={deep} ch={ch} cvs={cvs} pts={pts} nparts={nparts}: {ms*1e3:.1f} us/step {gb/ms*1e3:.0f} GB/s frac {gb/ms*1e3/peak:.2f}", flush=True)
        c.close()
