import sys, os, torch
sys.path.insert(0, os.getcwd())
import opencl_fft_b200 as eng
for mp in (1000, 16, 8, 4):
    eng.set_option("pconv_deep_min_parts", mp)
    for ch, cvs, pts in ((16, 480000, 4096), (32, 480000, 4096), (16, 480000, 2048), (32, 480000, 2048), (1, 96000, 2048), (1, 65536, 2048)):
        c = eng.Clpconv(0, cvs, pts, channels=ch)
        x = torch.randn(ch, pts, device="cuda"); y = torch.empty_like(x)
        for _ in range(5): c.convolution_dev(y, x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): c.convolution_dev(y, x)
        e1.record(); torch.cuda.synchronize()
        print(f"min_parts={mp} ch={ch} pts={pts} nparts={cvs//pts}: {e0.elapsed_time(e1)/50*1e3:.1f} us/step", flush=True)
        c.close()
