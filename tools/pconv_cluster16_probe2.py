import sys, os, torch
sys.path.insert(0, os.getcwd())
import opencl_fft_b200 as eng
for ch, cvs, pts in ((16, 480000, 512), (32, 480000, 512), (16, 480000, 1024), (32, 480000, 1024), (16, 1 << 22, 512), (12, 96000, 512)):
    for S in (0, 8, 16):
        eng.set_option("pconv_cluster", S)
        c = eng.Clpconv(0, cvs, pts, channels=ch)
        x = torch.randn(ch, pts, device="cuda"); y = torch.empty_like(x)
        for _ in range(5): c.convolution_dev(y, x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): c.convolution_dev(y, x)
        e1.record(); torch.cuda.synchronize()
        print(f"ch={ch} pts={pts} nparts={cvs//pts} cluster={S}: {e0.elapsed_time(e1)/50*1e3:.1f} us/step", flush=True)
        c.close()
