"""Launch ONE hot kernel family a couple of times at its BASELINE shape, for `ncu --set full`:
    python tools/profile_all.py cfft1024 | rfft4096 | rfft65536 | pconv5b | pconv_general | dconv4 | push_ir | pconv_mono_deep | pconv_general_mono
(see profiles/r01_ncu_all_kernels.md for the command line used)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import opencl_fft_b200 as eng  # noqa: E402

wl = sys.argv[1]
reps = 2
if wl == "cfft1024":
    p = eng.Clcfft(0, 1024, True, max_batch=65536)
    x = torch.randn(65536, 1024, 2, device="cuda")
    y = torch.empty_like(x)
    for _ in range(reps):
        p.transform_dev(x, y, 65536)
elif wl == "rfft4096":
    f, i = eng.Clrfft(0, 4096, True, max_batch=32768), eng.Clrfft(0, 4096, False, max_batch=32768)
    r = torch.rand(32768, 4096, device="cuda")
    c = torch.empty_like(r)
    for _ in range(reps):
        f.transform_dev(r, c, 32768)
        i.transform_dev(c, c, 32768)
elif wl == "rfft65536":
    f5 = eng.Clrfft(0, 65536, True, max_batch=1024)
    r5 = torch.rand(1024, 65536, device="cuda")
    c5 = torch.empty_like(r5)
    for _ in range(reps):
        f5.transform_dev(r5, c5, 1024)
elif wl == "pconv5b":  # config 5b, 256-channel slice (1.97 GB of state) to keep the capture short
    conv = eng.Clpconv(0, 480000, 512, channels=256)
    ir = torch.randn(256, 480000, device="cuda") * 0.01
    conv.push_ir_dev(ir, 480000)
    xb = torch.rand(256, 512, device="cuda")
    yb = torch.empty_like(xb)
    for _ in range(reps):
        conv.convolution_dev(yb, xb)
elif wl == "pconv_general":  # 64 channels x 2^20-tap IR, 8192-sample partitions
    cg = eng.Clpconv(0, 1 << 20, 8192, channels=64)
    irg = torch.randn(64, 1 << 20, device="cuda") * 0.01
    cg.push_ir_dev(irg, 1 << 20)
    xg = torch.rand(64, 8192, device="cuda")
    yg = torch.empty_like(xg)
    for _ in range(reps):
        cg.convolution_dev(yg, xg)
elif wl == "push_ir":  # config 5b's set-up, 256-channel slice
    conv = eng.Clpconv(0, 480000, 512, channels=256)
    ir = torch.randn(256, 480000, device="cuda") * 0.01
    for _ in range(reps):
        conv.push_ir_dev(ir, 480000)
elif wl == "pconv_mono_deep":  # one channel, 4M-tap IR, 2048-sample partitions: cluster of 16, 32 KB TMA stages
    cm = eng.Clpconv(0, 1 << 22, 2048, channels=1)
    xm = torch.rand(1, 2048, device="cuda")
    ym = torch.empty_like(xm)
    for _ in range(reps + 2):
        cm.convolution_dev(ym, xm, xm)
elif wl == "pconv_general_mono":  # one channel, 4M-tap IR, 8192-sample partitions: fused frames / split MAC / inverse + OLA
    cm = eng.Clpconv(0, 1 << 22, 8192, channels=1)
    xm = torch.rand(1, 8192, device="cuda")
    ym = torch.empty_like(xm)
    for _ in range(reps + 2):
        cm.convolution_dev(ym, xm, xm)
elif wl == "dconv4":
    d = eng.Cldconv(0, 4096, 256, channels=64)
    d.push_ir_dev(torch.randn(64, 4096, device="cuda") / 64, 4096)
    xd = torch.rand(64, 375 * 256, device="cuda")
    yd = torch.empty_like(xd)
    for _ in range(reps):
        d.convolution_dev(yd, xd, nblocks=375)
torch.cuda.synchronize()
print("done", wl)
