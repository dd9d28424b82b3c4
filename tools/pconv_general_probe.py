"""General path (pts >= 8192) in the throughput regime: us per block and fraction of the measured HBM peak of a whole
block step (all launches), device-resident. usage: python tools/pconv_general_probe.py"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import opencl_fft_b200 as eng  # noqa: E402
peak = 6544.7
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(p):
    peak = json.load(open(p))["hbm_gbs"]
for ch, cvs, pts in ((64, 1 << 20, 8192), (256, 1 << 20, 8192), (1024, 480000, 8192), (64, 1 << 21, 16384), (256, 1 << 20, 16384),
                     (64, 1 << 21, 32768), (256, 1 << 20, 32768)):
    c = eng.Clpconv(0, cvs, pts, channels=ch)
    x = torch.randn(ch, pts, device="cuda")
    y = torch.empty_like(x)
    for tv in (False, True):
        for _ in range(4):
            c.convolution_dev(y, x, x if tv else None)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        e0.record()
        for _ in range(n):
            c.convolution_dev(y, x, x if tv else None)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        nparts = cvs // pts
        gb = ch * 8 * pts * (2 * nparts + 3) / 1e9
        print(f"ch={ch} pts={pts} nparts={nparts} tv={int(tv)}: {ms*1e3:.1f} us/block {gb/ms*1e3:.0f} GB/s frac {gb/ms*1e3/peak:.2f}", flush=True)
    c.close()
