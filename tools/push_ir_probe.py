import sys, os, torch
sys.path.insert(0, os.getcwd())
import opencl_fft_b200 as eng
for reg in (0, 1):
  eng.set_option("pconv_push_reg", reg)
  for ch, cvs, pts in ((1024, 480000, 512), (256, 480000, 2048), (1024, 96000, 128), (256, 480000, 4096), (1024, 96000, 64)):
      conv = eng.Clpconv(0, cvs, pts, channels=ch)
      ir = torch.randn(ch, cvs, device="cuda") * 0.01
      for _ in range(2):
          conv.push_ir_dev(ir, cvs)
      torch.cuda.synchronize()
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record()
      for _ in range(5):
          conv.push_ir_dev(ir, cvs)
      e1.record()
      torch.cuda.synchronize()
      ms = e0.elapsed_time(e1) / 5
      nparts = cvs // pts
      gb = ch * nparts * pts * 12 / 1e9
      print(f"push_ir reg={reg} {ch} x {cvs}/{pts}: {ms:.3f} ms, {gb/ms*1e3:.0f} GB/s")
      conv.close()
