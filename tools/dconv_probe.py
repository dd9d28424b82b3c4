"""dconv throughput probe: config 4 (4096 taps, 256-sample blocks, 64 channels), many blocks per launch."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import opencl_fft_b200 as eng  # noqa: E402

irsize, vsize, ch = 4096, 256, 64
nblocks = int(sys.argv[1]) if len(sys.argv) > 1 else 375
conv = eng.Cldconv(0, irsize, vsize, channels=ch, max_blocks=1)
ir = torch.randn(ch, irsize, device="cuda") / 64
assert conv.push_ir_dev(ir, irsize) == 0
x = torch.rand(ch, nblocks * vsize, device="cuda") * 2 - 1
y = torch.empty_like(x)
for _ in range(3):
    conv.convolution_dev(y, x, nblocks=nblocks)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
it = 10
e0.record()
for _ in range(it):
    conv.convolution_dev(y, x, nblocks=nblocks)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / it
flop = 2.0 * irsize * vsize * nblocks * ch
print(f"dconv {nblocks} blocks: {ms:.3f} ms, {flop / ms / 1e9:.1f} TFLOP/s")
