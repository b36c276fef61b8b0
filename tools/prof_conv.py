"""Micro-driver for ncu: run ONE representative tensor-core convolution (default: the FPN/TCEA 3x3 256->256
at 256x512, 154.6 GFLOP) a few times.  Usage: python tools/prof_conv.py [cin cout h w k stride] [--iters N] [--tc32]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from vps_b200 import ops  # noqa: E402
from vps_b200.layers import Conv  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
cin, cout, h, w, k, s = (int(v) for v in args[:6]) if len(args) >= 6 else (256, 256, 256, 512, 3, 1)
iters = int(sys.argv[sys.argv.index("--iters") + 1]) if "--iters" in sys.argv else 5
tc32 = "--tc32" in sys.argv                    # fp32 activations through vps_conv2d_tc32 (the parity precision)
ops.F32_TC[0] = tc32
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
x = torch.randn(1, h, w, cin, generator=g).to(dev)
if not tc32:
    x = x.bfloat16()
conv = Conv((torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).to(dev), torch.zeros(cout, device=dev),
            stride=s, pad=k // 2, act=ops.ACT_RELU)
y = conv(x)
torch.cuda.synchronize()
evs = []
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for i in range(iters):
    flush.fill_(i)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    conv(x, y)
    b.record()
    evs.append((a, b))
torch.cuda.synchronize()
ms = sorted(a.elapsed_time(b) for a, b in evs)
fl = 2.0 * y.shape[1] * y.shape[2] * cout * cin * k * k
print("conv %dx%d s%d %d->%d @%dx%d: median %.4f ms  %.1f TFLOP/s" % (k, k, s, cin, cout, y.shape[1], y.shape[2], ms[len(ms) // 2],
                                                                 fl / ms[len(ms) // 2] / 1e9))
