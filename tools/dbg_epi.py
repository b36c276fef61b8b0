"""debug: identity 1x1 tc32 convolution -- shows how the epilogue permutes (pixel, channel)"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vps_b200 import ops
from vps_b200.layers import empty_nhwc
dev = torch.device("cuda:0")
ops.F32_TC[0] = True
for (h, w, c, k) in [(16, 32, 64, 1), (16, 32, 64, 3), (8, 8, 128, 1)]:
    x = empty_nhwc(1, h, w, c, torch.float32, dev)
    pix = torch.arange(h * w, device=dev).view(1, h, w, 1).float()
    ch = torch.arange(c, device=dev).view(1, 1, 1, c).float()
    x.copy_(pix + ch / 256.0)
    wt = torch.zeros(c, c, k, k, device=dev)
    wt[torch.arange(c), torch.arange(c), k // 2, k // 2] = 1.0
    pk = ops.PackedConv(wt, None)
    y = torch.full((1, h, w, c), float("nan"), device=dev)
    ops.conv2d(x, pk, y, stride=1, pad=k // 2, use_tc=True)
    torch.cuda.synchronize()
    bad = (y - x).abs() > 1e-3
    bad |= torch.isnan(y)
    print("case", (h, w, c, k), "bad", int(bad.sum()), "of", bad.numel(), "nan", int(torch.isnan(y).sum()))
    if bad.any():
        idx = bad.nonzero()[:12]
        for i in idx:
            _, yy, xx, cc = i.tolist()
            v = float(y[0, yy, xx, cc])
            print("  at pixel (%d,%d)=%d ch %d: got %.4f = pixel %d ch %.0f" % (yy, xx, yy * w + xx, cc, v, int(v), (v - int(v)) * 256))
