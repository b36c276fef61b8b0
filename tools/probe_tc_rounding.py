"""Which way does tcgen05.mma round when it adds into its fp32 accumulator?  Long chains (VPS_TC32_GROUP large) of a 1x1
convolution on (a) zero-mean data, (b) all-positive data, no activation; error vs fp64 split by the sign of the result."""
import os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vps_b200 import ops
from vps_b200.layers import empty_nhwc

dev = torch.device("cuda:0")
ops.F32_TC[0] = True
print("GROUP", os.environ.get("VPS_TC32_GROUP"))
for name, fx, fw in (("zero-mean", lambda t: t, lambda t: t), ("positive", lambda t: t.abs(), lambda t: t.abs()),
                     ("x>=0, w mixed", lambda t: t.abs(), lambda t: t)):
    g = torch.Generator().manual_seed(3)
    cin, cout, h, w = 4096, 64, 16, 64
    x = fx(torch.randn(1, h, w, cin, generator=g))
    wt = fw(torch.randn(cout, cin, 1, 1, generator=g)) / cin ** 0.5
    pk = ops.PackedConv(wt.to(dev))
    xd = empty_nhwc(1, h, w, cin, torch.float32, dev); xd.copy_(x)
    y = empty_nhwc(1, h, w, cout, torch.float32, dev)
    ops.conv2d(xd, pk, y, use_tc=True)
    torch.cuda.synchronize()
    ref = F.conv2d(x.permute(0, 3, 1, 2).double(), wt.double())
    d = y.cpu().permute(0, 3, 1, 2).double() - ref
    pos, neg = ref > 0, ref < 0
    ulp = 2.0 ** (torch.floor(torch.log2(ref.abs().clamp_min(1e-30))) - 23)
    print("%-14s mean|ref| %.3f  mean(d)/ulp: ref>0 %+.2f  ref<0 %+.2f   rms(d)/ulp %.2f   max|d|/max|ref| %.2e" %
          (name, float(ref.abs().mean()), float((d / ulp)[pos].mean()) if pos.any() else 0.0,
           float((d / ulp)[neg].mean()) if neg.any() else 0.0, float(((d / ulp) ** 2).mean().sqrt()),
           float(d.abs().max() / ref.abs().max())))
