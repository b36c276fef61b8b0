"""tc32 (tensor cores) vs CUDA-core fp32 on the thin layers of the path (tiny cout or cin): which one should a layer use?"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vps_b200 import ops
from vps_b200.layers import empty_nhwc

SH = [(82, 16, 1024, 2048, 3, 1), (16, 2, 1024, 2048, 3, 1), (194, 2, 256, 512, 3, 1), (386, 2, 128, 256, 3, 1), (770, 2, 64, 128, 3, 1),
      (1026, 2, 32, 64, 3, 1), (1024, 2, 16, 32, 3, 1), (256, 18, 256, 512, 3, 1), (6, 64, 1024, 2048, 3, 1), (11, 64, 1024, 2048, 3, 1),
      (48, 64, 512, 1024, 4, 1), (12, 64, 512, 1024, 4, 1), (64, 32, 256, 512, 3, 1), (339, 64, 256, 512, 3, 1), (32, 2, 256, 512, 3, 1),
      (162, 32, 512, 1024, 3, 1), (64, 64, 512, 1024, 3, 2)]
dev = torch.device("cuda:0")
for cin, cout, h, w, k, s in SH:
    g = torch.Generator().manual_seed(1)
    wt = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).to(dev)
    pk = ops.PackedConv(wt, torch.zeros(cout, device=dev))
    x = empty_nhwc(1, h, w, cin, torch.float32, dev); x.normal_()
    pad = k // 2
    oh, ow = (h + 2 * pad - k) // s + 1, (w + 2 * pad - k) // s + 1
    y = empty_nhwc(1, oh, ow, cout, torch.float32, dev)
    res = {}
    for name, tc in (("tc32", True), ("simt", False)):
        ops.F32_TC[0] = tc
        f = lambda: ops.conv2d(x, pk, y, stride=s, pad=pad, act=ops.ACT_LRELU, use_tc=tc)
        for _ in range(2): f()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5): f()
        b.record(); torch.cuda.synchronize()
        res[name] = a.elapsed_time(b) / 5
    print("%dx%d s%d %4d->%-3d @%dx%-5d tc32 %7.3f ms  simt %7.3f ms  -> %s" % (k, k, s, cin, cout, oh, ow, res["tc32"], res["simt"],
                                                                             "simt" if res["simt"] < res["tc32"] else "tc32"), flush=True)
