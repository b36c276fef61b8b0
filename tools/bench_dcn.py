"""Time the UPSNetFPN deformable convolutions: fused vps_deform_conv_tc vs vps_deform_im2col + 1x1 GEMM."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from vps_b200 import ops  # noqa: E402
from vps_b200.layers import Conv  # noqa: E402

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=5):
    fn()
    torch.cuda.synchronize()
    evs = []
    for i in range(iters):
        flush.fill_(i)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    return sorted(a.elapsed_time(b) for a, b in evs)[len(evs) // 2]


tot_f = tot_u = 0.0
for (h, w) in [(256, 512), (128, 256), (64, 128), (32, 64)]:
    for (ci, co) in [(256, 256), (256, 128), (128, 128)]:
        g = torch.Generator().manual_seed(0)
        x = torch.randn(1, h, w, ci, generator=g).to(dev).bfloat16()
        off = (torch.randn(1, h, w, 18, generator=g) * 1.5).to(dev)
        wt = (torch.randn(co, ci, 3, 3, generator=g) / (ci * 9) ** 0.5).to(dev)
        pk = ops.PackedConv(wt, None)
        y = torch.empty(1, h, w, co, dtype=torch.bfloat16, device=dev)
        gemm = Conv(wt.permute(0, 2, 3, 1).reshape(co, 9 * ci, 1, 1).contiguous(), None)
        cols = torch.empty(1, h, w, 9 * ci, dtype=torch.bfloat16, device=dev)
        tf = timeit(lambda: ops.deform_conv_tc(x, off, pk, y))
        tu = timeit(lambda: (ops.deform_im2col(x, off, cols), gemm(cols, y)))
        tot_f += tf; tot_u += tu
        print("dcn %3d->%3d @%3dx%3d: fused %.4f ms   im2col+gemm %.4f ms" % (ci, co, h, w, tf, tu))
print("total fused %.3f ms, unfused %.3f ms" % (tot_f, tot_u))
