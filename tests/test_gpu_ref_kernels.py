"""GPU: the oracle's native-op restatements (oracle/ops.py) against the REFERENCE's OWN CUDA kernels, compiled for sm_100a
from the sources under /root/reference by oracle/ref_kernels/build.py (oracle/_ref/libvps_ref_kernels.so: the extracted
`__global__` bodies of resample2d / channelnorm / correlation / ROIAlign / nms / deformable_im2col with the launch geometry
of the reference's launchers).  This pins the oracle to the reference's actual kernels rather than to transcriptions."""
import ctypes
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libvps_ref_kernels.so")


@pytest.fixture(scope="module")
def ref(cuda):
    if not os.path.exists(LIB):
        from oracle.ref_kernels.build import build
        build()                                    # only possible where /root/reference is mounted
    return ctypes.CDLL(LIB)


def P(t):
    return ctypes.c_void_p(t.data_ptr())


def test_resample2d_and_channelnorm(ref):
    from oracle import ops as O
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 5, 19, 27, generator=g)
    flow = (torch.rand(2, 2, 19, 27, generator=g) - 0.5) * 14
    out = torch.empty(2, 5, 19, 27, device="cuda")
    assert ref.ref_resample2d(P(x.cuda()), P(flow.cuda()), P(out), 2, 5, 19, 27, 19, 27) == 0
    assert float((out.cpu() - O.resample2d(x, flow)).abs().max()) <= 2e-6 * float(x.abs().max())
    o2 = torch.empty(2, 1, 19, 27, device="cuda")
    assert ref.ref_channelnorm(P(x.cuda()), P(o2), 2, 5, 19, 27) == 0
    assert float((o2.cpu() - O.channelnorm(x)).abs().max()) <= 1e-6 * float(O.channelnorm(x).max())


@pytest.mark.parametrize("pad,md,s2,C", [(20, 20, 2, 64), (4, 4, 1, 96)])
def test_correlation(ref, pad, md, s2, C):
    """both call sites of the path: FlowNetC (pad 20, d 20, s2 2) and LiteFlowNetCorr (pad 4, d 4, s2 1)"""
    from oracle import ops as O
    g = torch.Generator().manual_seed(2)
    B, H, W = 1, 24, 32
    f1, f2 = torch.randn(B, C, H, W, generator=g), torch.randn(B, C, H, W, generator=g)
    want = O.correlation(f1, f2, pad, 1, md, 1, s2)
    D = 2 * (md // s2) + 1
    oh, ow = want.shape[2:]
    rb1 = torch.empty(B, H + 2 * pad, W + 2 * pad, C, device="cuda")
    rb2 = torch.empty_like(rb1)
    out = torch.empty(B, D * D, oh, ow, device="cuda")
    assert ref.ref_correlation(P(f1.cuda()), P(f2.cuda()), P(rb1), P(rb2), P(out), B, C, H, W, D * D, oh, ow, pad, 1, md, 1, s2) == 0
    assert float((out.cpu() - want).abs().max()) <= 1e-5 * max(1.0, float(want.abs().max()))


def test_roi_align(ref):
    from oracle import ops as O
    g = torch.Generator().manual_seed(3)
    feat = torch.randn(1, 16, 40, 56, generator=g)
    n = 37
    xy = torch.rand(n, 2, generator=g) * torch.tensor([200.0, 140.0])
    wh = torch.rand(n, 2, generator=g) * 90 + 1
    rois = torch.cat([torch.zeros(n, 1), xy, xy + wh], 1)
    rois[0, 1:] = torch.tensor([-20.0, -10.0, 5.0, 3.0])             # partly outside
    for S, scale in ((7, 0.25), (14, 0.25)):
        out = torch.empty(n, 16, S, S, device="cuda")
        assert ref.ref_roi_align(P(feat.cuda()), P(rois.cuda()), n, ctypes.c_float(scale), 2, 16, 40, 56, S, S, P(out)) == 0
        want = O.roi_align(feat, rois, S, scale, 2)
        assert float((out.cpu() - want).abs().max()) <= 1e-5 * max(1.0, float(want.abs().max()))


def test_nms(ref):
    """nms_kernel's bit mask + the host reduction of nms_cuda (nms_kernel.cu:99-121) == oracle.ops.nms, exactly"""
    from oracle import ops as O
    g = torch.Generator().manual_seed(4)
    for n in (5, 64, 65, 700):
        xy = torch.rand(n, 2, generator=g) * 300
        wh = torch.rand(n, 2, generator=g) * 80 + 2
        dets = torch.cat([xy, xy + wh, torch.rand(n, 1, generator=g)], 1)
        order = torch.sort(dets[:, 4], descending=True, stable=True)[1]
        bs = dets[order].contiguous()
        cb = (n + 63) // 64
        mask = torch.zeros(n * cb, dtype=torch.int64, device="cuda")
        assert ref.ref_nms_mask(P(bs.cuda()), n, ctypes.c_float(0.5), P(mask)) == 0
        m = mask.cpu().numpy().view(np.uint64).reshape(n, cb)
        remv = np.zeros(cb, np.uint64)
        keep = []
        for i in range(n):                                            # the reference's host loop
            if not (int(remv[i // 64]) >> (i % 64)) & 1:
                keep.append(i)
                remv |= m[i]
        got = torch.sort(order[torch.tensor(keep, dtype=torch.long)])[0]
        _, want = O.nms(dets, 0.5)
        assert torch.equal(got, want.sort()[0]), n


def test_deformable_im2col(ref):
    from oracle import ops as O
    g = torch.Generator().manual_seed(5)
    B, C, H, W = 2, 12, 13, 17
    x = torch.randn(B, C, H, W, generator=g)
    off = torch.randn(B, 18, H, W, generator=g) * 2.5
    off[:, :, 0] -= 4.0
    col = torch.empty(C * 9, B, H, W, device="cuda")
    assert ref.ref_deform_im2col(P(x.cuda()), P(off.cuda()), B, C, H, W, 3, 1, 1, 1, 1, P(col)) == 0
    want = O.deform_im2col(x, off)                                     # [B, C*9, H*W] (c-major, tap-minor)
    got = col.cpu().permute(1, 0, 2, 3).reshape(B, C * 9, H * W)
    assert float((got - want.reshape(B, C * 9, H * W)).abs().max()) <= 1e-5 * max(1.0, float(want.abs().max()))
