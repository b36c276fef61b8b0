"""Host logic of the thin-convolution path (vps_b200/layers.py:Conv, vps_tap_gather3x3): a 3x3 / stride 1 / pad 1 convolution with
<= 3 output channels == a 1x1 convolution over tap-major output channels followed by a 9-tap gather.  Checked on the CPU with the
weights exactly as layers.Conv packs them (the GPU test is tests/test_gpu_conv_tc32.py::test_thin_3x3_as_tap_major_1x1)."""
import torch
import torch.nn.functional as F


def _gather(z, cout):
    """out[n, co, y, x] = sum_t z[n, t*cout+co, y + t//3 - 1, x + t%3 - 1], zero outside (include/vps_b200.h: vps_tap_gather3x3)"""
    n, _, h, w = z.shape
    zp = F.pad(z, (1, 1, 1, 1))
    out = torch.zeros(n, cout, h, w, dtype=z.dtype)
    for t in range(9):
        r, s = t // 3, t % 3
        out += zp[:, t * cout:(t + 1) * cout, r:r + h, s:s + w]
    return out


def test_tap_major_weights_reproduce_the_3x3_convolution():
    from vps_b200.layers import Conv
    g = torch.Generator().manual_seed(3)
    for cin, cout in ((194, 2), (16, 2), (64, 3), (7, 1)):
        x = torch.randn(2, cin, 9, 13, generator=g, dtype=torch.float64)
        w = torch.randn(cout, cin, 3, 3, generator=g, dtype=torch.float64)
        b = torch.randn(cout, generator=g, dtype=torch.float64)
        scale = torch.rand(cout, generator=g, dtype=torch.float64) + 0.5
        layer = Conv(w.float(), b.float(), stride=1, pad=1, scale=scale.float())
        assert layer.pk_tap is not None and layer.pk_tap.kh == 1 and layer.pk_tap.cout == 9 * cout
        wt = layer.pk_tap.weight.double()                       # [9*cout, cin, 1, 1], scale folded in
        z = F.conv2d(x, wt)
        got = _gather(z, cout) + b.view(1, -1, 1, 1)
        ref = F.conv2d(x, (w.float() * scale.float().view(-1, 1, 1, 1)).double(), b, padding=1)
        assert (got - ref).abs().max().item() <= 1e-9 * max(1.0, ref.abs().max().item())


def test_only_thin_3x3_layers_take_the_path():
    from vps_b200.layers import Conv
    w = torch.randn(4, 8, 3, 3)
    assert Conv(w, None, stride=1, pad=1).pk_tap is None                       # 4 output channels: 36 > 32 tap-major columns
    assert Conv(w[:2], None, stride=2, pad=1).pk_tap is None                   # strided
    assert Conv(w[:2, :, :1, :1].contiguous(), None).pk_tap is None            # 1x1
    assert Conv(w[:2], None, stride=1, pad=1).pk_tap is not None
