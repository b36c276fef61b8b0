"""CPU: pin the oracle's native-op restatements against literal transcriptions of the reference's CUDA
sources (tiny inputs) and against independent torch formulations where semantics coincide."""
import numpy as np
import torch
import torch.nn.functional as F

from oracle import ops as O


def test_correlation_matches_literal_loops():
    g = torch.Generator().manual_seed(0)
    for md, s2 in ((4, 1), (4, 2), (6, 2)):
        f1 = torch.randn(1, 5, 6, 7, generator=g)
        f2 = torch.randn(1, 5, 6, 7, generator=g)
        a = O.correlation(f1, f2, md, 1, md, 1, s2)
        b = O.correlation_loops(f1, f2, md, md, s2)
        assert a.shape == b.shape == (1, (2 * (md // s2) + 1) ** 2, 6, 7)
        assert (a - b).abs().max() < 1e-6


def test_correlation_zero_displacement_is_channel_mean_product():
    g = torch.Generator().manual_seed(1)
    f1 = torch.randn(2, 8, 5, 5, generator=g)
    f2 = torch.randn(2, 8, 5, 5, generator=g)
    out = O.correlation(f1, f2, 4, 1, 4, 1, 1)
    centre = out[:, 4 * 9 + 4]
    assert (centre - (f1 * f2).mean(1)).abs().max() < 1e-6


def test_resample2d_matches_literal_loops_and_grid_sample_interior():
    g = torch.Generator().manual_seed(2)
    src = torch.randn(1, 3, 9, 11, generator=g)
    flow = (torch.rand(1, 2, 9, 11, generator=g) - 0.5) * 8
    a = O.resample2d(src, flow)
    b = O.resample2d_loops(src, flow)
    assert (a - b).abs().max() < 1e-6
    # interior (no clamping): equals grid_sample(border, align_corners=True) with a pixel-unit grid
    H, W = 9, 11
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    xf, yf = xs + flow[0, 0], ys + flow[0, 1]
    grid = torch.stack([2 * xf / (W - 1) - 1, 2 * yf / (H - 1) - 1], -1)[None]
    gs = F.grid_sample(src, grid, mode="bilinear", padding_mode="border", align_corners=True)
    inside = (xf >= 0) & (xf <= W - 1) & (yf >= 0) & (yf <= H - 1)
    assert ((a - gs)[0, :, inside]).abs().max() < 1e-5


def test_channelnorm():
    x = torch.tensor([[[[3.0]], [[4.0]]]])
    assert float(O.channelnorm(x)) == 5.0


def test_flow_warp_mixed_convention():
    """zero flow is NOT the identity: sample position is x*W/(W-1) - 0.5."""
    W, H = 8, 4
    x = torch.arange(W, dtype=torch.float32).view(1, 1, 1, W).expand(1, 1, H, W).contiguous()
    out = O.flow_warp(x, torch.zeros(1, 2, H, W))
    pos = torch.arange(W, dtype=torch.float32) * W / (W - 1) - 0.5
    # a linear ramp sampled at `pos` (zero padding only bites at the two ends)
    assert (out[0, 0, 1, 1:-1] - pos[1:-1]).abs().max() < 1e-5


def test_deform_conv_zero_offset_equals_conv_and_integer_shift():
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 6, 7, 9, generator=g)
    w = torch.randn(4, 6, 3, 3, generator=g)
    off = torch.zeros(2, 18, 7, 9)
    assert (O.deform_conv(x, off, w) - F.conv2d(x, w, padding=1)).abs().max() < 1e-5
    # all taps shifted by (+1 row, -2 cols) == conv over the shifted (zero padded) image
    off[:, 0::2] = 1.0
    off[:, 1::2] = -2.0
    xs = torch.zeros_like(x)
    xs[:, :, :-1, 2:] = x[:, :, 1:, :-2]
    ref = F.conv2d(xs, w, padding=1)
    got = O.deform_conv(x, off, w)
    # borders differ (the shifted image loses the conv padding ring); compare the interior
    assert (got - ref)[:, :, 2:-2, 3:-3].abs().max() < 1e-5


def _roi_align_literal(feat, rois, ps, scale, sn):
    """literal transcription of ROIAlignForward + bilinear_interpolate (roi_align_kernel.cu:16-128)"""
    f = feat.numpy()
    B, C, H, W = f.shape
    out = np.zeros((rois.shape[0], C, ps, ps), np.float32)

    def bil(d, y, x):
        if y < -1.0 or y > H or x < -1.0 or x > W:
            return 0.0
        y = max(y, 0.0); x = max(x, 0.0)
        yl, xl = int(y), int(x)
        if yl >= H - 1:
            yh = yl = H - 1; y = float(yl)
        else:
            yh = yl + 1
        if xl >= W - 1:
            xh = xl = W - 1; x = float(xl)
        else:
            xh = xl + 1
        ly, lx = y - yl, x - xl
        hy, hx = 1 - ly, 1 - lx
        return hy * hx * d[yl, xl] + hy * lx * d[yl, xh] + ly * hx * d[yh, xl] + ly * lx * d[yh, xh]
    for n, r in enumerate(rois.numpy()):
        b = int(r[0])
        sw, sh, ew, eh = r[1] * scale, r[2] * scale, (r[3] + 1) * scale, (r[4] + 1) * scale
        rw, rh = max(ew - sw, 0.0), max(eh - sh, 0.0)
        bh, bw = rh / ps, rw / ps
        for c in range(C):
            for ph in range(ps):
                for pw in range(ps):
                    acc = 0.0
                    for iy in range(sn):
                        y = sh + ph * bh + (iy + 0.5) * bh / sn
                        for ix in range(sn):
                            x = sw + pw * bw + (ix + 0.5) * bw / sn
                            acc += bil(f[b, c], y, x)
                    out[n, c, ph, pw] = acc / (sn * sn)
    return torch.from_numpy(out)


def test_roi_align_matches_literal_transcription():
    g = torch.Generator().manual_seed(4)
    feat = torch.randn(2, 3, 12, 16, generator=g)
    rois = torch.tensor([[0, 3.2, 4.1, 30.7, 22.9], [1, -5.0, -3.0, 80.0, 60.0], [0, 10, 10, 10, 10],
                         [1, 40.0, 2.0, 63.5, 47.0]])
    for ps in (3, 7):
        a = O.roi_align(feat, rois, ps, 0.25, 2)
        b = _roi_align_literal(feat, rois, ps, 0.25, 2)
        assert (a - b).abs().max() < 1e-5


def test_nms_matches_bruteforce_plus_one_iou():
    g = torch.Generator().manual_seed(5)
    n = 200
    xy = torch.rand(n, 2, generator=g) * 100
    wh = torch.rand(n, 2, generator=g) * 40 + 1
    dets = torch.cat([xy, xy + wh, torch.rand(n, 1, generator=g)], 1)
    kept, inds = O.nms(dets, 0.5)
    d = dets.numpy()
    order = np.argsort(-d[:, 4], kind="stable")
    keep = []
    for i in order:
        ok = True
        for j in keep:
            xx1, yy1 = max(d[i, 0], d[j, 0]), max(d[i, 1], d[j, 1])
            xx2, yy2 = min(d[i, 2], d[j, 2]), min(d[i, 3], d[j, 3])
            w, h = max(np.float32(0), xx2 - xx1 + 1), max(np.float32(0), yy2 - yy1 + 1)
            inter = w * h
            a1 = (d[i, 2] - d[i, 0] + 1) * (d[i, 3] - d[i, 1] + 1)
            a2 = (d[j, 2] - d[j, 0] + 1) * (d[j, 3] - d[j, 1] + 1)
            if inter / (a1 + a2 - inter) > 0.5:
                ok = False
                break
        if ok:
            keep.append(i)
    assert sorted(keep) == inds.tolist()
    assert torch.equal(kept, dets[inds])
    assert O.gpu_nms_upsnet(d, 0.5) == keep       # UPSNet variant returns score order


def test_resnet_trunk_equals_torchvision():
    import torchvision
    from oracle.model import ResNet50
    torch.manual_seed(0)
    tv = torchvision.models.resnet50(weights=None).eval()
    for m in tv.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.1); m.running_var.uniform_(0.5, 1.5)
    mine = ResNet50().eval()
    sd = {k: v for k, v in tv.state_dict().items() if not k.startswith("fc.")}
    mine.load_state_dict(sd, strict=True)
    x = torch.randn(1, 3, 64, 96)
    with torch.no_grad():
        outs = mine(x)
        t = tv.maxpool(tv.relu(tv.bn1(tv.conv1(x))))
        refs = []
        for l in (tv.layer1, tv.layer2, tv.layer3, tv.layer4):
            t = l(t)
            refs.append(t)
    for a, b in zip(outs, refs):
        assert (a - b).abs().max() < 1e-4
