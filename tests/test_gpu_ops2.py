"""GPU parity, per kernel: resampling / neck / detection / fusion kernels vs the oracle restatements on
identical seeded inputs.  Index-valued outputs (NMS keep lists, sort order, track ids, keep flags, label maps)
must be bit-exact; float outputs within 1e-5 (relative to the output scale)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def to_nhwc(t, dtype=torch.float32, dev="cuda"):
    n, c, h, w = t.shape
    cs = (c + 7) // 8 * 8
    buf = torch.zeros(n, h, w, cs, dtype=dtype)
    buf[..., :c] = t.permute(0, 2, 3, 1).to(dtype)
    return buf.to(dev)[..., :c]


def to_nchw(t):
    return t.float().permute(0, 3, 1, 2).contiguous().cpu()


def new_nhwc(n, h, w, c, dtype=torch.float32):
    cs = (c + 7) // 8 * 8
    return torch.zeros(n, h, w, cs, dtype=dtype, device="cuda")[..., :c]


def close(a, b, tol=1e-5):
    return float((a - b).abs().max()) <= tol * max(1.0, float(b.abs().max()))


def test_resize_pool_layout(cuda):
    from vps_b200 import ops
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 19, 13, 17, generator=g)
    xd = to_nhwc(x)
    for (oh, ow) in ((52, 68), (26, 34), (3, 4)):
        out = new_nhwc(2, oh, ow, 19)
        ops.resize_bilinear(xd, out, mul=0.25)
        assert close(to_nchw(out), F.interpolate(x, size=(oh, ow), mode="bilinear", align_corners=False) * 0.25)
        ops.resize_nearest(xd, out)
        assert torch.equal(to_nchw(out), F.interpolate(x, size=(oh, ow), mode="nearest"))
    y = torch.randn(1, 8, 26, 34, generator=g)
    acc = to_nhwc(torch.randn(1, 8, 52, 68, generator=g))
    ref = to_nchw(acc) + F.interpolate(y, scale_factor=2, mode="nearest")
    ops.resize_nearest(to_nhwc(y), acc, accumulate=True)
    assert close(to_nchw(acc), ref)
    for (k, s, p, avg) in ((3, 2, 1, False), (3, 2, 1, True), (1, 2, 0, False)):
        x2 = torch.randn(1, 8, 15, 22, generator=g)
        ref = F.avg_pool2d(x2, k, s, p) if avg else F.max_pool2d(x2, k, s, p)
        out = new_nhwc(1, ref.shape[2], ref.shape[3], 8)
        ops.pool2d(to_nhwc(x2), out, k, s, p, avg)
        assert close(to_nchw(out), ref)
    src = torch.randn(1, 3, 9, 11, generator=g).cuda()
    d = new_nhwc(1, 9, 11, 3)
    ops.nchw_to_nhwc(src, d)
    back = torch.empty(1, 3, 9, 11, device="cuda")
    ops.nhwc_to_nchw(d, back)
    assert torch.equal(back, src)


def test_groupnorm(cuda):
    from vps_b200 import ops
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, 128, 20, 33, generator=g) * 2 + 0.5
    gamma, beta = torch.rand(128, generator=g) + 0.5, torch.randn(128, generator=g)
    ref = F.relu(F.group_norm(x, 32, gamma, beta, 1e-5))
    out = new_nhwc(1, 20, 33, 128)
    ops.groupnorm(to_nhwc(x), out, gamma.cuda(), beta.cuda(), 32, 1e-5, relu=True)
    assert close(to_nchw(out), ref, 2e-5)


def test_flownet_input(cuda):
    from oracle.model import IMG_MEAN, IMG_STD
    from vps_b200 import ops
    g = torch.Generator().manual_seed(2)
    img, ref = torch.randn(1, 3, 32, 48, generator=g), torch.randn(1, 3, 32, 48, generator=g)
    std, mean = torch.tensor(IMG_STD).view(1, 3, 1, 1), torch.tensor(IMG_MEAN).view(1, 3, 1, 1)
    rgbs = torch.stack([img * std + mean, ref * std + mean], dim=2)
    m = rgbs.contiguous().view(1, 3, -1).mean(-1).view(1, 3, 1, 1, 1)
    x = (rgbs - m) / 255.0
    x = torch.cat((x[:, :, 0], x[:, :, 1]), 1)
    out = new_nhwc(1, 32, 48, 6)
    ops.flownet_input(img.cuda(), ref.cuda(), IMG_STD, IMG_MEAN, 255.0, torch.empty(3, dtype=torch.float64, device="cuda"), out)
    assert close(to_nchw(out), x, 1e-5)


def test_bfp_gather_scatter_warp(cuda):
    from oracle import ops as O
    from vps_b200 import ops
    g = torch.Generator().manual_seed(3)
    sizes = [(24, 40), (12, 20), (6, 10), (3, 5), (2, 3)]
    lv = [torch.randn(1, 16, h, w, generator=g) for h, w in sizes]
    ref = sum(F.interpolate(t, size=sizes[0], mode="nearest") for t in lv) / 5
    out = new_nhwc(1, 24, 40, 16)
    ops.bfp_gather([to_nhwc(t) for t in lv], out)
    assert close(to_nchw(out), ref)
    bsf = torch.randn(1, 16, 24, 40, generator=g)
    for t, (h, w) in zip(lv, sizes):
        o = new_nhwc(1, h, w, 16)
        ops.bfp_scatter(to_nhwc(bsf), to_nhwc(t), o)
        assert close(to_nchw(o), F.adaptive_max_pool2d(bsf, (h, w)) + t)
    flow = (torch.rand(1, 2, 24, 40, generator=g) - 0.5) * 12
    o = new_nhwc(1, 24, 40, 16)
    ops.flow_warp(to_nhwc(bsf), to_nhwc(flow), o)
    assert close(to_nchw(o), O.flow_warp(bsf, flow), 2e-5)


def test_tcea_kernels(cuda):
    from vps_b200 import ops
    g = torch.Generator().manual_seed(4)
    C, H, W = 64, 10, 14
    f0, f1, e0, e1, er = [torch.randn(1, C, H, W, generator=g) * 0.3 for _ in range(5)]
    p0 = torch.sigmoid((e0 * er).sum(1, keepdim=True))
    p1 = torch.sigmoid((e1 * er).sum(1, keepdim=True))
    ref = torch.cat([f0 * p0, f1 * p1], 1)
    out = new_nhwc(1, H, W, 2 * C)
    ops.tcea_temporal(*[to_nhwc(t) for t in (f0, f1, e0, e1, er)], out)
    assert close(to_nchw(out), ref, 2e-5)
    o2 = new_nhwc(1, H, W, C)
    ops.tcea_combine(to_nhwc(f0), to_nhwc(e0), to_nhwc(e1), o2)
    assert close(to_nchw(o2), f0 * torch.sigmoid(e0) * 2 + e1, 2e-5)


def test_deform_im2col_and_gemm(cuda):
    from oracle import ops as O
    from vps_b200 import ops
    from vps_b200.layers import Conv
    g = torch.Generator().manual_seed(5)
    C, Co, H, W = 32, 24, 11, 15
    x = torch.randn(1, C, H, W, generator=g)
    off = torch.randn(1, 18, H, W, generator=g) * 2.5
    w = torch.randn(Co, C, 3, 3, generator=g) / (C * 9) ** 0.5
    ref = O.deform_conv(x, off, w)
    cols = new_nhwc(1, H, W, 9 * C)
    ops.deform_im2col(to_nhwc(x), to_nhwc(off), cols)
    ref_cols = O.deform_im2col(x, off).view(1, C, 9, H, W).permute(0, 2, 1, 3, 4).reshape(1, 9 * C, H, W)   # tap-major
    assert close(to_nchw(cols), ref_cols)
    w1 = w.permute(0, 2, 3, 1).reshape(Co, 9 * C, 1, 1).contiguous().cuda()
    y = Conv(w1, None)(cols)
    assert close(to_nchw(y), ref, 2e-5)


def test_deform_conv_tc_fused(cuda):
    """vps_deform_conv_tc (sampling fused into the tensor-core operand ring) vs vps_deform_im2col + 1x1 GEMM (same bf16
    columns, different fp32 accumulation order over K: chunk-major vs tap-major -> equal up to one bf16 rounding), and
    vs the oracle DCNv1 (deform_conv_cuda forward) on bf16-representable operands; ragged tiles, offsets that leave the
    image, two images."""
    from oracle import ops as O
    from vps_b200 import ops
    from vps_b200.layers import Conv
    g = torch.Generator().manual_seed(15)
    for (N, C, Co, H, W) in [(1, 128, 128, 19, 37), (2, 256, 256, 12, 20), (1, 64, 32, 8, 16)]:
        x = torch.randn(N, C, H, W, generator=g).bfloat16().float()
        off = torch.randn(N, 18, H, W, generator=g) * 3.0
        off[:, :, 0, :] -= 4.0                       # some samples fall outside the image
        w = (torch.randn(Co, C, 3, 3, generator=g) / (C * 9) ** 0.5).bfloat16().float()
        ref = O.deform_conv(x, off, w)
        xd = to_nhwc(x).bfloat16()
        od = to_nhwc(off)
        pk = ops.PackedConv(w.cuda(), None)
        y = torch.full((N, H, W, Co), float("nan"), dtype=torch.bfloat16, device="cuda")
        ops.deform_conv_tc(xd, od, pk, y)
        cols = torch.empty(N, H, W, 9 * C, dtype=torch.bfloat16, device="cuda")
        ops.deform_im2col(xd, od, cols)
        w1 = w.permute(0, 2, 3, 1).reshape(Co, 9 * C, 1, 1).contiguous().cuda()
        y2 = Conv(w1, None)(cols)
        torch.cuda.synchronize()
        assert not torch.isnan(y.float()).any()
        scale = max(1.0, float(y2.float().abs().max()))
        assert float((y.float() - y2.float()).abs().max()) <= 2.0 ** -7 * scale, "fused vs im2col + GEMM"
        assert float((y.float() != y2.float()).float().mean()) < 0.05          # almost all outputs are bit-identical
        got = y.float().permute(0, 3, 1, 2).cpu()
        assert float((got - ref).abs().max()) <= 2e-2 * max(1.0, float(ref.abs().max()))


def test_deform_conv_tc32_fused(cuda):
    """vps_deform_conv_tc32 (fp32 activations, sampling warps feed the split fp16 operand planes) vs the oracle DCNv1
    (deform_conv_cuda forward) in fp32: ragged tiles, offsets that leave the image, two images, cout > 128 (two N tiles)."""
    from oracle import ops as O
    from vps_b200 import ops
    g = torch.Generator().manual_seed(16)
    old = ops.F32_TC[0]
    ops.F32_TC[0] = True
    try:
        for (N, C, Co, H, W) in [(1, 128, 128, 19, 37), (2, 256, 256, 12, 20), (1, 64, 32, 8, 16), (1, 256, 128, 64, 128)]:
            x = torch.randn(N, C, H, W, generator=g)
            off = torch.randn(N, 18, H, W, generator=g) * 3.0
            off[:, :, 0, :] -= 4.0                       # some samples fall outside the image
            w = torch.randn(Co, C, 3, 3, generator=g) / (C * 9) ** 0.5
            ref = O.deform_conv(x, off, w)
            pk = ops.PackedConv(w.cuda(), None)
            y = torch.full((N, H, W, Co), float("nan"), dtype=torch.float32, device="cuda")
            ops.deform_conv_tc32(to_nhwc(x), to_nhwc(off), pk, y)
            torch.cuda.synchronize()
            got = y.permute(0, 3, 1, 2).cpu()
            assert not torch.isnan(got).any()
            assert float((got - ref).abs().max()) <= 1e-5 * max(1.0, float(ref.abs().max())), (N, C, Co, H, W)
        assert ops.tc32_overflow() == 0
    finally:
        ops.F32_TC[0] = old


def test_roi_align_multilevel(cuda):
    from oracle.model import roi_extract
    from vps_b200 import ops
    g = torch.Generator().manual_seed(6)
    feats = [torch.randn(1, 16, 64 // s * 2, 128 // s * 2, generator=g) for s in (4, 8, 16, 32)]   # 128x256 image
    n = 300
    xy = torch.rand(n, 2, generator=g) * torch.tensor([256.0, 128.0])
    wh = torch.exp(torch.rand(n, 2, generator=g) * 5.5)
    boxes = torch.cat([xy - wh / 2, xy + wh / 2], 1)
    boxes[:5] = torch.tensor([[0, 0, 0, 0], [-20, -20, 300, 200], [10, 10, 9, 9], [255, 127, 255, 127], [0, 0, 255, 127]]).float()
    rois = torch.cat([torch.zeros(n, 1), boxes], 1)
    fd = [to_nhwc(f) for f in feats]
    for ps in (7, 14):
        ref = roi_extract(feats, rois, ps)
        out = torch.empty(n, ps, ps, 16, device="cuda")
        ops.roi_align(fd, [4, 8, 16, 32], rois.cuda(), n, out, 2)
        got = out.permute(0, 3, 1, 2).cpu()
        assert close(got, ref, 2e-5)
    cnt = torch.tensor([100], dtype=torch.int32, device="cuda")
    out = torch.full((n, 7, 7, 16), 5.0, device="cuda")
    ops.roi_align(fd, [4, 8, 16, 32], rois.cuda(), n, out, 2, nroi_dev=cnt)
    assert float(out[100:].abs().max()) == 0.0


@pytest.mark.parametrize("n,thr", [(1, 0.5), (63, 0.7), (64, 0.5), (1000, 0.7), (1777, 0.5), (8000, 0.5)])
def test_sort_and_nms_bit_exact(cuda, n, thr):
    from oracle import ops as O
    from vps_b200 import ops
    g = torch.Generator().manual_seed(n)
    xy = torch.rand(n, 2, generator=g) * 200
    wh = torch.rand(n, 2, generator=g) * 60 + 1
    sc = torch.rand(n, generator=g)
    sc[::7] = sc[0]                                   # ties: stable order must win
    dets = torch.cat([xy, xy + wh, sc[:, None]], 1)
    s_sorted = torch.empty(n, device="cuda")
    i_sorted = torch.empty(n, dtype=torch.int32, device="cuda")
    ops.sort_desc(sc.cuda(), s_sorted, i_sorted, n, torch.empty(ops.sort_ws_bytes(n), dtype=torch.uint8, device="cuda"))
    order = np.argsort(-sc.numpy(), kind="stable")
    assert np.array_equal(i_sorted.cpu().numpy(), order)
    d_sorted = dets[torch.from_numpy(order)].contiguous().cuda()
    keep = torch.empty(n, dtype=torch.int32, device="cuda")
    nk = torch.zeros(1, dtype=torch.int32, device="cuda")
    ops.nms(d_sorted, n, thr, keep, nk, torch.empty(max(8, ops.nms_ws_bytes(n)), dtype=torch.uint8, device="cuda"))
    k = int(nk.item())
    ref_keep = O.nms_sorted_numpy(dets.numpy()[order], thr)
    assert k == len(ref_keep) and np.array_equal(keep[:k].cpu().numpy(), ref_keep)
    # reference-facing form: indices into the original array, ascending (nms_wrapper.py:49)
    _, inds = O.nms(dets, thr)
    assert np.array_equal(np.sort(order[keep[:k].cpu().numpy()]), inds.numpy())


def test_nms_empty_and_device_count(cuda):
    from vps_b200 import ops
    nk = torch.full((1,), 7, dtype=torch.int32, device="cuda")
    ops.nms(torch.empty(0, 5, device="cuda"), 0, 0.5, torch.empty(1, dtype=torch.int32, device="cuda"), nk,
            torch.empty(8, dtype=torch.uint8, device="cuda"))
    assert int(nk.item()) == 0
    d = torch.tensor([[0, 0, 10, 10, .9], [1, 1, 11, 11, .8], [50, 50, 60, 60, .7], [0, 0, 10, 10, .6]], device="cuda")
    keep = torch.empty(4, dtype=torch.int32, device="cuda")
    cnt = torch.tensor([3], dtype=torch.int32, device="cuda")
    ops.nms(d, 4, 0.5, keep, nk, torch.empty(64, dtype=torch.uint8, device="cuda"), n_dev=cnt)
    assert keep[:int(nk.item())].tolist() == [0, 2]


def test_nms_batch_vs_oracle(cuda):
    """vps_nms_batch: several independent problems (RPN levels) of different sizes in one launch pair, each equal to
    the oracle's greedy NMS (nms_kernel.cu:13-131) on the same boxes."""
    from oracle import ops as O
    from vps_b200 import ops
    g = torch.Generator().manual_seed(21)
    seg, ns = 1000, [1000, 777, 64, 1, 0]
    dets = torch.zeros(len(ns) * seg, 5)
    refs = []
    for b, n in enumerate(ns):
        if n == 0:
            refs.append([])
            continue
        xy = torch.rand(n, 2, generator=g) * 300
        wh = torch.rand(n, 2, generator=g) * 80 + 4
        sc = torch.sort(torch.rand(n, generator=g), descending=True).values
        d = torch.cat([xy, xy + wh, sc[:, None]], 1)
        dets[b * seg:b * seg + n] = d
        _, keep = O.nms(d, 0.7)
        refs.append(keep.tolist())
    dd = dets.cuda()
    keep = torch.full((len(ns) * seg,), -1, dtype=torch.int32, device="cuda")
    nk = torch.full((len(ns),), -5, dtype=torch.int32, device="cuda")
    ws = torch.empty(len(ns) * ops.nms_ws_bytes(seg), dtype=torch.uint8, device="cuda")
    ops.nms_batch(dd, ns, seg, 0.7, keep, nk, ws)
    torch.cuda.synchronize()
    for b, n in enumerate(ns):
        k = int(nk[b].item())
        assert keep[b * seg:b * seg + k].tolist() == refs[b], "problem %d" % b


def test_rpn_level_pipeline(cuda):
    """sigmoid -> stable top-k -> decode -> NMS of one RPN level vs RPNHead.get_bboxes' per-level body."""
    from oracle import ops as O
    from oracle.model import delta2bbox, gen_base_anchors, grid_anchors, stable_topk
    from vps_b200 import ops
    g = torch.Generator().manual_seed(9)
    h, w, A, stride = 24, 40, 3, 8
    cls = torch.randn(1, A, h, w, generator=g) * 2
    reg = torch.randn(1, 4 * A, h, w, generator=g) * 0.4
    img_shape = (h * stride, w * stride, 3)
    anchors = grid_anchors(gen_base_anchors(stride), (h, w), stride)
    scores = cls[0].permute(1, 2, 0).reshape(-1).sigmoid()
    bp = reg[0].permute(1, 2, 0).reshape(-1, 4)
    sc, idx = stable_topk(scores, 1000)
    props = delta2bbox(anchors[idx], bp[idx], (0, 0, 0, 0), (1, 1, 1, 1), img_shape)
    ref, _ = O.nms(torch.cat([props, sc[:, None]], 1), 0.7)
    head = to_nhwc(torch.cat([cls, reg], 1))
    n = h * w * A
    s = torch.empty(n, device="cuda")
    ops.sigmoid_flat(head[..., :A], s)
    assert close(s.cpu(), scores, 1e-6)
    s.copy_(scores.cuda())                       # identical keys from here on -> indices must be bit-exact
    s_sorted, i_sorted = torch.empty(n, device="cuda"), torch.empty(n, dtype=torch.int32, device="cuda")
    ops.sort_desc(s, s_sorted, i_sorted, n, torch.empty(ops.sort_ws_bytes(n), dtype=torch.uint8, device="cuda"))
    assert torch.equal(i_sorted[:1000].cpu().long(), idx)
    dets = torch.empty(1000, 5, device="cuda")
    ops.rpn_decode(s_sorted, i_sorted, 1000, head[..., A:5 * A], stride, gen_base_anchors(stride).cuda(),
                   float(img_shape[0]), float(img_shape[1]), dets)
    assert float((dets[:, :4].cpu() - props).abs().max()) <= 1e-3
    keep = torch.empty(1000, dtype=torch.int32, device="cuda")
    nk = torch.zeros(1, dtype=torch.int32, device="cuda")
    dets[:, :4] = props.cuda()
    ops.nms(dets, 1000, 0.7, keep, nk, torch.empty(ops.nms_ws_bytes(1000), dtype=torch.uint8, device="cuda"))
    got = dets[keep[:int(nk.item())].long()].cpu()
    assert got.shape == ref.shape and torch.equal(got, ref)       # reference returns rows in original (= score) order


def test_maskroi_vs_oracle(cuda):
    from oracle.model import mask_roi
    from vps_b200 import ops
    from vps_b200.detector import MAX_DET_CAP
    g = torch.Generator().manual_seed(10)
    n, nc, H, W = 400, 9, 256.0, 512.0
    xy = torch.rand(n, 2, generator=g) * torch.tensor([W, H])
    wh = torch.rand(n, 2, generator=g) * 120 + 4
    rois = torch.cat([torch.zeros(n, 1), (xy - wh / 2).clamp(min=0), xy + wh / 2], 1)
    rois[:, 3].clamp_(max=W - 1); rois[:, 4].clamp_(max=H - 1)
    cls_score = torch.randn(n, nc, generator=g) * 3
    bbox_pred = torch.randn(n, 4 * nc, generator=g) * 0.7
    prob = F.softmax(cls_score, 1)
    o_prob, o_rois, o_cls = mask_roi(rois, bbox_pred, prob, np.array([[H, W, 1.0]]))
    m = n * (nc - 1)
    dev = "cuda"
    y = torch.cat([cls_score, bbox_pred], 1).cuda()
    cand, ccls, cprob = torch.empty(m, 5, device=dev), torch.empty(m, dtype=torch.int32, device=dev), torch.empty(m, device=dev)
    ncand = torch.zeros(1, dtype=torch.int32, device=dev)
    ops.maskroi_candidates(rois.cuda(), y[:, :nc], y[:, nc:], n, nc, 0.6, H, W, cand, ccls, cprob, ncand)
    psort, slot = torch.empty(m, device=dev), torch.empty(m, dtype=torch.int32, device=dev)
    ops.sort_desc(cprob, psort, slot, m, torch.empty(ops.sort_ws_bytes(m), dtype=torch.uint8, device=dev))
    csort = torch.empty(m, 5, device=dev)
    ops.gather_rows(cand, slot, m, 5, csort)
    keep, nk = torch.empty(m, dtype=torch.int32, device=dev), torch.zeros(1, dtype=torch.int32, device=dev)
    ops.nms(csort, m, 0.5, keep, nk, torch.empty(ops.nms_ws_bytes(m), dtype=torch.uint8, device=dev), n_dev=ncand)
    det_rois = torch.empty(MAX_DET_CAP, 5, device=dev)
    cidx, cp = torch.empty(MAX_DET_CAP, dtype=torch.int32, device=dev), torch.empty(MAX_DET_CAP, device=dev)
    kout = torch.zeros(2, dtype=torch.int32, device=dev)
    ops.maskroi_finalize(csort, slot, ccls, keep, nk, 100, MAX_DET_CAP, det_rois, cidx, cp, kout)
    k, dummy = kout.tolist()
    assert dummy == 0 and k == o_rois.shape[0]
    assert torch.equal(cidx[:k].cpu().long(), o_cls)
    assert float((det_rois[:k].cpu() - o_rois).abs().max()) <= 1e-3
    assert float((cp[:k].cpu() - o_prob).abs().max()) <= 1e-5


def test_track_assign_vs_oracle(cuda):
    from oracle.model import PanopticFuseTrack as Oracle
    from vps_b200 import ops
    g = torch.Generator().manual_seed(11)
    k, m, dim, cap = 37, 52, 64, 256
    emb, ref_emb = torch.randn(k, dim, generator=g) * 0.4, torch.randn(m, dim, generator=g) * 0.4
    ref_emb[:20] = emb[:20] + 0.05 * torch.randn(20, dim, generator=g)      # real matches
    ref_emb[20:24] = emb[3:7]                                                # duplicates -> "undo" branch
    mk = lambda n: torch.cat([torch.rand(n, 2, generator=g) * 100, torch.rand(n, 2, generator=g) * 100 + 100], 1)
    db, rb = mk(k), mk(m)
    rb[:20] = db[:20] + 1.0
    dl, rl = torch.randint(0, 8, (k,), generator=g), torch.randint(0, 8, (m,), generator=g)
    rl[:20] = dl[:20]
    prob = torch.rand(k, generator=g) * 0.4 + 0.6

    class TH(object):
        match_coeff = (1.0, 2.0, 10.0)

        def __call__(self, x, r):
            return torch.cat([torch.zeros(x.size(0), 1), x @ r.t()], 1)
        compute_comp_scores = Oracle.__dict__  # placeholder, replaced below
    o = Oracle.__new__(Oracle)
    torch.nn.Module.__init__(o)
    from oracle.model import TrackHead
    th = TrackHead.__new__(TrackHead)
    torch.nn.Module.__init__(th)
    th.match_coeff = (1.0, 2.0, 10.0)
    th.forward = lambda x, r: torch.cat([torch.zeros(x.size(0), 1), x @ r.t()], 1)
    o.track_head = th
    o.prev_bboxes, o.prev_roi_feats, o.prev_det_labels = rb.clone(), ref_emb.clone(), rl.clone()
    taps = {}
    ids_ref = o.track(db, dl, emb, prob, False, taps)
    dev = "cuda"
    ids = torch.empty(k, dtype=torch.int32, device=dev)
    mids = torch.empty(k, dtype=torch.int32, device=dev)
    comp = torch.empty(k, m + 1, device=dev)
    mem_src = torch.empty(cap, dtype=torch.int32, device=dev)
    new_m = torch.zeros(1, dtype=torch.int32, device=dev)
    ops.track_assign(emb.cuda(), ref_emb.cuda(), k, m, dim, db.cuda(), rb.cuda(), dl.int().cuda(), rl.int().cuda(), prob.cuda(),
                     (1.0, 2.0, 10.0), cap, ids, mids, comp, mem_src, new_m,
                     torch.empty((k * m + k + 2 * cap) * 4, dtype=torch.uint8, device=dev))
    assert float((comp.cpu() - taps["comp_scores"]).abs().max()) <= 1e-4
    assert np.array_equal(ids.cpu().numpy(), np.asarray(ids_ref))            # track ids bit-exact
    assert int(new_m.item()) == o.prev_roi_feats.shape[0]
    # memory contents after the update equal the oracle's (features = embeddings here)
    feats = torch.zeros(cap, dim, device=dev); feats[:m] = ref_emb.cuda()
    boxes = torch.zeros(cap, 4, device=dev); boxes[:m] = rb.cuda()
    labels = torch.zeros(cap, dtype=torch.int32, device=dev); labels[:m] = rl.int().cuda()
    ops.track_update(feats, emb.cuda(), dim, boxes, db.cuda(), labels, dl.int().cuda(), mem_src, m, cap, new_m)
    nm = int(new_m.item())
    assert torch.equal(feats[:nm].cpu(), o.prev_roi_feats) and torch.equal(boxes[:nm].cpu(), o.prev_bboxes)
    assert torch.equal(labels[:nm].cpu().long(), o.prev_det_labels)


def test_mask_removal_and_fusion_bit_exact(cuda):
    """MaskRemoval keep decisions and the fused panoptic / semantic argmax vs the oracle's cv2 + torch pipeline."""
    from oracle.model import mask_removal, seg_term
    from vps_b200 import ops
    g = torch.Generator().manual_seed(12)
    H, W, k, ms, NS = 96, 160, 40, 28, 11
    fcn_score = torch.randn(1, 19, H // 4, W // 4, generator=g) * 2
    fcn_output = F.interpolate(fcn_score, scale_factor=4, mode="bilinear", align_corners=False)
    xy = torch.rand(k, 2, generator=g) * torch.tensor([W * 1.0, H * 1.0])
    wh = torch.exp(torch.rand(k, 2, generator=g) * 3.5) + 1
    boxes = torch.cat([(xy - wh / 2), (xy + wh / 2)], 1)
    boxes[:, 0::2].clamp_(0, W - 1); boxes[:, 1::2].clamp_(0, H - 1)
    boxes[0] = torch.tensor([0.0, 0.0, W - 1.0, H - 1.0])
    boxes[1] = torch.tensor([5.3, 7.9, 5.9, 8.2])
    cls_idx = torch.randint(1, 9, (k,), generator=g)
    cls_idx[:8] = 3                                   # force same-class overlaps
    prob = torch.rand(k, generator=g) * 0.39 + 0.6
    mlog = torch.randn(k, 1, ms, ms, generator=g) * 2 + 0.3
    rois = torch.cat([torch.zeros(k, 1), boxes], 1)
    keep_ref, energy = mask_removal(boxes, prob, mlog, cls_idx, (H, W))
    stuff, inst = seg_term(cls_idx[keep_ref], fcn_output, rois[keep_ref] * 4.0)
    logits = torch.cat([stuff, inst + energy], 1)
    pano_ref = torch.max(F.softmax(logits, 1), 1)[1][0]
    sem_ref = torch.max(F.softmax(fcn_output, 1), 1)[1][0]
    dev = "cuda"
    order = torch.from_numpy(np.argsort(-prob.numpy(), kind="stable").astype(np.int32)).cuda()
    keep_sorted = torch.zeros(128, dtype=torch.int32, device=dev)
    nkeep = torch.zeros(1, dtype=torch.int32, device=dev)
    ml = mlog[:, 0].contiguous().cuda()
    ops.mask_removal(boxes.cuda(), order, k, ml, ms, cls_idx.int().cuda(), H, W, 0.3,
                     torch.empty(8, H, W, dtype=torch.uint8, device=dev), 8, torch.empty(2 * k, dtype=torch.int32, device=dev),
                     torch.empty(k, dtype=torch.int32, device=dev), keep_sorted, nkeep)
    nk = int(nkeep.item())
    assert np.array_equal(keep_sorted[:nk].cpu().numpy(), keep_ref.numpy())          # keep list bit-exact
    pano = torch.empty(H, W, dtype=torch.int64, device=dev)
    sem = torch.empty(H, W, dtype=torch.int64, device=dev)
    ops.panoptic_fuse(to_nhwc(fcn_score), boxes.cuda(), cls_idx.int().cuda(), ml, ms, keep_sorted, nkeep, 128, NS, False, H, W,
                      pano, sem)
    # softmax-then-max == argmax except for float ties; inputs here are continuous random => identical maps
    assert torch.equal(sem.cpu(), sem_ref)
    assert torch.equal(pano.cpu(), pano_ref)
