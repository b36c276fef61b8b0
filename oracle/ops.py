"""ORACLE (test infrastructure, NOT product code): CPU restatements of the reference's native ops.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  Everything here is plain PyTorch-CPU / numpy fp32 following the cited reference lines.
Layouts are the reference's (NCHW).

PARITY PIN STATUS: the reference ships no golden vectors for these ops.  The
restatements are pinned (a) against brute-force loops transcribed from the CUDA sources
(tests/test_oracle_ops.py) and (b) through the reference's own Python modules executed here with
these functions substituted for the absent CUDA extensions (tests/golden/make_golden.py).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------- correlation
def correlation(f1, f2, pad_size, kernel_size, max_displacement, stride1, stride2, corr_multiply=1):
    """correlation_cuda.forward: correlation_cuda.cc:10-87, correlation_cuda_kernel.cu:74-147.

    out[b, (tj+R)*D + (ti+R), y, x] = (1/(k*k*C)) * sum_c pad(f1)[b,c,y+md,x+md] * pad(f2)[b,c,y+md+tj*s2, x+md+ti*s2]
    kernel_size must be 1 (both call sites: FlowNetC.py:31, flow_modules.py:54-56).  fp32 accumulation.
    """
    assert kernel_size == 1 and stride1 == 1
    B, C, H, W = f1.shape
    R = max_displacement // stride2
    D = 2 * R + 1
    border = max_displacement  # kernel_radius 0
    pH, pW = H + 2 * pad_size, W + 2 * pad_size
    oH = int(math.ceil((pH - 2 * border) / float(stride1)))
    oW = int(math.ceil((pW - 2 * border) / float(stride1)))
    p1 = F.pad(f1.float(), (pad_size,) * 4)
    p2 = F.pad(f2.float(), (pad_size,) * 4)
    out = f1.new_zeros((B, D * D, oH, oW), dtype=torch.float32)
    a = p1[:, :, max_displacement:max_displacement + oH, max_displacement:max_displacement + oW]
    for tj in range(-R, R + 1):
        for ti in range(-R, R + 1):
            y0 = max_displacement + tj * stride2
            x0 = max_displacement + ti * stride2
            b = p2[:, :, y0:y0 + oH, x0:x0 + oW]
            out[:, (tj + R) * D + (ti + R)] = (a * b).sum(1) / float(C)
    return out


def correlation_loops(f1, f2, pad_size, max_displacement, stride2):
    """Literal triple-loop transcription of correlation_cuda_kernel.cu:74-147 (tiny inputs only)."""
    B, C, H, W = f1.shape
    R = max_displacement // stride2
    D = 2 * R + 1
    p1 = F.pad(f1.float(), (pad_size,) * 4).numpy()
    p2 = F.pad(f2.float(), (pad_size,) * 4).numpy()
    out = np.zeros((B, D * D, H, W), np.float32)
    for b in range(B):
        for y in range(H):
            for x in range(W):
                y1, x1 = y + max_displacement, x + max_displacement
                for tj in range(-R, R + 1):
                    for ti in range(-R, R + 1):
                        y2, x2 = y1 + tj * stride2, x1 + ti * stride2
                        acc = np.float32(0)
                        for c in range(C):
                            acc += p1[b, c, y1, x1] * p2[b, c, y2, x2]
                        out[b, (tj + R) * D + (ti + R), y, x] = acc / np.float32(C)
    return torch.from_numpy(out)


# ----------------------------------------------------------------------------- resample2d
def resample2d(src, flow):
    """resample2d_cuda.forward (kernel_size=1, bilinear): resample2d_kernel.cu:16-71.

    xf = x + flow[:,0], yf = y + flow[:,1]; taps clamped to the border, fractional weights kept.
    """
    B, C, H, W = src.shape
    _, _, oH, oW = flow.shape
    ys, xs = torch.meshgrid(torch.arange(oH, dtype=torch.float32), torch.arange(oW, dtype=torch.float32),
                            indexing="ij")
    xf = xs[None] + flow[:, 0].float()
    yf = ys[None] + flow[:, 1].float()
    alpha = xf - torch.floor(xf)
    beta = yf - torch.floor(yf)
    xL = torch.floor(xf).long().clamp(0, W - 1)
    xR = (torch.floor(xf).long() + 1).clamp(0, W - 1)
    yT = torch.floor(yf).long().clamp(0, H - 1)
    yB = (torch.floor(yf).long() + 1).clamp(0, H - 1)
    out = torch.zeros((B, C, oH, oW), dtype=torch.float32)
    srcf = src.float()
    for b in range(B):
        s = srcf[b]
        out[b] = ((1 - alpha[b]) * (1 - beta[b]) * s[:, yT[b], xL[b]] + alpha[b] * (1 - beta[b]) * s[:, yT[b], xR[b]]
                  + (1 - alpha[b]) * beta[b] * s[:, yB[b], xL[b]] + alpha[b] * beta[b] * s[:, yB[b], xR[b]])
    return out


def resample2d_loops(src, flow):
    """Literal transcription of kernel_resample2d_update_output (resample2d_kernel.cu:16-71)."""
    B, C, H, W = src.shape
    s = src.float().numpy()
    f = flow.float().numpy()
    out = np.zeros((B, C, flow.shape[2], flow.shape[3]), np.float32)
    for b in range(B):
        for y in range(flow.shape[2]):
            for x in range(flow.shape[3]):
                dx, dy = f[b, 0, y, x], f[b, 1, y, x]
                xf, yf = np.float32(x) + dx, np.float32(y) + dy
                alpha, beta = xf - np.floor(xf), yf - np.floor(yf)
                xL = max(min(int(np.floor(xf)), W - 1), 0)
                xR = max(min(int(np.floor(xf)) + 1, W - 1), 0)
                yT = max(min(int(np.floor(yf)), H - 1), 0)
                yB = max(min(int(np.floor(yf)) + 1, H - 1), 0)
                for c in range(C):
                    out[b, c, y, x] = ((1 - alpha) * (1 - beta) * s[b, c, yT, xL] + alpha * (1 - beta) * s[b, c, yT, xR]
                                       + (1 - alpha) * beta * s[b, c, yB, xL] + alpha * beta * s[b, c, yB, xR])
    return torch.from_numpy(out)


# ----------------------------------------------------------------------------- channelnorm
def channelnorm(x):
    """channelnorm_cuda.forward: channelnorm_kernel.cu:19-60 -- sqrt(sum_c x^2), no epsilon."""
    return torch.sqrt((x.float() ** 2).sum(1, keepdim=True))


# ----------------------------------------------------------------------------- WarpingLayer
def flow_warp(x, flow):
    """WarpingLayer.forward (flow_modules.py:126-148) under the pinned torch-1.4 defaults:
    grid = linspace(-1,1) + flow/((dim-1)/2); F.grid_sample(bilinear, zeros, align_corners=False)."""
    B, C, H, W = x.shape
    gx = torch.linspace(-1.0, 1.0, W).view(1, 1, 1, W).expand(B, 1, H, W)
    gy = torch.linspace(-1.0, 1.0, H).view(1, 1, H, 1).expand(B, 1, H, W)
    grid = torch.cat([gx, gy], 1)
    fg = torch.zeros_like(flow)
    fg[:, 0] = flow[:, 0] / ((flow.size(3) - 1.0) / 2.0)
    fg[:, 1] = flow[:, 1] / ((flow.size(2) - 1.0) / 2.0)
    grid = (grid + fg).permute(0, 2, 3, 1)
    return F.grid_sample(x, grid, mode="bilinear", padding_mode="zeros", align_corners=False)


# ----------------------------------------------------------------------------- DCNv1
def deform_im2col(x, offset, kh=3, kw=3, stride=1, pad=1, dil=1):
    """deformable_im2col (deform_conv_cuda_kernel.cu:83-113,189-242), deformable_group=1.

    Returns columns [B, C*kh*kw, H_out, W_out] ordered (c, i, j) as data_col (c_col = c*kh*kw + i*kw + j).
    """
    B, C, H, W = x.shape
    Ho = (H + 2 * pad - (dil * (kh - 1) + 1)) // stride + 1
    Wo = (W + 2 * pad - (dil * (kw - 1) + 1)) // stride + 1
    ys, xs = torch.meshgrid(torch.arange(Ho, dtype=torch.float32), torch.arange(Wo, dtype=torch.float32),
                            indexing="ij")
    cols = torch.zeros((B, C, kh * kw, Ho, Wo), dtype=torch.float32)
    xf = x.float()
    for i in range(kh):
        for j in range(kw):
            k = i * kw + j
            off_h = offset[:, 2 * k].float()
            off_w = offset[:, 2 * k + 1].float()
            h_im = ys[None] * stride - pad + i * dil + off_h
            w_im = xs[None] * stride - pad + j * dil + off_w
            inside = (h_im > -1) & (w_im > -1) & (h_im < H) & (w_im < W)
            h_low = torch.floor(h_im)
            w_low = torch.floor(w_im)
            lh, lw = h_im - h_low, w_im - w_low
            hh, hw = 1 - lh, 1 - lw
            h_low, w_low = h_low.long(), w_low.long()
            h_high, w_high = h_low + 1, w_low + 1
            for b in range(B):
                def tap(hi, wi, ok):
                    v = xf[b][:, hi.clamp(0, H - 1), wi.clamp(0, W - 1)]
                    return v * ok.float()
                v1 = tap(h_low[b], w_low[b], (h_low[b] >= 0) & (w_low[b] >= 0))
                v2 = tap(h_low[b], w_high[b], (h_low[b] >= 0) & (w_high[b] <= W - 1))
                v3 = tap(h_high[b], w_low[b], (h_high[b] <= H - 1) & (w_low[b] >= 0))
                v4 = tap(h_high[b], w_high[b], (h_high[b] <= H - 1) & (w_high[b] <= W - 1))
                val = hh[b] * hw[b] * v1 + hh[b] * lw[b] * v2 + lh[b] * hw[b] * v3 + lh[b] * lw[b] * v4
                cols[b, :, k] = val * inside[b].float()
    return cols.view(B, C * kh * kw, Ho, Wo)


def deform_conv(x, offset, weight, stride=1, pad=1, dil=1):
    """DeformConvFunction.forward (deform_conv.py:14-56 -> deform_conv_cuda.cpp:152-260): columns x weight,
    groups=1, no bias."""
    Co, Ci, kh, kw = weight.shape
    cols = deform_im2col(x, offset, kh, kw, stride, pad, dil)
    B, _, Ho, Wo = cols.shape
    out = torch.matmul(weight.view(Co, -1).float(), cols.view(B, Ci * kh * kw, Ho * Wo))
    return out.view(B, Co, Ho, Wo)


# ----------------------------------------------------------------------------- RoIAlign (legacy mmdet v1)
def roi_align(feat, rois, out_size, spatial_scale, sample_num):
    """ROIAlignForward (roi_align_kernel.cu:16-128).  feat [B,C,H,W], rois [n,5] (batch, x1,y1,x2,y2)."""
    B, C, H, W = feat.shape
    n = rois.shape[0]
    ph = pw = out_size
    out = torch.zeros((n, C, ph, pw), dtype=torch.float32)
    if n == 0:
        return out
    f = feat.float()
    rois = rois.float()
    bi = rois[:, 0].long()
    sw_ = rois[:, 1] * spatial_scale
    sh_ = rois[:, 2] * spatial_scale
    ew = (rois[:, 3] + 1) * spatial_scale
    eh = (rois[:, 4] + 1) * spatial_scale
    rw = torch.clamp(ew - sw_, min=0.0)
    rh = torch.clamp(eh - sh_, min=0.0)
    bh, bw = rh / ph, rw / pw
    assert sample_num > 0
    acc = torch.zeros((n, C, ph, pw), dtype=torch.float32)
    pidx_h = torch.arange(ph, dtype=torch.float32)
    pidx_w = torch.arange(pw, dtype=torch.float32)
    for iy in range(sample_num):
        y = sh_[:, None] + pidx_h[None] * bh[:, None] + (iy + 0.5) * bh[:, None] / sample_num  # [n,ph]
        for ix in range(sample_num):
            x = sw_[:, None] + pidx_w[None] * bw[:, None] + (ix + 0.5) * bw[:, None] / sample_num  # [n,pw]
            yy = y[:, :, None].expand(n, ph, pw)
            xx = x[:, None, :].expand(n, ph, pw)
            zero = (yy < -1.0) | (yy > H) | (xx < -1.0) | (xx > W)
            yc = torch.where(yy <= 0, torch.zeros_like(yy), yy)
            xc = torch.where(xx <= 0, torch.zeros_like(xx), xx)
            y_low = yc.long()
            x_low = xc.long()
            ytop = y_low >= H - 1
            xtop = x_low >= W - 1
            y_low = torch.where(ytop, torch.full_like(y_low, H - 1), y_low)
            x_low = torch.where(xtop, torch.full_like(x_low, W - 1), x_low)
            y_high = torch.where(ytop, y_low, y_low + 1)
            x_high = torch.where(xtop, x_low, x_low + 1)
            yc = torch.where(ytop, y_low.float(), yc)
            xc = torch.where(xtop, x_low.float(), xc)
            ly, lx = yc - y_low.float(), xc - x_low.float()
            hy, hx = 1.0 - ly, 1.0 - lx
            def idx(yi, xi):
                # gather f[b, :, yi, xi] for every (roi, ph, pw) without materialising f[bi]
                out_ = torch.empty((n, C, ph, pw), dtype=torch.float32)
                lin = yi * W + xi
                for b in torch.unique(bi).tolist():
                    sel = (bi == b).nonzero(as_tuple=True)[0]
                    fb = f[b].reshape(C, H * W)
                    out_[sel] = fb[:, lin[sel].reshape(-1)].reshape(C, sel.numel(), ph, pw).permute(1, 0, 2, 3)
                return out_
            lt, rt, lb, rb = idx(y_low, x_low), idx(y_low, x_high), idx(y_high, x_low), idx(y_high, x_high)
            w1, w2, w3, w4 = (hy * hx)[:, None], (hy * lx)[:, None], (ly * hx)[:, None], (ly * lx)[:, None]
            val = w1 * lt + w2 * rt + w3 * lb + w4 * rb
            val = val * (~zero)[:, None].float()
            acc += val
    out = acc / float(sample_num * sample_num)
    return out


# ----------------------------------------------------------------------------- NMS
def nms(dets, iou_thr):
    """nms_cuda (nms_kernel.cu:13-131): greedy, IoU with +1 extents, suppress when IoU > thr (strict).
    Processing order = stable descending score sort.  Returns (dets[inds], inds) with inds ascending
    (nms_wrapper.py:49)."""
    n = dets.shape[0]
    if n == 0:
        return dets, torch.zeros(0, dtype=torch.long)
    d = dets.detach().float().numpy()
    order = np.argsort(-d[:, 4], kind="stable")
    keep = nms_sorted_numpy(d[order], iou_thr)
    inds = np.sort(order[keep])
    inds = torch.from_numpy(inds.astype(np.int64))
    return dets[inds], inds


def nms_sorted_numpy(b, thr):
    """Greedy pass over boxes already sorted by score (descending). Returns kept positions."""
    n = b.shape[0]
    x1, y1, x2, y2 = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    areas = (x2 - x1 + np.float32(1)) * (y2 - y1 + np.float32(1))
    suppressed = np.zeros(n, bool)
    keep = []
    for i in range(n):
        if suppressed[i]:
            continue
        keep.append(i)
        if i + 1 >= n:
            break
        xx1 = np.maximum(x1[i], x1[i + 1:])
        yy1 = np.maximum(y1[i], y1[i + 1:])
        xx2 = np.minimum(x2[i], x2[i + 1:])
        yy2 = np.minimum(y2[i], y2[i + 1:])
        w = np.maximum(xx2 - xx1 + np.float32(1), np.float32(0))
        h = np.maximum(yy2 - yy1 + np.float32(1), np.float32(0))
        inter = w * h
        iou = inter / (areas[i] + areas[i + 1:] - inter)
        suppressed[i + 1:] |= iou > np.float32(thr)
    return np.array(keep, dtype=np.int64)


def gpu_nms_upsnet(dets, thresh):
    """UPSNet gpu_nms (gpu_nms.pyx:22-37 + nms_kernel.cu:40-150): host argsort descending, same bitmask
    kernel; returns order[keep] (score order).  Stable descending sort pins the reference's
    unspecified tie order."""
    d = np.asarray(dets, np.float32)
    order = np.argsort(-d[:, 4], kind="stable")
    keep = nms_sorted_numpy(d[order], thresh)
    return list(order[keep])
