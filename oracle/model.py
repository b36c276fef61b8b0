"""ORACLE (test infrastructure): CPU restatement of PanopticFuseTrack inference
(mmdet/models/detectors/panoptic_fusetrack.py:502-606) and every module it calls, in plain
PyTorch-CPU fp32.  Parameter names equal the reference's state_dict layout.

Reference lines followed are cited per class/function.  Device-specific calls of the reference
(.cuda(), torch.cuda.current_device()) are dropped; numerics are the torch-1.4 defaults the reference
pins (align_corners=False for bilinear interpolate / grid_sample).
"""
import math

import cv2
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops as O
from .flownet2 import FlowNet2

IMG_MEAN = [123.675, 116.28, 103.53]   # panoptic_fusetrack.py:92-93 / fusetrack.py:153-154
IMG_STD = [58.395, 57.12, 57.375]
CLASS_MAPPING = {1: 11, 2: 12, 3: 13, 4: 14, 5: 15, 6: 16, 7: 17, 8: 18}   # fusetrack.py:148
BBOX_REG_WEIGHTS = (10.0, 10.0, 5.0, 5.0)   # tools/config/config.py:47
MAX_DET = 100                                # tools/config/config.py:169


# ============================================================================ backbone
class Bottleneck(nn.Module):
    """resnet.py:86-266, style='pytorch' (stride on the 3x3)."""
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride=stride, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.downsample = downsample

    def forward(self, x):
        identity = x
        out = F.relu(self.bn1(self.conv1(x)))
        out = F.relu(self.bn2(self.conv2(out)))
        out = self.bn3(self.conv3(out))
        if self.downsample is not None:
            identity = self.downsample(x)
        return F.relu(out + identity)


class ResNet50(nn.Module):
    """resnet.py:333-526 with depth=50, out_indices=(0,1,2,3), norm_eval (frozen BN)."""

    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        inplanes = 64
        for i, (planes, blocks) in enumerate(zip((64, 128, 256, 512), (3, 4, 6, 3))):
            stride = 1 if i == 0 else 2
            layers = []
            ds = nn.Sequential(nn.Conv2d(inplanes, planes * 4, 1, stride=stride, bias=False),
                               nn.BatchNorm2d(planes * 4))
            layers.append(Bottleneck(inplanes, planes, stride, ds))
            inplanes = planes * 4
            for _ in range(1, blocks):
                layers.append(Bottleneck(inplanes, planes))
            setattr(self, "layer%d" % (i + 1), nn.Sequential(*layers))

    def forward(self, x):
        x = F.relu(self.bn1(self.conv1(x)))
        x = F.max_pool2d(x, 3, 2, 1)
        outs = []
        for i in range(4):
            x = getattr(self, "layer%d" % (i + 1))(x)
            outs.append(x)
        return tuple(outs)


class ConvModule(nn.Module):
    """conv_module.py:44-166 with norm_cfg=None: .conv (+ optional activation)."""

    def __init__(self, cin, cout, k, padding=0, activation="relu"):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, padding=padding, bias=True)
        self.activation = activation

    def forward(self, x):
        x = self.conv(x)
        if self.activation == "relu":
            x = F.relu(x)
        elif self.activation == "leaky_relu":
            x = F.leaky_relu(x, 0.1)
        return x


class FPN(nn.Module):
    """fpn.py:10-139, in_channels [256,512,1024,2048], out 256, num_outs 5, no activation."""

    def __init__(self):
        super().__init__()
        self.lateral_convs = nn.ModuleList([ConvModule(c, 256, 1, activation=None) for c in (256, 512, 1024, 2048)])
        self.fpn_convs = nn.ModuleList([ConvModule(256, 256, 3, padding=1, activation=None) for _ in range(4)])

    def forward(self, inputs):
        laterals = [l(inputs[i]) for i, l in enumerate(self.lateral_convs)]
        for i in range(3, 0, -1):
            laterals[i - 1] = laterals[i - 1] + F.interpolate(laterals[i], scale_factor=2, mode="nearest")
        outs = [self.fpn_convs[i](laterals[i]) for i in range(4)]
        outs.append(F.max_pool2d(outs[-1], 1, stride=2))     # fpn.py:126
        return tuple(outs)


# ============================================================================ BFPTcea
class OpticalFlowEstimatorCorr(nn.Module):
    """flow_modules.py:37-48."""

    def __init__(self, ch_in):
        super().__init__()

        def conv(a, b):
            return nn.Sequential(nn.Conv2d(a, b, 3, 1, 1, bias=True), nn.LeakyReLU(0.1, inplace=True))
        self.convs = nn.Sequential(conv(ch_in, 64), conv(64, 64), conv(64, 32), nn.Conv2d(32, 2, 3, 1, 1, bias=True))

    def forward(self, x):
        return self.convs(x)


class LiteFlowNetCorr(nn.Module):
    """flow_modules.py:50-73, search_range 4."""

    def __init__(self, in_ch):
        super().__init__()
        self.flow_estimator = OpticalFlowEstimatorCorr(in_ch + 81)

    def forward(self, x1, x2, flow_init):
        corr = O.correlation(x1.contiguous(), x2.contiguous(), 4, 1, 4, 1, 1)
        return self.flow_estimator(torch.cat([x1, corr, flow_init], dim=1))


class TCEA_Fusion(nn.Module):
    """tcea_modules.py:17-78 (nf=256, nframes=2, center=0)."""

    def __init__(self, nf=256, nframes=2, center=0):
        super().__init__()
        self.center = center
        self.tAtt_1 = nn.Conv2d(nf, nf, 3, 1, 1)
        self.tAtt_2 = nn.Conv2d(nf, nf, 3, 1, 1)
        self.fea_fusion = nn.Conv2d(nframes * nf, nf, 1, 1)
        self.sAtt_1 = nn.Conv2d(nframes * nf, nf, 1, 1)
        self.sAtt_2 = nn.Conv2d(nf * 2, nf, 1, 1)
        self.sAtt_3 = nn.Conv2d(nf, nf, 3, 1, 1)
        self.sAtt_4 = nn.Conv2d(nf, nf, 3, 1, 1)
        self.sAtt_add_1 = nn.Conv2d(nf, nf, 1, 1)
        self.sAtt_add_2 = nn.Conv2d(nf, nf, 1, 1)

    def forward(self, aligned_fea):
        B, N, C, H, W = aligned_fea.size()
        lrelu = lambda t: F.leaky_relu(t, 0.1)
        emb_ref = self.tAtt_2(aligned_fea[:, self.center].clone())
        emb = self.tAtt_1(aligned_fea.view(-1, C, H, W)).view(B, N, -1, H, W)
        cor_l = [torch.sum(emb[:, i] * emb_ref, 1).unsqueeze(1) for i in range(N)]
        cor_prob = torch.sigmoid(torch.cat(cor_l, dim=1))
        cor_prob = cor_prob.unsqueeze(2).repeat(1, 1, C, 1, 1).view(B, -1, H, W)
        aligned = aligned_fea.view(B, -1, H, W) * cor_prob
        fea = lrelu(self.fea_fusion(aligned))
        att = lrelu(self.sAtt_1(aligned))
        att_max = F.max_pool2d(att, 3, stride=2, padding=1)
        att_avg = F.avg_pool2d(att, 3, stride=2, padding=1)
        att = lrelu(self.sAtt_2(torch.cat([att_max, att_avg], dim=1)))
        att = lrelu(self.sAtt_3(att))
        att = F.interpolate(att, scale_factor=2, mode="bilinear", align_corners=False)
        att = self.sAtt_4(att)
        att_add = self.sAtt_add_2(lrelu(self.sAtt_add_1(att)))
        att = torch.sigmoid(att)
        return fea * att * 2 + att_add


class BFPTcea(nn.Module):
    """bfp_tcea.py:13-149 with refine_level=0, refine_type='conv', nframes=2, center=0."""

    def __init__(self, in_channels=256, num_levels=5):
        super().__init__()
        self.num_levels = num_levels
        self.liteflownet = LiteFlowNetCorr(in_channels + 2)
        self.tcea_fusion = TCEA_Fusion(in_channels, 2, 0)
        self.refine = ConvModule(in_channels, in_channels, 3, padding=1)   # default activation relu

    def gather(self, inputs):
        size = inputs[0].shape[2:]
        feats = [F.interpolate(inputs[i], size=size, mode="nearest") for i in range(self.num_levels)]
        return sum(feats) / len(feats)

    def forward(self, inputs, ref_inputs, flow_init, taps=None):
        bsf = self.gather(inputs)
        ref_bsf = self.gather(ref_inputs)
        warp_bsf = O.flow_warp(ref_bsf, flow_init)
        flow_fine = self.liteflownet(bsf, warp_bsf, flow_init)
        warp_bsf2 = O.flow_warp(warp_bsf, flow_fine)
        fused = self.tcea_fusion(torch.stack([bsf, warp_bsf2], dim=1))
        refined = self.refine(fused)
        outs = []
        for i in range(self.num_levels):
            residual = F.adaptive_max_pool2d(refined, output_size=inputs[i].shape[2:])
            outs.append(residual + inputs[i])
        if taps is not None:
            taps.update(bsf=bsf, ref_bsf=ref_bsf, warp_bsf=warp_bsf, flow_fine=flow_fine, warp_bsf2=warp_bsf2,
                        fused=fused, refined=refined)
        return tuple(outs)


# ============================================================================ UPSNetFPN
class DeformConv(nn.Module):
    """mmdet.ops.DeformConv (deform_conv.py:190-236): weight only, no bias."""

    def __init__(self, cin, cout, k=3, padding=1):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, cin, k, k))
        stdv = 1.0 / math.sqrt(cin * k * k)
        self.weight.data.uniform_(-stdv, stdv)
        self.padding = padding

    def forward(self, x, offset):
        return O.deform_conv(x, offset, self.weight, 1, self.padding, 1)


class DeformConvWithOffset(nn.Module):
    """deform_conv_with_offset.py:8-37."""

    def __init__(self, cin, cout):
        super().__init__()
        self.conv_offset = nn.Conv2d(cin, 18, 3, 1, 1)
        self.conv_offset.weight.data.zero_()
        self.conv_offset.bias.data.zero_()
        self.conv = DeformConv(cin, cout)

    def forward(self, x):
        return self.conv(x, self.conv_offset(x))


class UPSNetFPN(nn.Module):
    """upsnetFPN.py:14-81 (in 256, out 128, 4 levels, 19 classes)."""

    def __init__(self, cin=256, cout=128, num_classes=19):
        super().__init__()
        self.num_levels = 4
        self.num_classes = num_classes
        self.deform_convs = nn.ModuleList([nn.Sequential(
            DeformConvWithOffset(cin, cin), nn.GroupNorm(32, cin), nn.ReLU(inplace=True),
            DeformConvWithOffset(cin, cout), nn.GroupNorm(32, cout), nn.ReLU(inplace=True),
            DeformConvWithOffset(cout, cout), nn.GroupNorm(32, cout), nn.ReLU(inplace=True))])
        self.conv_pred = ConvModule(cout * 4, num_classes, 1, activation=None)

    def forward(self, inputs):
        px = [self.deform_convs[0](inputs[i]) for i in range(4)]
        p3 = F.interpolate(px[1], None, 2, mode="bilinear", align_corners=False)
        p4 = F.interpolate(px[2], None, 4, mode="bilinear", align_corners=False)
        p5 = F.interpolate(px[3], None, 8, mode="bilinear", align_corners=False)
        feat = torch.cat([px[0], p3, p4, p5], dim=1)
        fcn_score = self.conv_pred(feat)
        fcn_output = F.interpolate(fcn_score, scale_factor=4, mode="bilinear", align_corners=False)
        return fcn_output, fcn_score


# ============================================================================ RPN
def gen_base_anchors(base_size, scales=(8,), ratios=(0.5, 1.0, 2.0)):
    """anchor_generator.py:18-49 (scale_major, ctr None)."""
    w = h = base_size
    x_ctr, y_ctr = 0.5 * (w - 1), 0.5 * (h - 1)
    scales = torch.Tensor(list(scales))
    ratios = torch.Tensor(list(ratios))
    h_ratios = torch.sqrt(ratios)
    w_ratios = 1 / h_ratios
    ws = (w * w_ratios[:, None] * scales[None, :]).view(-1)
    hs = (h * h_ratios[:, None] * scales[None, :]).view(-1)
    return torch.stack([x_ctr - 0.5 * (ws - 1), y_ctr - 0.5 * (hs - 1), x_ctr + 0.5 * (ws - 1),
                        y_ctr + 0.5 * (hs - 1)], dim=-1).round()


def grid_anchors(base_anchors, featmap_size, stride):
    """anchor_generator.py:55-72."""
    feat_h, feat_w = featmap_size
    shift_x = torch.arange(0, feat_w) * stride
    shift_y = torch.arange(0, feat_h) * stride
    xx = shift_x.repeat(len(shift_y))
    yy = shift_y.view(-1, 1).repeat(1, len(shift_x)).view(-1)
    shifts = torch.stack([xx, yy, xx, yy], dim=-1).type_as(base_anchors)
    return (base_anchors[None, :, :] + shifts[:, None, :]).view(-1, 4)


def delta2bbox(rois, deltas, means=(0, 0, 0, 0), stds=(1, 1, 1, 1), max_shape=None, wh_ratio_clip=16 / 1000):
    """transforms.py:34-68."""
    means = deltas.new_tensor(means).repeat(1, deltas.size(1) // 4)
    stds = deltas.new_tensor(stds).repeat(1, deltas.size(1) // 4)
    d = deltas * stds + means
    dx, dy, dw, dh = d[:, 0::4], d[:, 1::4], d[:, 2::4], d[:, 3::4]
    max_ratio = np.abs(np.log(wh_ratio_clip))
    dw = dw.clamp(min=-max_ratio, max=max_ratio)
    dh = dh.clamp(min=-max_ratio, max=max_ratio)
    px = ((rois[:, 0] + rois[:, 2]) * 0.5).unsqueeze(1).expand_as(dx)
    py = ((rois[:, 1] + rois[:, 3]) * 0.5).unsqueeze(1).expand_as(dy)
    pw = (rois[:, 2] - rois[:, 0] + 1.0).unsqueeze(1).expand_as(dw)
    ph = (rois[:, 3] - rois[:, 1] + 1.0).unsqueeze(1).expand_as(dh)
    gw = pw * dw.exp()
    gh = ph * dh.exp()
    gx = torch.addcmul(px, pw, dx, value=1)
    gy = torch.addcmul(py, ph, dy, value=1)
    x1 = gx - gw * 0.5 + 0.5
    y1 = gy - gh * 0.5 + 0.5
    x2 = gx + gw * 0.5 - 0.5
    y2 = gy + gh * 0.5 - 0.5
    if max_shape is not None:
        x1 = x1.clamp(min=0, max=max_shape[1] - 1)
        y1 = y1.clamp(min=0, max=max_shape[0] - 1)
        x2 = x2.clamp(min=0, max=max_shape[1] - 1)
        y2 = y2.clamp(min=0, max=max_shape[0] - 1)
    return torch.stack([x1, y1, x2, y2], dim=-1).view_as(deltas)


def stable_topk(scores, k):
    """scores.topk(k) with the tie order pinned: descending score, ascending index."""
    order = torch.sort(scores, descending=True, stable=True)[1]
    return scores[order[:k]], order[:k]


class RPNHead(nn.Module):
    """rpn_head.py:12-104 + anchor_head.py:198-223; anchor scale 8, ratios .5/1/2, strides 4..64."""
    strides = (4, 8, 16, 32, 64)

    def __init__(self, cin=256, feat=256):
        super().__init__()
        self.rpn_conv = nn.Conv2d(cin, feat, 3, padding=1)
        self.rpn_cls = nn.Conv2d(feat, 3, 1)
        self.rpn_reg = nn.Conv2d(feat, 12, 1)
        self.base_anchors = [gen_base_anchors(s) for s in self.strides]

    def forward(self, feats):
        cls, reg = [], []
        for x in feats:
            x = F.relu(self.rpn_conv(x))
            cls.append(self.rpn_cls(x))
            reg.append(self.rpn_reg(x))
        return cls, reg

    def get_bboxes(self, cls_scores, bbox_preds, img_shape, cfg, taps=None):
        """get_bboxes_single (rpn_head.py:55-104), batch 1, nms_across_levels False, min_bbox_size 0."""
        mlvl = []
        for idx in range(len(cls_scores)):
            sc = cls_scores[idx][0]
            bp = bbox_preds[idx][0]
            anchors = grid_anchors(self.base_anchors[idx], sc.shape[-2:], self.strides[idx])
            scores = sc.permute(1, 2, 0).reshape(-1).sigmoid()
            bp = bp.permute(1, 2, 0).reshape(-1, 4)
            if cfg["nms_pre"] > 0 and scores.shape[0] > cfg["nms_pre"]:
                scores, topk_inds = stable_topk(scores, cfg["nms_pre"])
                bp = bp[topk_inds, :]
                anchors = anchors[topk_inds, :]
            else:
                # reference keeps original order here; NMS sorts internally
                pass
            proposals = delta2bbox(anchors, bp, (0, 0, 0, 0), (1, 1, 1, 1), img_shape)
            proposals = torch.cat([proposals, scores.unsqueeze(-1)], dim=-1)
            proposals, _ = O.nms(proposals, cfg["nms_thr"])
            proposals = proposals[:cfg["nms_post"], :]
            mlvl.append(proposals)
        proposals = torch.cat(mlvl, 0)
        num = min(cfg["max_num"], proposals.shape[0])
        _, topk_inds = stable_topk(proposals[:, 4], num)
        if taps is not None:
            taps["rpn_mlvl"] = mlvl
        return proposals[topk_inds, :]


# ============================================================================ RoI heads
def map_roi_levels(rois, num_levels, finest_scale=56):
    """single_level.py:54-73."""
    scale = torch.sqrt((rois[:, 3] - rois[:, 1] + 1) * (rois[:, 4] - rois[:, 2] + 1))
    lvls = torch.floor(torch.log2(scale / finest_scale + 1e-6))
    return lvls.clamp(min=0, max=num_levels - 1).long()


def roi_extract(feats, rois, out_size, strides=(4, 8, 16, 32), sample_num=2):
    """SingleRoIExtractor.forward (single_level.py:89-107)."""
    num_levels = len(strides)
    lvls = map_roi_levels(rois, num_levels)
    out = feats[0].new_zeros(rois.size(0), feats[0].shape[1], out_size, out_size)
    for i in range(num_levels):
        inds = lvls == i
        if inds.any():
            out[inds] = O.roi_align(feats[i], rois[inds, :], out_size, 1.0 / strides[i], sample_num)
    return out


class SharedFCBBoxHead(nn.Module):
    """convfc_bbox_head.py:132-185 with num_fcs=2, 9 classes."""

    def __init__(self, cin=256, fc=1024, num_classes=9):
        super().__init__()
        self.shared_fcs = nn.ModuleList([nn.Linear(cin * 49, fc), nn.Linear(fc, fc)])
        self.fc_cls = nn.Linear(fc, num_classes)
        self.fc_reg = nn.Linear(fc, 4 * num_classes)

    def forward(self, x):
        x = x.view(x.size(0), -1)
        for fc in self.shared_fcs:
            x = F.relu(fc(x))
        return self.fc_cls(x), self.fc_reg(x)


class TrackHead(nn.Module):
    """track_head.py:20-132 (num_fcs 2, match_coeff [1,2,10], bbox_dummy_iou 0, dynamic)."""

    def __init__(self, cin=256, fc=1024, match_coeff=(1.0, 2.0, 10.0)):
        super().__init__()
        self.fcs = nn.ModuleList([nn.Linear(cin * 49, fc), nn.Linear(fc, fc)])
        self.match_coeff = match_coeff

    def embed(self, x):
        x = x.view(x.size(0), -1)
        for idx, fc in enumerate(self.fcs):
            x = fc(x)
            if idx < len(self.fcs) - 1:
                x = F.relu(x)
        return x

    def forward(self, x, ref_x):
        prod = torch.mm(self.embed(x), self.embed(ref_x).t())
        return torch.cat([torch.zeros(prod.size(0), 1), prod], dim=1)

    def compute_comp_scores(self, match_ll, bbox_scores, bbox_ious, label_delta):
        bbox_ious = torch.cat((torch.zeros(bbox_ious.size(0), 1), bbox_ious), dim=1)
        label_delta = torch.cat((torch.ones(bbox_ious.size(0), 1), label_delta), dim=1)
        c = self.match_coeff
        return match_ll + c[0] * torch.log(bbox_scores) + c[1] * bbox_ious + c[2] * label_delta


class FCNMaskHead(nn.Module):
    """fcn_mask_head.py:14-103 (4 convs, deconv x2, 9 classes)."""

    def __init__(self, cin=256, num_classes=9):
        super().__init__()
        self.convs = nn.ModuleList([ConvModule(cin, 256, 3, padding=1) for _ in range(4)])
        self.upsample = nn.ConvTranspose2d(256, 256, 2, stride=2)
        self.conv_logits = nn.Conv2d(256, num_classes, 1)

    def forward(self, x):
        for c in self.convs:
            x = c(x)
        x = F.relu(self.upsample(x))
        return self.conv_logits(x)


def bbox_overlaps(b1, b2):
    """geometry.py:4-63, mode iou, not aligned."""
    rows, cols = b1.size(0), b2.size(0)
    if rows * cols == 0:
        return b1.new_zeros(rows, cols)
    lt = torch.max(b1[:, None, :2], b2[:, :2])
    rb = torch.min(b1[:, None, 2:], b2[:, 2:])
    wh = (rb - lt + 1).clamp(min=0)
    overlap = wh[:, :, 0] * wh[:, :, 1]
    a1 = (b1[:, 2] - b1[:, 0] + 1) * (b1[:, 3] - b1[:, 1] + 1)
    a2 = (b2[:, 2] - b2[:, 0] + 1) * (b2[:, 3] - b2[:, 1] + 1)
    return overlap / (a1[:, None] + a2 - overlap)


# ---------------------------------------------------------------------------- MaskROI (UPSNet)
def upsnet_bbox_transform(boxes, deltas, weights):
    """upsnet/bbox/bbox_transform.py:290-330 (numpy float32)."""
    if boxes.shape[0] == 0:
        return np.zeros((0, deltas.shape[1]), dtype=deltas.dtype)
    boxes = boxes.astype(deltas.dtype, copy=False)
    widths = boxes[:, 2] - boxes[:, 0] + 1.0
    heights = boxes[:, 3] - boxes[:, 1] + 1.0
    ctr_x = boxes[:, 0] + 0.5 * widths
    ctr_y = boxes[:, 1] + 0.5 * heights
    wx, wy, ww, wh = weights
    dx = deltas[:, 0::4] / wx
    dy = deltas[:, 1::4] / wy
    dw = deltas[:, 2::4] / ww
    dh = deltas[:, 3::4] / wh
    dw = np.minimum(dw, np.log(1000. / 16.))
    dh = np.minimum(dh, np.log(1000. / 16.))
    pred_ctr_x = dx * widths[:, np.newaxis] + ctr_x[:, np.newaxis]
    pred_ctr_y = dy * heights[:, np.newaxis] + ctr_y[:, np.newaxis]
    pred_w = np.exp(dw) * widths[:, np.newaxis]
    pred_h = np.exp(dh) * heights[:, np.newaxis]
    pred = np.zeros(deltas.shape, dtype=deltas.dtype)
    pred[:, 0::4] = pred_ctr_x - 0.5 * pred_w
    pred[:, 1::4] = pred_ctr_y - 0.5 * pred_h
    pred[:, 2::4] = pred_ctr_x + 0.5 * pred_w - 1
    pred[:, 3::4] = pred_ctr_y + 0.5 * pred_h - 1
    return pred


def upsnet_clip_boxes(boxes, im_shape):
    """upsnet/bbox/bbox_transform.py:45-60."""
    boxes[:, 0::4] = np.maximum(np.minimum(boxes[:, 0::4], im_shape[1] - 1), 0)
    boxes[:, 1::4] = np.maximum(np.minimum(boxes[:, 1::4], im_shape[0] - 1), 0)
    boxes[:, 2::4] = np.maximum(np.minimum(boxes[:, 2::4], im_shape[1] - 1), 0)
    boxes[:, 3::4] = np.maximum(np.minimum(boxes[:, 3::4], im_shape[0] - 1), 0)
    return boxes


def mask_roi(rois, bbox_delta, cls_prob, im_info, num_classes=9, nms_thresh=0.5, score_thresh=0.6,
             max_det=MAX_DET):
    """MaskROI.forward (mask_roi.py:37-147) with clip_boxes=True, class_agnostic=True, top_n unused.
    Returns (cls_prob [k], det_rois [k,5], cls_idx [k] in 1..8) or the dummy result (:136-142)."""
    rois_np = rois.detach().numpy()
    delta_np = bbox_delta.detach().numpy()
    prob_np = cls_prob.detach().numpy()
    proposal = upsnet_bbox_transform(rois_np[:, 1:], delta_np, BBOX_REG_WEIGHTS)
    proposal = upsnet_clip_boxes(proposal, im_info[0, :2])
    n = proposal.shape[0]
    cls_idx = [[c for _ in range(n)] for c in range(num_classes)]
    # class-agnostic fold (:60-74)
    prob_fold = prob_np[:, 1:].reshape((-1, 1))
    prob_fold = np.hstack((np.zeros_like(prob_fold), prob_fold))
    prob_t = cls_prob[:, 1:].contiguous().view(-1, 1)
    prob_t = torch.cat([torch.zeros_like(prob_t), prob_t], dim=1)
    proposal = proposal.reshape((n, -1, 4))[:, 1:, :].reshape((-1, 4))
    proposal = np.hstack((np.zeros_like(proposal), proposal))
    cls_idx = np.array(cls_idx).T[:, 1:].reshape((1, -1))
    cls_idx = np.vstack((np.zeros_like(cls_idx), cls_idx))
    j = 1
    inds = np.where(prob_fold[:, j] > score_thresh)[0]
    scores_j = prob_fold[inds, j]
    boxes_j = proposal[inds, j * 4:(j + 1) * 4]
    dets_j = np.hstack((boxes_j, scores_j[:, np.newaxis])).astype(np.float32)
    keep = [] if len(dets_j) == 0 else O.gpu_nms_upsnet(dets_j, nms_thresh)
    keep = np.array(keep, dtype=np.int64)
    nms_dets = dets_j[keep, :]
    scores_th = prob_t[torch.from_numpy(inds).long(), j][torch.from_numpy(keep).long()]
    cls_j = cls_idx[j][inds][keep]
    if max_det > 0 and len(nms_dets) > max_det:
        image_thresh = np.sort(nms_dets[:, -1])[-max_det]
        k2 = np.where(nms_dets[:, -1] >= image_thresh)[0]
        nms_dets = nms_dets[k2, :]
        scores_th = scores_th[torch.from_numpy(k2)]
        cls_j = cls_j[k2]
    if nms_dets.shape[0] == 0:
        return torch.ones(1), torch.zeros(1, 5), torch.zeros(1, dtype=torch.long)
    boxes = np.zeros((nms_dets.shape[0], 5))
    boxes[:, 1:] = nms_dets[:, :-1]
    return scores_th, torch.from_numpy(boxes).float(), torch.from_numpy(cls_j).long()


# ---------------------------------------------------------------------------- MaskRemoval / SegTerm
def mask_removal(mask_rois, cls_prob, mask_logit, cls_idx, im_shape, fraction_threshold=0.3):
    """MaskRemoval.forward (mask_removal.py:29-92).  mask_rois [k,4]; mask_logit [k,1,28,28]; cls_idx 1-based.
    Returns keep_inds (LongTensor) and mask_energy [1,k',H,W].  np.argsort made stable."""
    k = mask_rois.size(0)
    mask_energy = torch.zeros(1, k, im_shape[0], im_shape[1])
    rois = mask_rois.detach().numpy()
    prob = cls_prob.detach().numpy()
    logit_np = mask_logit.detach().numpy()
    cidx = cls_idx.detach().numpy()
    mask_image = np.zeros((int(np.max(cidx)),) + tuple(im_shape), dtype=np.uint8)
    sorted_inds = np.argsort(-prob, kind="stable")
    rois, prob, logit_np = rois[sorted_inds], prob[sorted_inds], logit_np[sorted_inds]
    cidx = cidx[sorted_inds] - 1
    if len(cidx) == 1 and cidx[0] == -1:
        return torch.zeros(1, dtype=torch.long), torch.zeros(1, 1, im_shape[0], im_shape[1])
    keep_inds = []
    ref_boxes = rois.astype(np.int32)
    frame_id = 0
    for i in range(sorted_inds.shape[0]):
        ref_box = ref_boxes[i, :].astype(np.int32)
        w = max(ref_box[2] - ref_box[0] + 1, 1)
        h = max(ref_box[3] - ref_box[1] + 1, 1)
        logit = cv2.resize(logit_np[i].squeeze(), (int(w), int(h)))
        mask = np.array(logit > 0, dtype=np.uint8)
        x_0 = max(ref_box[0], 0)
        x_1 = min(ref_box[2] + 1, im_shape[1])
        y_0 = max(ref_box[1], 0)
        y_1 = min(ref_box[3] + 1, im_shape[0])
        crop_mask = mask[(y_0 - ref_box[1]):(y_1 - ref_box[1]), (x_0 - ref_box[0]):(x_1 - ref_box[0])]
        mask_sum = crop_mask.sum()
        crop_img = mask_image[cidx[i]][y_0:y_1, x_0:x_1]
        if mask_sum == 0 or (np.logical_and(crop_img >= 1, crop_mask == 1).sum() / mask_sum > fraction_threshold):
            continue
        keep_inds.append(sorted_inds[i])
        mask_image[cidx[i]][y_0:y_1, x_0:x_1] += crop_mask
        mask_energy[0, frame_id, y_0:y_1, x_0:x_1] = torch.from_numpy(
            logit[(y_0 - ref_box[1]):(y_1 - ref_box[1]), (x_0 - ref_box[0]):(x_1 - ref_box[0])])
        frame_id += 1
    mask_energy = mask_energy[:, :len(keep_inds)]
    if len(keep_inds) == 0:
        return torch.zeros(1, dtype=torch.long), torch.zeros(1, 1, im_shape[0], im_shape[1])
    return torch.from_numpy(np.array(keep_inds, dtype=np.int64)), mask_energy


def seg_term(cls_indices, seg_score, boxes, box_scale=0.25, num_stuff=11):
    """SegTerm.forward (unary_logits.py:81-108)."""
    cls_np = cls_indices.numpy()
    stuff = seg_score[[0], :num_stuff]
    b = boxes.numpy()[:, 1:] * box_scale
    inst = torch.zeros((1, cls_np.shape[0], seg_score.shape[2], seg_score.shape[3]))
    for i in range(cls_np.shape[0]):
        if cls_np[i] == 0:
            continue
        y0 = int(b[i][1]); y1 = int(b[i][3].round() + 1)
        x0 = int(b[i][0]); x1 = int(b[i][2].round() + 1)
        inst[0, i, y0:y1, x0:x1] = seg_score[0, CLASS_MAPPING[int(cls_np[i])], y0:y1, x0:x1]
    return stuff, inst


# ============================================================================ detector
TEST_CFG_RPN = dict(nms_across_levels=False, nms_pre=1000, nms_post=1000, max_num=1000, nms_thr=0.7, min_bbox_size=0)


class PanopticFuseTrack(nn.Module):
    """panoptic_fusetrack.py:25-606, inference path, configured by configs/cityscapes/fusetrack.py."""

    def __init__(self):
        super().__init__()
        self.backbone = ResNet50()
        self.neck = FPN()
        self.extra_neck = BFPTcea()
        self.panopticFPN = UPSNetFPN()
        self.rpn_head = RPNHead()
        self.bbox_head = SharedFCBBoxHead()
        self.track_head = TrackHead()
        self.mask_head = FCNMaskHead()
        self.flownet2 = FlowNet2()
        self.prev_bboxes = self.prev_roi_feats = self.prev_det_labels = None
        self.eval()

    def extract_feat(self, img):
        return self.neck(self.backbone(img))

    def compute_flow(self, img, ref_img, scale_factor=0.25, taps=None):
        """:117-143 + flow_utils.denormalize (flow_utils.py:5-10)."""
        std = torch.tensor(IMG_STD).view(1, 3, 1, 1)
        mean = torch.tensor(IMG_MEAN).view(1, 3, 1, 1)
        rgb = img * std + mean
        ref_rgb = ref_img * std + mean
        rgbs = torch.stack([rgb, ref_rgb], dim=2)
        H, W = rgbs.shape[-2:]
        if H == 800 and W == 1600:
            rgbs = F.pad(rgbs, (0, 64, 0, 32))
        elif H == 200 and W == 400:
            rgbs = F.pad(rgbs, (0, 48, 0, 56))
        assert rgbs.size(-2) % 64 == 0 and rgbs.size(-1) % 64 == 0
        flow = self.flownet2(rgbs, taps)[:, :, :H, :W]
        if taps is not None:
            taps["flow_full"] = flow
        return F.interpolate(flow, scale_factor=scale_factor, mode="bilinear", align_corners=False) * scale_factor

    def track(self, det_bboxes, det_labels, det_roi_feats, cls_prob, is_first, taps=None):
        """:391-469."""
        if is_first or self.prev_bboxes is None:
            det_obj_ids = np.arange(det_bboxes.size(0))
            self.prev_bboxes = det_bboxes.clone()
            self.prev_roi_feats = det_roi_feats.clone()
            self.prev_det_labels = det_labels.clone()
            return det_obj_ids
        match_score = self.track_head(det_roi_feats, self.prev_roi_feats)
        match_logprob = F.log_softmax(match_score, dim=1)
        label_delta = (self.prev_det_labels == det_labels.view(-1, 1)).float()
        bbox_ious = bbox_overlaps(det_bboxes[:, :4], self.prev_bboxes[:, :4])
        comp_scores = self.track_head.compute_comp_scores(match_logprob, cls_prob.view(-1, 1), bbox_ious, label_delta)
        match_likelihood, match_ids = torch.max(comp_scores, dim=1)
        if taps is not None:
            taps["comp_scores"] = comp_scores
        match_likelihood = match_likelihood.numpy()
        match_ids = match_ids.numpy().astype(np.int32)
        det_obj_ids = np.ones((match_ids.shape[0]), dtype=np.int32) * (-1)
        best_match_scores = np.ones((self.prev_bboxes.size(0))) * (-100)
        best_match_ids = np.ones((self.prev_bboxes.size(0)), dtype=np.int32) * (-1)
        for idx, match_id in enumerate(match_ids):
            if match_id == 0:
                det_obj_ids[idx] = self.prev_roi_feats.size(0)
                self.prev_roi_feats = torch.cat((self.prev_roi_feats, det_roi_feats[idx][None]), dim=0)
                self.prev_bboxes = torch.cat((self.prev_bboxes, det_bboxes[idx][None]), dim=0)
                self.prev_det_labels = torch.cat((self.prev_det_labels, det_labels[idx][None]), dim=0)
            else:
                obj_id = match_id - 1
                match_score_i = match_likelihood[idx]
                if match_score_i > best_match_scores[obj_id]:
                    det_obj_ids[idx] = obj_id
                    if best_match_ids[obj_id] >= 0:
                        det_obj_ids[best_match_ids[obj_id]] = -1
                    best_match_scores[obj_id] = match_score_i
                    best_match_ids[obj_id] = idx
                    self.prev_roi_feats[obj_id] = det_roi_feats[idx]
                    self.prev_bboxes[obj_id] = det_bboxes[idx]
        for idx, det_obj_id in enumerate(det_obj_ids):
            if det_obj_id >= 0:
                continue
            det_obj_ids[idx] = self.prev_roi_feats.size(0)
            self.prev_roi_feats = torch.cat((self.prev_roi_feats, det_roi_feats[idx][None]), dim=0)
            self.prev_bboxes = torch.cat((self.prev_bboxes, det_bboxes[idx][None]), dim=0)
            self.prev_det_labels = torch.cat((self.prev_det_labels, det_labels[idx][None]), dim=0)
        return det_obj_ids

    @torch.no_grad()
    def simple_test(self, img, img_meta, ref_img, taps=None):
        """:502-606.  img, ref_img [1,3,H,W] fp32 normalised; img_meta dict with iid, img_shape."""
        im_info = np.array([[float(img.shape[2]), float(img.shape[3]), 1.0]])
        flow = self.compute_flow(img.clone(), ref_img.clone(), 0.25, taps)
        x = self.extract_feat(img)
        ref_x = self.extract_feat(ref_img)
        xf = self.extra_neck(x, ref_x, flow, taps)
        fcn_output, fcn_score = self.panopticFPN(xf[0:4])
        cls_scores, bbox_preds = self.rpn_head(xf)
        proposals = self.rpn_head.get_bboxes(cls_scores, bbox_preds, img_meta["img_shape"], TEST_CFG_RPN, taps)
        # ---- simple_test_bboxes (:358-471)
        rois = torch.cat([proposals.new_zeros(proposals.size(0), 1), proposals[:, :4]], dim=-1)   # bbox2roi
        roi_feats = roi_extract(xf[:4], rois, 7)
        cls_score, bbox_pred = self.bbox_head(roi_feats)
        is_first = (img_meta["iid"] % 10000) == 1
        cls_prob_all = F.softmax(cls_score, dim=1)
        cls_prob, det_rois, cls_idx = mask_roi(rois, bbox_pred, cls_prob_all, im_info)
        det_labels = cls_idx - 1
        det_roi_feats = roi_extract(xf[:4], det_rois, 7)
        det_bboxes = det_rois[:, 1:]
        det_obj_ids = self.track(det_bboxes, det_labels, det_roi_feats, cls_prob, is_first, taps)
        # ---- panoptic head (:559-597)
        mask_feats = roi_extract(xf[:4], det_rois, 14)
        mask_score = self.mask_head(mask_feats)
        nobj, _, mh, mw = mask_score.shape
        mask_score = mask_score.gather(1, cls_idx.view(-1, 1, 1, 1).expand(-1, -1, mh, mw))
        keep_inds, mask_logits = mask_removal(det_rois[:, 1:], cls_prob, mask_score, cls_idx, tuple(fcn_output.shape[2:]))
        det_obj_ids_t = torch.from_numpy(np.asarray(det_obj_ids))
        mask_rois = det_rois[keep_inds]
        cls_idx_k = cls_idx[keep_inds]
        det_labels_k = det_labels[keep_inds]
        det_obj_ids_k = det_obj_ids_t[keep_inds]
        cls_prob_k = cls_prob[keep_inds]
        stuff, inst = seg_term(cls_idx_k, fcn_output, mask_rois * 4.0)
        panoptic_logits = torch.cat([stuff, inst + mask_logits], dim=1)
        panoptic_output = torch.max(F.softmax(panoptic_logits, dim=1), dim=1)[1]
        sem_output = torch.max(F.softmax(fcn_output, dim=1), dim=1)[1]
        h0, w0 = img_meta["img_shape"][:2]
        sem_output = sem_output[:, 0:h0, 0:w0]
        panoptic_output = panoptic_output[:, 0:h0, 0:w0]
        if taps is not None:
            taps.update(flow=flow, fpn=x, ref_fpn=ref_x, fused=xf, fcn_score=fcn_score, fcn_output=fcn_output,
                        rpn_cls=cls_scores, rpn_reg=bbox_preds, proposals=proposals, roi_feats=roi_feats,
                        cls_score=cls_score, bbox_pred=bbox_pred, det_rois=det_rois, cls_idx=cls_idx,
                        cls_prob=cls_prob, det_roi_feats=det_roi_feats, mask_score=mask_score,
                        keep_inds=keep_inds, panoptic_logits=panoptic_logits, det_obj_ids_all=det_obj_ids_t)
        bbox_results = {}
        for bbox, label, obj_id in zip(det_bboxes.numpy(), det_labels.numpy(), np.asarray(det_obj_ids)):
            if obj_id >= 0:
                bbox_results[int(obj_id)] = {"bbox": bbox, "label": label}
        pano_results = {
            "fcn_outputs": sem_output,
            "panoptic_cls_inds": cls_idx_k,
            "panoptic_cls_prob": cls_prob_k,
            "panoptic_det_labels": det_labels_k,
            "panoptic_det_obj_ids": det_obj_ids_k,
            "panoptic_outputs": panoptic_output,
        }
        segm_result = [[] for _ in range(8)]   # simple_test_mask :484-485 always returns empties
        return bbox_results, segm_result, pano_results
