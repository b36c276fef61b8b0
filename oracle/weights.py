"""ORACLE (test infrastructure): deterministic synthetic weight sets for the FuseTrack oracle
(SURVEY.md 8d).  There are no trained checkpoints offline; these exercise every code path.

  "A": the reference's init rules (kaiming / xavier / normal as cited in SURVEY A.13).  bn3.gamma = 0
       and score_thresh 0.6 mean residual branches vanish and nothing is detected -> plumbing only.
  "B": dense-random: every conv/linear kaiming-uniform-ish, BN gamma~U(.5,1.5), beta~N(0,.1), running
       stats random, non-zero DCN offsets, FlowNet2 xavier with small biases.
  "C": B + detection forcing: bbox_head.fc_cls scaled up / fc_reg scaled down so that >= a few dozen
       RoIs pass the 0.6 threshold and the mask / tracking / fusion stages see real instances.
"""
import math

import torch
import torch.nn as nn


def _kaiming_uniform(w, gain=1.0):
    fan_in = w[0].numel()
    bound = gain * math.sqrt(3.0 / fan_in)
    w.uniform_(-bound, bound)


@torch.no_grad()
def init_weights(model, kind="C", seed=0):
    g = torch.Generator().manual_seed(seed)

    def U(t, a, b):
        t.copy_(torch.rand(t.shape, generator=g) * (b - a) + a)

    def N(t, mean, std):
        t.copy_(torch.randn(t.shape, generator=g) * std + mean)

    if kind == "A":
        for name, m in model.named_modules():
            if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d, nn.Linear)):
                if name.startswith("flownet2"):
                    fan_in = m.weight[0].numel() if not isinstance(m, nn.ConvTranspose2d) else m.weight.shape[0] * m.weight[0, 0].numel()
                    fan_out = m.weight.shape[0] * m.weight[0, 0].numel() if not isinstance(m, nn.ConvTranspose2d) else m.weight[0].numel()
                    b = math.sqrt(6.0 / (fan_in + fan_out))
                    U(m.weight, -b, b)
                    if m.bias is not None:
                        U(m.bias, 0, 1)
                elif name.startswith("backbone"):
                    fan_out = m.weight.shape[0] * m.weight[0, 0].numel()
                    N(m.weight, 0, math.sqrt(2.0 / fan_out))
                elif name.startswith(("rpn_head", "track_head")):
                    N(m.weight, 0, 0.01); m.bias.zero_()
                elif name == "bbox_head.fc_cls":
                    N(m.weight, 0, 0.01); m.bias.zero_()
                elif name == "bbox_head.fc_reg":
                    N(m.weight, 0, 0.001); m.bias.zero_()
                else:   # xavier uniform (FPN, BFPTcea, UPSNetFPN, shared fcs); kaiming for mask head is close enough
                    fan_in = m.weight[0].numel()
                    fan_out = m.weight.shape[0] * (m.weight[0, 0].numel() if m.weight.dim() > 2 else 1)
                    b = math.sqrt(6.0 / (fan_in + fan_out))
                    U(m.weight, -b, b)
                    if m.bias is not None:
                        m.bias.zero_()
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.fill_(1); m.bias.zero_(); m.running_mean.zero_(); m.running_var.fill_(1)
                if name.endswith("bn3"):
                    m.weight.zero_()
        return model

    for name, m in model.named_modules():
        if isinstance(m, (nn.Conv2d, nn.Linear)):
            _g = 1.4 if not name.startswith("flownet2") else 1.0
            fan_in = m.weight[0].numel()
            bound = _g * math.sqrt(3.0 / fan_in)
            U(m.weight, -bound, bound)
            if m.bias is not None:
                N(m.bias, 0, 0.05)
        elif isinstance(m, nn.ConvTranspose2d):
            fan_in = m.weight.shape[0] * m.weight[0, 0].numel() / 4.0   # stride-2: ~1/4 of taps hit each output
            bound = math.sqrt(3.0 / fan_in)
            U(m.weight, -bound, bound)
            if m.bias is not None:
                N(m.bias, 0, 0.05)
        elif isinstance(m, nn.BatchNorm2d):
            U(m.weight, 0.5, 1.5); N(m.bias, 0, 0.1); N(m.running_mean, 0, 0.1); U(m.running_var, 0.5, 1.5)
            if name.endswith("bn3"):
                m.weight.mul_(0.5)      # keep the residual trunk from blowing up over 16 blocks
        elif isinstance(m, nn.GroupNorm):
            U(m.weight, 0.5, 1.5); N(m.bias, 0, 0.1)
        elif m.__class__.__name__ == "DeformConv":
            fan_in = m.weight[0].numel()
            bound = 1.4 * math.sqrt(3.0 / fan_in)
            U(m.weight, -bound, bound)
    # DCN offsets: moderate non-zero offsets (a few pixels)
    for name, m in model.named_modules():
        if name.endswith("conv_offset"):
            N(m.weight, 0, 0.02); N(m.bias, 0, 0.5)
    # FlowNet2 predicts flows through 5 stacked nets; keep predictions O(1 px)
    for name, m in model.named_modules():
        if name.startswith("flownet2") and "predict_flow" in name:
            m.weight.mul_(0.2)
    # RPN: spread objectness and keep deltas moderate
    model.rpn_head.rpn_cls.weight.mul_(2.0)
    model.rpn_head.rpn_reg.weight.mul_(0.3)
    if kind == "C":
        model.bbox_head.fc_cls.weight.mul_(6.0)
        model.bbox_head.fc_cls.bias[0] -= 1.0
        model.bbox_head.fc_reg.weight.mul_(0.3)
        model.mask_head.conv_logits.weight.mul_(3.0)
    return model


@torch.no_grad()
def calibrate(model, size=(128, 256), seed=123):
    """Data-dependent rescaling (LSUV-style) so the synthetic network is numerically non-degenerate:
    O(1) pyramid features, sub-pixel..pixel flows, un-saturated class scores, O(1) mask logits and
    tracker scores.  Deterministic (fixed seed, CPU fp32).  Only linear output layers are rescaled."""
    import torch.nn.functional as F
    from . import model as M
    g = torch.Generator().manual_seed(seed)
    H, W = size
    img = torch.randn(1, 3, H, W, generator=g)
    ref = torch.roll(img, shifts=(1, 2), dims=(2, 3)) + 0.05 * torch.randn(1, 3, H, W, generator=g)

    def scale_(mod, s):
        mod.weight.mul_(s)
        if getattr(mod, "bias", None) is not None:
            mod.bias.mul_(s)

    # FlowNet2: final flow ~ 1.5 px RMS
    taps = {}
    model.compute_flow(img.clone(), ref.clone(), 0.25, taps)
    rms = taps["flow_full"].pow(2).mean().sqrt().item()
    scale_(model.flownet2.flownetfusion.predict_flow0, 1.5 / max(rms, 1e-6))
    # pyramid features ~ unit RMS at every level
    c = model.backbone(img)
    for i, l in enumerate(model.neck.lateral_convs):
        rms = l(c[i]).pow(2).mean().sqrt().item()
        scale_(l.conv, 1.0 / max(rms, 1e-6))
    for _ in range(2):
        p = model.neck(c)
        for i in range(4):
            rms = p[i].pow(2).mean().sqrt().item()
            scale_(model.neck.fpn_convs[i].conv, 1.0 / max(rms, 1e-6))
    x = model.neck(c)
    rx = model.neck(model.backbone(ref))
    flow = torch.zeros(1, 2, H // 4, W // 4)
    en = model.extra_neck
    bsf, rbsf = en.gather(x), en.gather(rx)
    ff = en.liteflownet(bsf, M.O.flow_warp(rbsf, flow), flow)
    scale_(en.liteflownet.flow_estimator.convs[3], 0.7 / max(ff.pow(2).mean().sqrt().item(), 1e-6))
    taps = {}
    en(x, rx, flow, taps)
    scale_(en.tcea_fusion.tAtt_1, 0.25 / max(en.tcea_fusion.tAtt_1(bsf).pow(2).mean().sqrt().item(), 1e-6))
    scale_(en.tcea_fusion.tAtt_2, 0.25 / max(en.tcea_fusion.tAtt_2(bsf).pow(2).mean().sqrt().item(), 1e-6))
    taps = {}
    en(x, rx, flow, taps)
    scale_(en.tcea_fusion.sAtt_4, 1.0 / max(taps["fused"].pow(2).mean().sqrt().item(), 1e-6) ** 0.5)
    taps = {}
    en(x, rx, flow, taps)
    scale_(en.refine.conv, 0.7 / max(taps["refined"].pow(2).mean().sqrt().item(), 1e-6))
    xf = en(x, rx, flow)
    # semantic logits ~ 2 RMS
    _, fs = model.panopticFPN(xf[:4])
    scale_(model.panopticFPN.conv_pred.conv, 2.0 / max(fs.pow(2).mean().sqrt().item(), 1e-6))
    # RPN: logits ~1.5, deltas ~0.3
    cls, reg = model.rpn_head(xf)
    scale_(model.rpn_head.rpn_cls, 1.5 / max(cls[0].pow(2).mean().sqrt().item(), 1e-6))
    scale_(model.rpn_head.rpn_reg, 0.3 / max(reg[0].pow(2).mean().sqrt().item(), 1e-6))
    cls, reg = model.rpn_head(xf)
    props = model.rpn_head.get_bboxes(cls, reg, (H, W, 3), M.TEST_CFG_RPN)
    rois = torch.cat([props.new_zeros(props.size(0), 1), props[:, :4]], dim=-1)
    rf = M.roi_extract(xf[:4], rois, 7)
    bh = model.bbox_head
    h0 = F.relu(bh.shared_fcs[0](rf.view(rf.size(0), -1)))
    scale_(bh.shared_fcs[0], 1.0 / max(h0.pow(2).mean().sqrt().item(), 1e-6))
    h1 = F.relu(bh.shared_fcs[1](F.relu(bh.shared_fcs[0](rf.view(rf.size(0), -1)))))
    scale_(bh.shared_fcs[1], 1.0 / max(h1.pow(2).mean().sqrt().item(), 1e-6))
    cs, bp = bh(rf)
    scale_(bh.fc_cls, 4.0 / max(cs.pow(2).mean().sqrt().item(), 1e-6))
    scale_(bh.fc_reg, 0.5 / max(bp.pow(2).mean().sqrt().item(), 1e-6))
    th = model.track_head
    e0 = F.relu(th.fcs[0](rf.view(rf.size(0), -1)))
    scale_(th.fcs[0], 1.0 / max(e0.pow(2).mean().sqrt().item(), 1e-6))
    e1 = th.embed(rf)
    scale_(th.fcs[1], (3.0 / max((e1 @ e1.t()).pow(2).mean().sqrt().item(), 1e-6)) ** 0.5)
    mf = M.roi_extract(xf[:4], rois[:32], 14)
    ms = model.mask_head(mf)
    scale_(model.mask_head.conv_logits, 2.0 / max(ms.pow(2).mean().sqrt().item(), 1e-6))
    return model


def make_model(kind="C", seed=0, calibrated=True, cache_dir="/tmp/vps_oracle_weights"):
    """Oracle model with synthetic weights; the calibrated state_dict is cached on disk."""
    import os
    from .model import PanopticFuseTrack
    m = PanopticFuseTrack()
    path = os.path.join(cache_dir, "w_%s_%d_%d.pt" % (kind, seed, int(calibrated)))
    if os.path.exists(path):
        m.load_state_dict(torch.load(path))
        return m
    init_weights(m, kind, seed)
    if calibrated and kind != "A":
        calibrate(m)
    try:
        os.makedirs(cache_dir, exist_ok=True)
        torch.save(m.state_dict(), path)
    except OSError:
        pass
    return m
