"""ORACLE (test infrastructure): CPU restatement of the reference's FlowNet2 (inference only).

Follows mmdet/models/flow_modules/{flownet2.py:32-198, FlowNetC.py:13-128, FlowNetS.py:15-94,
FlowNetSD.py:11-106, FlowNetFusion.py:11-67, submodules.py:7-38}; batchNorm=False, eval mode
(encoders return only flow2).  Module/parameter names equal the reference's so state_dicts are
interchangeable.  The three CUDA extensions are replaced by oracle.ops.{correlation,resample2d,channelnorm}.
nn.Upsample(mode='bilinear') is evaluated with align_corners=False (torch-1.4 default).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn import init

from . import ops as O


def conv(cin, cout, kernel_size=3, stride=1):
    return nn.Sequential(nn.Conv2d(cin, cout, kernel_size, stride, (kernel_size - 1) // 2, bias=True),
                         nn.LeakyReLU(0.1, inplace=True))


def i_conv(cin, cout, kernel_size=3, stride=1, bias=True):
    return nn.Sequential(nn.Conv2d(cin, cout, kernel_size, stride, (kernel_size - 1) // 2, bias=bias))


def predict_flow(cin):
    return nn.Conv2d(cin, 2, 3, 1, 1, bias=True)


def deconv(cin, cout):
    return nn.Sequential(nn.ConvTranspose2d(cin, cout, 4, 2, 1, bias=True), nn.LeakyReLU(0.1, inplace=True))


def _flownet_init(mod):
    # flownet2.py:94-103: xavier_uniform weights, bias ~ U(0,1)
    for m in mod.modules():
        if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
            if m.bias is not None:
                init.uniform_(m.bias)
            init.xavier_uniform_(m.weight)


class _Decoder5(nn.Module):
    """Shared decoder wiring of FlowNetC / FlowNetS (refinement with raw concat prediction)."""

    def _make_decoder(self, flow_bias):
        self.deconv5 = deconv(1024, 512)
        self.deconv4 = deconv(1026, 256)
        self.deconv3 = deconv(770, 128)
        self.deconv2 = deconv(386, 64)
        self.predict_flow6 = predict_flow(1024)
        self.predict_flow5 = predict_flow(1026)
        self.predict_flow4 = predict_flow(770)
        self.predict_flow3 = predict_flow(386)
        self.predict_flow2 = predict_flow(194)
        self.upsampled_flow6_to_5 = nn.ConvTranspose2d(2, 2, 4, 2, 1, bias=flow_bias)
        self.upsampled_flow5_to_4 = nn.ConvTranspose2d(2, 2, 4, 2, 1, bias=flow_bias)
        self.upsampled_flow4_to_3 = nn.ConvTranspose2d(2, 2, 4, 2, 1, bias=flow_bias)
        self.upsampled_flow3_to_2 = nn.ConvTranspose2d(2, 2, 4, 2, 1, bias=flow_bias)

    def _decode(self, c2, c3, c4, c5, c6):
        flow6 = self.predict_flow6(c6)
        concat5 = torch.cat((c5, self.deconv5(c6), self.upsampled_flow6_to_5(flow6)), 1)
        flow5 = self.predict_flow5(concat5)
        concat4 = torch.cat((c4, self.deconv4(concat5), self.upsampled_flow5_to_4(flow5)), 1)
        flow4 = self.predict_flow4(concat4)
        concat3 = torch.cat((c3, self.deconv3(concat4), self.upsampled_flow4_to_3(flow4)), 1)
        flow3 = self.predict_flow3(concat3)
        concat2 = torch.cat((c2, self.deconv2(concat3), self.upsampled_flow3_to_2(flow3)), 1)
        return self.predict_flow2(concat2)


class FlowNetC(_Decoder5):
    def __init__(self):
        super().__init__()
        self.conv1 = conv(3, 64, 7, 2)
        self.conv2 = conv(64, 128, 5, 2)
        self.conv3 = conv(128, 256, 5, 2)
        self.conv_redir = conv(256, 32, 1, 1)
        self.conv3_1 = conv(473, 256)
        self.conv4 = conv(256, 512, stride=2)
        self.conv4_1 = conv(512, 512)
        self.conv5 = conv(512, 512, stride=2)
        self.conv5_1 = conv(512, 512)
        self.conv6 = conv(512, 1024, stride=2)
        self.conv6_1 = conv(1024, 1024)
        self._make_decoder(flow_bias=True)   # FlowNetC.py:48-51 bias=True

    def forward(self, x):
        x1, x2 = x[:, 0:3], x[:, 3:]
        c1a = self.conv1(x1); c2a = self.conv2(c1a); c3a = self.conv3(c2a)
        c1b = self.conv1(x2); c2b = self.conv2(c1b); c3b = self.conv3(c2b)
        corr = O.correlation(c3a, c3b, 20, 1, 20, 1, 2)          # FlowNetC.py:31,86
        corr = F.leaky_relu(corr, 0.1)                             # :33,87
        redir = self.conv_redir(c3a)
        c3_1 = self.conv3_1(torch.cat((redir, corr), 1))
        c4 = self.conv4_1(self.conv4(c3_1))
        c5 = self.conv5_1(self.conv5(c4))
        c6 = self.conv6_1(self.conv6(c5))
        return self._decode(c2a, c3_1, c4, c5, c6)


class FlowNetS(_Decoder5):
    def __init__(self, input_channels=12):
        super().__init__()
        self.conv1 = conv(input_channels, 64, 7, 2)
        self.conv2 = conv(64, 128, 5, 2)
        self.conv3 = conv(128, 256, 5, 2)
        self.conv3_1 = conv(256, 256)
        self.conv4 = conv(256, 512, stride=2)
        self.conv4_1 = conv(512, 512)
        self.conv5 = conv(512, 512, stride=2)
        self.conv5_1 = conv(512, 512)
        self.conv6 = conv(512, 1024, stride=2)
        self.conv6_1 = conv(1024, 1024)
        self._make_decoder(flow_bias=False)  # FlowNetS.py:45-48 bias=False

    def forward(self, x):
        c1 = self.conv1(x)
        c2 = self.conv2(c1)
        c3 = self.conv3_1(self.conv3(c2))
        c4 = self.conv4_1(self.conv4(c3))
        c5 = self.conv5_1(self.conv5(c4))
        c6 = self.conv6_1(self.conv6(c5))
        return self._decode(c2, c3, c4, c5, c6)


class FlowNetSD(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv0 = conv(6, 64)
        self.conv1 = conv(64, 64, stride=2)
        self.conv1_1 = conv(64, 128)
        self.conv2 = conv(128, 128, stride=2)
        self.conv2_1 = conv(128, 128)
        self.conv3 = conv(128, 256, stride=2)
        self.conv3_1 = conv(256, 256)
        self.conv4 = conv(256, 512, stride=2)
        self.conv4_1 = conv(512, 512)
        self.conv5 = conv(512, 512, stride=2)
        self.conv5_1 = conv(512, 512)
        self.conv6 = conv(512, 1024, stride=2)
        self.conv6_1 = conv(1024, 1024)
        self.deconv5 = deconv(1024, 512)
        self.deconv4 = deconv(1026, 256)
        self.deconv3 = deconv(770, 128)
        self.deconv2 = deconv(386, 64)
        self.inter_conv5 = i_conv(1026, 512)
        self.inter_conv4 = i_conv(770, 256)
        self.inter_conv3 = i_conv(386, 128)
        self.inter_conv2 = i_conv(194, 64)
        self.predict_flow6 = predict_flow(1024)
        self.predict_flow5 = predict_flow(512)
        self.predict_flow4 = predict_flow(256)
        self.predict_flow3 = predict_flow(128)
        self.predict_flow2 = predict_flow(64)
        self.upsampled_flow6_to_5 = nn.ConvTranspose2d(2, 2, 4, 2, 1)
        self.upsampled_flow5_to_4 = nn.ConvTranspose2d(2, 2, 4, 2, 1)
        self.upsampled_flow4_to_3 = nn.ConvTranspose2d(2, 2, 4, 2, 1)
        self.upsampled_flow3_to_2 = nn.ConvTranspose2d(2, 2, 4, 2, 1)

    def forward(self, x):
        c0 = self.conv0(x)
        c1 = self.conv1_1(self.conv1(c0))
        c2 = self.conv2_1(self.conv2(c1))
        c3 = self.conv3_1(self.conv3(c2))
        c4 = self.conv4_1(self.conv4(c3))
        c5 = self.conv5_1(self.conv5(c4))
        c6 = self.conv6_1(self.conv6(c5))
        flow6 = self.predict_flow6(c6)
        concat5 = torch.cat((c5, self.deconv5(c6), self.upsampled_flow6_to_5(flow6)), 1)
        flow5 = self.predict_flow5(self.inter_conv5(concat5))
        concat4 = torch.cat((c4, self.deconv4(concat5), self.upsampled_flow5_to_4(flow5)), 1)
        flow4 = self.predict_flow4(self.inter_conv4(concat4))
        concat3 = torch.cat((c3, self.deconv3(concat4), self.upsampled_flow4_to_3(flow4)), 1)
        flow3 = self.predict_flow3(self.inter_conv3(concat3))
        concat2 = torch.cat((c2, self.deconv2(concat3), self.upsampled_flow3_to_2(flow3)), 1)
        return self.predict_flow2(self.inter_conv2(concat2))


class FlowNetFusion(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv0 = conv(11, 64)
        self.conv1 = conv(64, 64, stride=2)
        self.conv1_1 = conv(64, 128)
        self.conv2 = conv(128, 128, stride=2)
        self.conv2_1 = conv(128, 128)
        self.deconv1 = deconv(128, 32)
        self.deconv0 = deconv(162, 16)
        self.inter_conv1 = i_conv(162, 32)
        self.inter_conv0 = i_conv(82, 16)
        self.predict_flow2 = predict_flow(128)
        self.predict_flow1 = predict_flow(32)
        self.predict_flow0 = predict_flow(16)
        self.upsampled_flow2_to_1 = nn.ConvTranspose2d(2, 2, 4, 2, 1)
        self.upsampled_flow1_to_0 = nn.ConvTranspose2d(2, 2, 4, 2, 1)

    def forward(self, x):
        c0 = self.conv0(x)
        c1 = self.conv1_1(self.conv1(c0))
        c2 = self.conv2_1(self.conv2(c1))
        flow2 = self.predict_flow2(c2)
        concat1 = torch.cat((c1, self.deconv1(c2), self.upsampled_flow2_to_1(flow2)), 1)
        flow1 = self.predict_flow1(self.inter_conv1(concat1))
        concat0 = torch.cat((c0, self.deconv0(concat1), self.upsampled_flow1_to_0(flow1)), 1)
        return self.predict_flow0(self.inter_conv0(concat0))


def _up4_bilinear(x):
    return F.interpolate(x, scale_factor=4, mode="bilinear", align_corners=False)


def _up4_nearest(x):
    return F.interpolate(x, scale_factor=4, mode="nearest")


class FlowNet2(nn.Module):
    """flownet2.py:32-198 (rgb_max=255, div_flow=20, fp16=False)."""

    def __init__(self, rgb_max=255.0, div_flow=20.0):
        super().__init__()
        self.rgb_max, self.div_flow = rgb_max, div_flow
        self.flownetc = FlowNetC()
        self.flownets_1 = FlowNetS()
        self.flownets_2 = FlowNetS()
        self.flownets_d = FlowNetSD()
        self.flownetfusion = FlowNetFusion()
        _flownet_init(self)
        for p in self.parameters():
            p.requires_grad = False

    def forward(self, inputs, taps=None):
        # inputs [B,3,2,H,W]
        rgb_mean = inputs.contiguous().view(inputs.size()[:2] + (-1,)).mean(dim=-1).view(inputs.size()[:2] + (1, 1, 1))
        x = (inputs - rgb_mean) / self.rgb_max
        x = torch.cat((x[:, :, 0], x[:, :, 1]), dim=1)
        img0, img1 = x[:, :3], x[:, 3:]

        c_flow2 = self.flownetc(x)
        c_flow = _up4_bilinear(c_flow2 * self.div_flow)
        res1 = O.resample2d(img1, c_flow)
        nd1 = O.channelnorm(img0 - res1)
        concat1 = torch.cat((x, res1, c_flow / self.div_flow, nd1), dim=1)

        s1_flow2 = self.flownets_1(concat1)
        s1_flow = _up4_bilinear(s1_flow2 * self.div_flow)
        res2 = O.resample2d(img1, s1_flow)
        nd2 = O.channelnorm(img0 - res2)
        concat2 = torch.cat((x, res2, s1_flow / self.div_flow, nd2), dim=1)

        s2_flow2 = self.flownets_2(concat2)
        s2_flow = _up4_nearest(s2_flow2 * self.div_flow)
        n_s2 = O.channelnorm(s2_flow)
        d_s2 = O.channelnorm(img0 - O.resample2d(img1, s2_flow))

        sd_flow2 = self.flownets_d(x)
        sd_flow = _up4_nearest(sd_flow2 / self.div_flow)
        n_sd = O.channelnorm(sd_flow)
        d_sd = O.channelnorm(img0 - O.resample2d(img1, sd_flow))

        concat3 = torch.cat((img0, sd_flow, s2_flow, n_sd, n_s2, d_sd, d_s2), dim=1)
        out = self.flownetfusion(concat3)
        if taps is not None:
            taps.update(c_flow2=c_flow2, s1_flow2=s1_flow2, s2_flow2=s2_flow2, sd_flow2=sd_flow2,
                        concat1=concat1, concat3=concat3)
        return out
