"""Layer helpers built on the C-ABI kernels: packed convolutions, transposed convolutions as
stride-phase sub-convolutions, linear layers.  Precision follows the activation dtype:
bf16 activations -> tcgen05 tensor-core kernel, fp32 activations -> fp32 CUDA-core kernel (parity mode).
"""
import torch

from . import ops
from .ops import ACT_LRELU, ACT_NONE, ACT_RELU, ACT_SIGMOID, PackedConv  # noqa: F401


def empty_nhwc(n, h, w, c, dtype, device, c_align=8):
    """NHWC buffer whose pixel stride is padded to a multiple of `c_align` (TMA needs 16-byte strides)."""
    cs = (c + c_align - 1) // c_align * c_align
    buf = torch.empty(n, h, w, cs, dtype=dtype, device=device)
    return buf[..., :c] if cs != c else buf


class Conv:
    """nn.Conv2d (+ folded frozen-BN scale/bias, + activation) on NHWC buffers."""

    def __init__(self, weight, bias=None, stride=1, pad=0, act=ACT_NONE, slope=0.1, scale=None):
        self.pk = PackedConv(weight, bias, scale)
        self.stride, self.pad, self.act, self.slope = stride, pad, act, slope
        self.cout, self.cin = self.pk.cout, self.pk.cin
        self.k = self.pk.kh

    def out_hw(self, h, w):
        return ((h + 2 * self.pad - self.pk.kh) // self.stride + 1,
                (w + 2 * self.pad - self.pk.kw) // self.stride + 1)

    def tc_ok(self, x):
        # tensor-core path needs bf16, 16-byte aligned pixel stride and enough K to be worth a 64-wide chunk
        return (x.dtype == torch.bfloat16 and self.cin >= 16 and ops.vt(x).cs % 8 == 0
                and x.data_ptr() % 16 == 0)

    def _small_cin_tc(self):
        """bf16 + tiny cin (3..15 channels, first layers): explicit im2col to a K-padded matrix, then the
        tensor-core kernel runs it as a 1x1 conv.  Weight columns follow vps_im2col: k = (r*kw+s)*cin + ci."""
        if not hasattr(self, "_pk_cols"):
            w = self.pk.weight
            if self.pk.scale is not None:
                w = w * self.pk.scale.view(-1, 1, 1, 1)
            co, ci, kh, kw = w.shape
            kk = kh * kw * ci
            kpad = (kk + 63) // 64 * 64
            w2 = torch.zeros(co, kpad, 1, 1, dtype=torch.float32, device=w.device)
            w2[:, :kk, 0, 0] = w.permute(0, 2, 3, 1).reshape(co, kk)
            self._pk_cols = PackedConv(w2, self.pk.bias)
            self._kpad = kpad
        return self._pk_cols, self._kpad

    def __call__(self, x, y=None, act=None, res=None, res_after_act=False, out_scale=1.0, out_dtype=None):
        n, h, w, _ = x.shape
        oh, ow = self.out_hw(h, w)
        if y is None:
            y = empty_nhwc(n, oh, ow, self.cout, out_dtype or x.dtype, x.device)
        act = self.act if act is None else act
        if x.dtype == torch.bfloat16 and self.cin < 16 and self.pk.kh * self.pk.kw > 1:
            pk, kpad = self._small_cin_tc()
            cols = torch.empty(n, oh, ow, kpad, dtype=torch.bfloat16, device=x.device)
            ops.im2col(x, cols, self.pk.kh, self.pk.kw, self.stride, self.stride, self.pad, self.pad)
            ops.conv2d(cols, pk, y, act=act, slope=self.slope, res=res, res_after_act=res_after_act,
                       out_scale=out_scale, use_tc=True)
            return y
        ops.conv2d(x, self.pk, y, stride=self.stride, pad=self.pad, act=act,
                   slope=self.slope, res=res, res_after_act=res_after_act, out_scale=out_scale,
                   use_tc=self.tc_ok(x))
        return y


class _PhaseDeconv:
    """ConvTranspose2d(stride 2) as 4 stride-phase convolutions that write interleaved output pixels."""

    def __init__(self, weight_iohw, bias, k):
        # weight_iohw: [cin, cout, k, k] fp32 on device
        self.cin, self.cout = weight_iohw.shape[:2]
        self.k = k
        self.phases = []
        w_oihw = weight_iohw.permute(1, 0, 2, 3)
        for py in range(2):
            for px in range(2):
                if k == 4:     # padding 1: out[2q+p] taps -> see DESIGN.md "transposed convolutions"
                    ky = [3, 1] if py == 0 else [2, 0]
                    kx = [3, 1] if px == 0 else [2, 0]
                    sub = w_oihw[:, :, ky][:, :, :, kx].contiguous()
                    pad = (1 - py, 1 - px)
                else:          # k == 2, padding 0: out[2q+p] = x[q] * w[p]
                    sub = w_oihw[:, :, py:py + 1, px:px + 1].contiguous()
                    pad = (0, 0)
                self.phases.append((py, px, pad, PackedConv(sub, bias)))

    def __call__(self, x, y, act=ACT_NONE, slope=0.1, out_scale=1.0):
        n, h, w, _ = x.shape
        use_tc = (x.dtype == torch.bfloat16 and self.cin >= 16 and ops.vt(x).cs % 8 == 0
                  and x.data_ptr() % 16 == 0)
        for py, px, pad, pk in self.phases:
            ops.conv2d(x, pk, y, stride=1, pad_hw=pad, act=act, slope=slope, oh=h, ow=w,
                       omap=(2, py, 2, px), out_scale=out_scale, use_tc=use_tc)
        return y


def deconv4x4_s2(weight_iohw, bias):
    """nn.ConvTranspose2d(cin, cout, 4, 2, 1) -- FlowNet2 `deconv` / `upsampled_flow*` (submodules.py:33-37)."""
    return _PhaseDeconv(weight_iohw, bias, 4)


def deconv2x2_s2(weight_iohw, bias):
    """nn.ConvTranspose2d(cin, cout, 2, stride=2) -- FCNMaskHead.upsample (fcn_mask_head.py:66-71)."""
    return _PhaseDeconv(weight_iohw, bias, 2)


class Linear:
    """nn.Linear over rows: x viewed as NHWC [1,1,M,K]."""

    def __init__(self, weight, bias=None, act=ACT_NONE):
        self.pk = PackedConv(weight.view(weight.shape[0], weight.shape[1], 1, 1), bias)
        self.act = act
        self.cout = weight.shape[0]

    def __call__(self, x2d, y2d=None, out_dtype=None):
        m, k = x2d.shape
        if y2d is None:
            y2d = torch.empty(m, (self.cout + 7) // 8 * 8, dtype=out_dtype or x2d.dtype, device=x2d.device)[:, :self.cout]
        x4 = x2d.unsqueeze(0).unsqueeze(0)
        y4 = y2d.unsqueeze(0).unsqueeze(0)
        use_tc = x2d.dtype == torch.bfloat16 and x2d.stride(0) % 8 == 0 and x2d.data_ptr() % 16 == 0
        ops.conv2d(x4, self.pk, y4, act=self.act, use_tc=use_tc)
        return y2d
