"""Python-side wrappers of the C-ABI kernels.

Tensors handed to these functions are torch CUDA tensors used purely as device-memory handles
(pointer + shape); all arithmetic happens in libvps_b200.so.  Activations are NHWC: a tensor of shape
[N, H, W, C] whose last dim is contiguous; a channel slice `buf[..., a:b]` of a wider buffer is a
valid view (pixel stride = buf.shape[-1]).
"""
import ctypes as C

import torch

from . import _lib
from ._lib import (ACT_LRELU, ACT_NONE, ACT_RELU, ACT_SIGMOID, VPS_BF16, VPS_F32, VpsConvArgs,
                   VpsTensor, check, lib)

_DT = {torch.float32: VPS_F32, torch.bfloat16: VPS_BF16}


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def vt(t):
    """torch NHWC tensor (possibly a channel-slice view) -> VpsTensor."""
    if t is None:
        return VpsTensor(None, 0, 0, 0, 0, 0, 0)
    assert t.is_cuda and t.dim() == 4, "expected a CUDA NHWC tensor, got %s" % (tuple(t.shape),)
    n, h, w, c = t.shape
    assert t.stride(3) == 1 or c == 1
    if w > 1:
        cs = t.stride(2)
    elif h > 1:
        cs = t.stride(1)
    elif n > 1:
        cs = t.stride(0)
    else:
        cs = max(c, 1)
    if h > 1:
        assert t.stride(1) == w * cs, "rows must be dense: %s %s" % (t.shape, t.stride())
    if n > 1:
        assert t.stride(0) == h * w * cs, "images must be dense: %s %s" % (t.shape, t.stride())
    return VpsTensor(t.data_ptr(), n, h, w, c, cs, _DT[t.dtype])


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(None)


def launch_count():
    return int(lib().vps_launch_count())


# ------------------------------------------------------------------ weights
class PackedConv:
    """Packed weights of one convolution for both kernels' layouts (built lazily per precision)."""

    def __init__(self, weight, bias=None, scale=None, transposed=False):
        # weight: OIHW fp32 CUDA tensor (IOHW if transposed); scale: per-cout multiplier folded in
        self.transposed = transposed
        if transposed:
            self.cin, self.cout, self.kh, self.kw = weight.shape
        else:
            self.cout, self.cin, self.kh, self.kw = weight.shape
        self.weight = weight.contiguous().float()
        self.scale = scale.contiguous().float() if scale is not None else None
        self.bias = bias.contiguous().float() if bias is not None else None
        self._tc = None
        self._simt = None

    def tc(self):
        if self._tc is None:
            nbytes = lib().vps_packed_tc_bytes(self.cout, self.cin, self.kh, self.kw)
            buf = torch.empty(nbytes // 2, dtype=torch.bfloat16, device=self.weight.device)
            check(lib().vps_pack_weights_tc(_ptr(self.weight), _ptr(self.scale), _ptr(buf), self.cout, self.cin,
                                            self.kh, self.kw, int(self.transposed), stream()), "pack_weights_tc")
            self._tc = buf
        return self._tc

    def simt(self):
        if self._simt is None:
            buf = torch.empty(self.kh * self.kw * self.cin * self.cout, dtype=torch.float32,
                              device=self.weight.device)
            check(lib().vps_pack_weights_simt(_ptr(self.weight), _ptr(self.scale), _ptr(buf), self.cout, self.cin,
                                              self.kh, self.kw, int(self.transposed), stream()),
                  "pack_weights_simt")
            self._simt = buf
        return self._simt


def conv2d(x, pw, y, stride=1, pad=0, act=ACT_NONE, slope=0.1, res=None, res_after_act=False, out_scale=1.0,
           oh=None, ow=None, omap=(1, 0, 1, 0), pad_hw=None, use_tc=None, w_override=None, khw=None):
    """y <- conv(x) with fused bias/activation/residual.  `omap` = (oy_mul, oy_off, ox_mul, ox_off)."""
    a = VpsConvArgs()
    a.x, a.y, a.res = vt(x), vt(y), vt(res)
    kh, kw = khw if khw is not None else (pw.kh, pw.kw)
    sh, sw = (stride, stride) if isinstance(stride, int) else stride
    ph, pw_ = pad_hw if pad_hw is not None else ((pad, pad) if isinstance(pad, int) else pad)
    a.kh, a.kw, a.sh, a.sw, a.ph, a.pw = kh, kw, sh, sw, ph, pw_
    if oh is None:
        oh = (x.shape[1] + 2 * ph - kh) // sh + 1
        ow = (x.shape[2] + 2 * pw_ - kw) // sw + 1
    a.oh, a.ow = oh, ow
    a.oy_mul, a.oy_off, a.ox_mul, a.ox_off = omap
    a.cin, a.cout = pw.cin, pw.cout
    a.act, a.slope, a.res_after_act, a.out_scale = act, slope, int(res_after_act), out_scale
    a.bias = pw.bias.data_ptr() if pw.bias is not None else None
    if use_tc is None:
        use_tc = x.dtype == torch.bfloat16
    if use_tc:
        a.w = (w_override if w_override is not None else pw.tc()).data_ptr()
        check(lib().vps_conv2d_tc(C.byref(a), stream()), "conv2d_tc")
    else:
        a.w = (w_override if w_override is not None else pw.simt()).data_ptr()
        check(lib().vps_conv2d_simt(C.byref(a), stream()), "conv2d_simt")
    return y


# ------------------------------------------------------------------ FlowNet2 native ops
def correlation(f1, f2, out, pad, max_disp, stride1, stride2, act=ACT_NONE, slope=0.1):
    check(lib().vps_correlation(C.byref(vt(f1)), C.byref(vt(f2)), C.byref(vt(out)), pad, max_disp, stride1,
                                stride2, act, C.c_float(slope), stream()), "correlation")
    return out


def resample2d(src, flow, out):
    check(lib().vps_resample2d(C.byref(vt(src)), C.byref(vt(flow)), C.byref(vt(out)), stream()), "resample2d")
    return out


def channelnorm(a, out, b=None):
    bb = C.byref(vt(b)) if b is not None else None
    check(lib().vps_channelnorm(C.byref(vt(a)), bb, C.byref(vt(out)), stream()), "channelnorm")
    return out
