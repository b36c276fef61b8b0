"""Import the reference's own PYTHON modules (mcahny/vps under /root/reference) on this CPU-only,
mmcv-less machine, so that golden vectors come from the reference's code rather than from our restatement.

What is stubbed (nothing of the reference is copied or modified):
  * mmcv (absent, pinned 0.2.14): a permissive stub; the initialisers come from the reference's vendored copy
    mmdet/models/utils/weight_init.py; is_str etc. are trivial.
  * pycocotools / matplotlib / pdb-free helpers: permissive stubs (not on the inference arithmetic path).
  * the compiled extensions (cpython-37/CUDA-10 binaries, unusable here): mmdet.ops.{nms, RoIAlign, DeformConv},
    correlation/resample2d/channelnorm packages, UPSNet gpu_nms  ->  oracle.ops restatements of the .cu sources.
  * tools.config.config (needs easydict): the two constants MaskROI reads (config.py:47,169).
  * .cuda() / torch.cuda.current_device(): identity / 'cpu' (the reference hard-codes CUDA placement).
  * torch-1.4 API drift: F.grid_sample / F.interpolate default align_corners (already False in torch 2),
    torch.addcmul(t, value, a, b) legacy signature, legacy autograd.Function (bypassed by the op stubs).
"""
import importlib
import importlib.util
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

REF = os.environ.get("VPS_REFERENCE", "/root/reference")


class _Dummy(object):
    """class usable as base class, decorator, callable, attribute bag"""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        if len(a) == 1 and callable(a[0]) and not k:
            return a[0]
        return _Dummy()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Dummy()


class _MagicModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        full = self.__name__ + "." + name
        if full in sys.modules:
            return sys.modules[full]
        return type(name, (_Dummy,), {})


def _stub(name, **attrs):
    m = _MagicModule(name)
    m.__path__ = []
    m.__dict__.update(attrs)
    sys.modules[name] = m
    parent, _, child = name.rpartition(".")
    if parent and parent in sys.modules:
        setattr(sys.modules[parent], child, m)
    return m


def _load_file(modname, path):
    spec = importlib.util.spec_from_file_location(modname, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


_done = False


def setup():
    """Install stubs and return the imported reference package `mmdet.models`."""
    global _done
    if _done:
        return importlib.import_module("mmdet.models")
    if not os.path.isdir(os.path.join(REF, "mmdet")):
        raise RuntimeError("reference tree not found at %s" % REF)
    root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    if root not in sys.path:
        sys.path.insert(0, root)
    from oracle import ops as O

    # ---- device shims
    torch.Tensor.cuda = lambda self, *a, **k: self
    nn.Module.cuda = lambda self, *a, **k: self
    torch.cuda.current_device = lambda: "cpu"
    torch.Tensor.pin_memory = lambda self, *a, **k: self
    _orig_get_device = torch.Tensor.get_device
    torch.Tensor.get_device = lambda self: "cpu"
    _orig_to = torch.Tensor.to

    def _cpu_dev(d):
        if isinstance(d, str) and d.startswith("cuda"):
            return "cpu"
        if isinstance(d, torch.device) and d.type == "cuda":
            return torch.device("cpu")
        if isinstance(d, int) and not isinstance(d, bool):
            return "cpu"
        return d

    def _to(self, *a, **k):
        k.pop("non_blocking", None)
        a = tuple(_cpu_dev(x) if isinstance(x, (str, torch.device)) else x for x in a)
        if "device" in k:
            k["device"] = _cpu_dev(k["device"])
        return _orig_to(self, *a, **k)
    torch.Tensor.to = _to
    for _name in ("arange", "zeros", "ones", "full", "empty", "tensor", "linspace", "zeros_like", "ones_like", "randn", "rand"):
        _f = getattr(torch, _name)

        def _mk(f):
            def w(*a, **k):
                if "device" in k:
                    k["device"] = _cpu_dev(k["device"])
                return f(*a, **k)
            return w
        setattr(torch, _name, _mk(_f))
    _addcmul = torch.addcmul

    def addcmul(inp, *args, **kw):
        if len(args) == 3 and not torch.is_tensor(args[0]):      # legacy (input, value, t1, t2)
            return _addcmul(inp, args[1], args[2], value=args[0])
        return _addcmul(inp, *args, **kw)
    torch.addcmul = addcmul

    # ---- mmcv
    # initialisers: values are irrelevant (every parameter is overwritten by load_state_dict), and the
    # reference's vendored copies crash on bias=None under torch 2, so they are no-ops here.
    _noop = lambda *a, **k: None
    inits = dict(xavier_init=_noop, normal_init=_noop, kaiming_init=_noop, uniform_init=_noop, constant_init=_noop,
                 bias_init_with_prob=lambda p: 0.0)
    mmcv = _stub("mmcv", is_str=lambda x: isinstance(x, str))
    _stub("mmcv.cnn", **inits)
    _stub("mmcv.cnn.weight_init", **inits)
    _stub("mmcv.runner"); _stub("mmcv.parallel"); _stub("mmcv.utils"); _stub("mmcv.image"); _stub("mmcv.visualization")
    _stub("pycocotools"); _stub("pycocotools.mask"); _stub("pycocotools.coco"); _stub("pycocotools.cocoeval")
    _stub("matplotlib"); _stub("matplotlib.pyplot"); _stub("terminaltables"); _stub("imagecorruptions")

    # ---- UPSNet global config (tools/config/config.py:47,169)
    class _NS(object):
        pass
    config = _NS(); config.network = _NS(); config.test = _NS(); config.dataset = _NS()
    config.network.bbox_reg_weights = (10.0, 10.0, 5.0, 5.0)
    config.test.max_det = 100
    _stub("tools"); _stub("tools.config"); _stub("tools.config.config", config=config)

    # ---- mmdet package skeleton (real sub-packages, stubbed compiled ops)
    mmdet = types.ModuleType("mmdet"); mmdet.__path__ = [os.path.join(REF, "mmdet")]
    sys.modules["mmdet"] = mmdet

    class RoIAlign(nn.Module):
        def __init__(self, out_size, spatial_scale, sample_num=0, use_torchvision=False):
            super().__init__()
            self.out_size = (out_size, out_size) if isinstance(out_size, int) else out_size
            self.spatial_scale, self.sample_num = float(spatial_scale), int(sample_num)

        def forward(self, features, rois):
            return O.roi_align(features, rois, self.out_size[0], self.spatial_scale, self.sample_num)

    class DeformConv(nn.Module):
        """mmdet/ops/dcn/deform_conv.py:190-236 parameterisation; forward = oracle restatement of the .cu"""

        def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                     deformable_groups=1, bias=False):
            super().__init__()
            import math
            k = kernel_size if isinstance(kernel_size, int) else kernel_size[0]
            self.stride, self.padding, self.dilation = stride, padding, dilation
            self.weight = nn.Parameter(torch.Tensor(out_channels, in_channels // groups, k, k))
            stdv = 1. / math.sqrt(in_channels * k * k)
            self.weight.data.uniform_(-stdv, stdv)

        def forward(self, x, offset):
            return O.deform_conv(x, offset, self.weight, self.stride, self.padding, self.dilation)

    _nms_fn = lambda dets, thr, device_id=None: O.nms(dets, thr)
    ops_stub = _stub("mmdet.ops", nms=_nms_fn, RoIAlign=RoIAlign,
                     DeformConv=DeformConv)
    mmdet.ops = ops_stub
    _nw = _stub("mmdet.ops.nms.nms_wrapper", nms=ops_stub.nms) if False else None
    _stub("mmdet.ops.nms", nms=ops_stub.nms)
    _stub("mmdet.ops.nms.nms_wrapper", nms=ops_stub.nms)
    _stub("mmdet.ops.dcn"); _stub("mmdet.ops.roi_align")
    ops_stub.nms = _nms_fn      # `from mmdet.ops import nms` must yield the function, not the sub-package

    class _Corr(nn.Module):
        def __init__(self, pad_size=3, kernel_size=3, max_displacement=20, stride1=1, stride2=2, corr_multiply=1):
            super().__init__()
            self.a = (pad_size, kernel_size, max_displacement, stride1, stride2, corr_multiply)

        def forward(self, x1, x2):
            return O.correlation(x1, x2, *self.a)

    class _Resample(nn.Module):
        def __init__(self, kernel_size=1, bilinear=True):
            super().__init__()

        def forward(self, a, b):
            return O.resample2d(a, b)

    class _CN(nn.Module):
        def __init__(self, norm_deg=2):
            super().__init__()

        def forward(self, x):
            return O.channelnorm(x)

    fm = "mmdet.models.flow_modules."
    for pkg, mod, attrs in ((fm + "correlation_package", "correlation", dict(Correlation=_Corr)),
                            (fm + "resample2d_package", "resample2d", dict(Resample2d=_Resample)),
                            (fm + "channelnorm_package", "channelnorm", dict(ChannelNorm=_CN))):
        sys.modules.setdefault(pkg, types.ModuleType(pkg))
        sys.modules[pkg].__path__ = []
        m = types.ModuleType(pkg + "." + mod)
        m.__dict__.update(attrs)
        sys.modules[pkg + "." + mod] = m
    # UPSNet nms / bbox cython modules
    nmsw = types.ModuleType("mmdet.models.utils.upsnet.nms.nms")
    nmsw.gpu_nms_wrapper = lambda thresh, device_id: (lambda dets: O.gpu_nms_upsnet(dets, thresh))
    nmsw.py_nms_wrapper = nmsw.cpu_nms_wrapper = lambda thresh: (lambda dets: O.gpu_nms_upsnet(dets, thresh))
    for name in ("mmdet.models.utils.upsnet", "mmdet.models.utils.upsnet.nms", "mmdet.models.utils.upsnet.bbox"):
        pm = types.ModuleType(name)
        pm.__path__ = [os.path.join(REF, name.replace(".", "/"))]
        sys.modules[name] = pm
    sys.modules["mmdet.models.utils.upsnet.nms.nms"] = nmsw
    _stub("mmdet.models.utils.upsnet.bbox.bbox")
    # dataset helpers imported by bfp_tcea.py for visualisation only
    _stub("mmdet.datasets"); _stub("mmdet.datasets.pipelines"); _stub("mmdet.datasets.pipelines.flow_utils")
    _done = True
    return importlib.import_module("mmdet.models")
