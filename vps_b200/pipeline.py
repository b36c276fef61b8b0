"""GPU input stage: what the reference's test pipeline does on the host for `img` and `ref_img` between
LoadImageFromFile / LoadRefImageFromFile and the model (configs/cityscapes/fusetrack.py:172-190, mmdet/datasets/pipelines/
{loading,transforms,formating}.py):  Resize(img_scale=(2048,1024), keep_ratio) -> Normalize(mean, std, to_rgb) -> Pad(32) ->
ImageToTensor, plus the `img_meta` fields `simple_test` reads.

`InputStage` takes the decoded uint8 HWC BGR frame (what cv2.imread / mmcv.imread return), uploads it as uint8 (12.6 MB per
1024x2048 pair instead of 50 MB of fp32) and normalises / pads / transposes it on the device in one pass (`vps_preprocess_u8`,
bit-identical to mmcv.imnormalize's float32 arithmetic).  Resize: the rescale factor mmcv.imrescale derives is
min(long_edge / max(h, w), short_edge / min(h, w)); for Cityscapes frames (1024x2048) it is exactly 1 and cv2.resize is the
identity, which is the only case handled here -- other sizes raise (cv2's fixed-point INTER_LINEAR is not restated)."""
import ctypes as C

import numpy as np
import torch

from . import ops
from ._lib import lib

CITYSCAPES_NORM = dict(mean=[123.675, 116.28, 103.53], std=[58.395, 57.12, 57.375], to_rgb=True)    # fusetrack.py:153-154


class InputStage:
    def __init__(self, mean=CITYSCAPES_NORM["mean"], std=CITYSCAPES_NORM["std"], to_rgb=True, img_scale=(2048, 1024), size_divisor=32,
                 device="cuda:0"):
        self.mean = (C.c_float * 3)(*[float(np.float32(v)) for v in mean])
        self.std = (C.c_float * 3)(*[float(np.float32(v)) for v in std])
        self.to_rgb, self.img_scale, self.div = bool(to_rgb), img_scale, int(size_divisor)
        self.dev = torch.device(device)

    def scale_factor(self, h, w):
        """mmcv.imrescale(img, scale=(long, short)): min(long / max(h, w), short / min(h, w))"""
        long_e, short_e = max(self.img_scale), min(self.img_scale)
        return min(long_e / max(h, w), short_e / min(h, w))

    def __call__(self, img_u8, out=None, stream_tensor=None):
        """img_u8: uint8 [H,W,3] BGR, host (pinned or not) or CUDA tensor.  Returns (fp32 CUDA [1,3,Hp,Wp], meta fields)."""
        assert img_u8.dtype == torch.uint8 and img_u8.dim() == 3 and img_u8.shape[2] == 3
        h, w = int(img_u8.shape[0]), int(img_u8.shape[1])
        sf = self.scale_factor(h, w)
        if abs(sf - 1.0) > 1e-12:
            raise NotImplementedError("InputStage: Resize with scale %.4f (frame %dx%d): only the identity case of the Cityscapes "
                                      "pipeline is implemented on the device" % (sf, h, w))
        hp, wp = (h + self.div - 1) // self.div * self.div, (w + self.div - 1) // self.div * self.div
        d = img_u8 if img_u8.is_cuda else img_u8.to(self.dev, non_blocking=True)
        d = d.contiguous()
        if out is None:
            out = torch.empty(1, 3, hp, wp, dtype=torch.float32, device=d.device)
        ops.check(lib().vps_preprocess_u8(ops._ptr(d), h, w, self.mean, self.std, int(self.to_rgb), ops._ptr(out), hp, wp, ops.stream()),
                  "preprocess_u8")
        d.record_stream(torch.cuda.current_stream(d.device))
        meta = dict(img_shape=(h, w, 3), ori_shape=(h, w, 3), pad_shape=(hp, wp, 3), scale_factor=1.0,
                    img_norm_cfg=dict(mean=np.array(list(self.mean), np.float32), std=np.array(list(self.std), np.float32), to_rgb=self.to_rgb))
        return out, meta
