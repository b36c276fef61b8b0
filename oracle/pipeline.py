"""ORACLE (test infrastructure): CPU restatement of the reference's host-side input pipeline for one frame --
mmcv 0.2.14 `imnormalize` (float32(img); BGR->RGB; (img - mean) / std, called from transforms.py:310-315), `impad_to_multiple`
(zero pad bottom / right, transforms.py:258-266) and ImageToTensor (HWC -> CHW, formating.py:46-68).  mmcv is not vendored in
/root/reference (pinned ==0.2.14, requirements.txt:1): the three functions are restated from its published source."""
import numpy as np


def imnormalize(img, mean, std, to_rgb=True):
    img = img.astype(np.float32)
    if to_rgb:
        img = img[..., ::-1]
    return (img - np.asarray(mean, np.float32)) / np.asarray(std, np.float32)


def impad_to_multiple(img, divisor, pad_val=0):
    h, w = img.shape[:2]
    hp, wp = int(np.ceil(h / divisor)) * divisor, int(np.ceil(w / divisor)) * divisor
    out = np.full((hp, wp, img.shape[2]), pad_val, dtype=img.dtype)
    out[:h, :w] = img
    return out


def prepare_frame(img_u8_bgr, mean, std, to_rgb=True, divisor=32):
    x = impad_to_multiple(imnormalize(img_u8_bgr, mean, std, to_rgb), divisor)
    return np.ascontiguousarray(x.transpose(2, 0, 1))[None]
