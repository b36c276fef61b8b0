"""ORACLE (test infrastructure, not product code) for SURVEY 8f rank 1b: the 2-channel -> PNG / JSON writer.

CPU restatement of `CityscapesVps.converter_2ch_track_core` and the file layout of `inference_panoptic_video`
(reference tools/dataset/cityscapes_vps.py:26-160).  The reference colours every segment with panopticapi's
`IdGenerator.get_color` (panopticapi is a third-party dependency that is NOT vendored in /root/reference; its generator draws
random colours), so segment ids are not reproducible even by the reference itself; everything VPQ consumes is invariant to
them.  This restatement therefore uses the deterministic id  1000 * semantic + track + 1  (0 = VOID) and is pinned against
the reference's own function run with a stand-in generator, modulo a bijection of ids
(tests/golden/make_writer_golden.py, tests/test_writer_cpu.py).

Kept from the reference: segments are keyed by 1000 * semantic + track channel (:104), VOID (semantic 255) pixels stay 0 (:110-
111), a thing keeps its id across the frames of a clip (:116-121), bbox = [x_min, y_min, x_max - x_min, y_max - y_min]
(:131-137, no +1), area = pixel count, iscrowd = 0, frames are sampled [(labeled_fid // lambda)::lambda] before writing (:35)."""
import numpy as np


def id2rgb(ids):
    ids = np.asarray(ids).astype(np.uint32)
    return np.stack([ids % 256, (ids // 256) % 256, ids // 65536], axis=-1).astype(np.uint8)


def rgb2id(rgb):
    rgb = np.asarray(rgb).astype(np.uint32)
    return rgb[..., 0] + 256 * rgb[..., 1] + 65536 * rgb[..., 2]


def convert_frame(pan_2ch):
    """one frame of converter_2ch_track_core: returns (segments_info list, id map uint32 [H,W])"""
    p = np.asarray(pan_2ch).astype(np.uint32)
    key = 1000 * p[..., 0] + p[..., 2]
    ids = np.where(p[..., 0] == 255, 0, key + 1).astype(np.uint32)
    segs = []
    for i in np.unique(ids).tolist():
        if i == 0:
            continue
        ys, xs = np.nonzero(ids == i)
        x, y = int(xs.min()), int(ys.min())
        segs.append({"category_id": int((i - 1) // 1000), "iscrowd": 0, "id": int(i),
                     "bbox": [x, y, int(xs.max()) - x, int(ys.max()) - y], "area": int(ys.size)})
    return segs, ids


def sample_frames(items, labeled_fid=20, lambda_=5):
    return items[(labeled_fid // lambda_)::lambda_]
