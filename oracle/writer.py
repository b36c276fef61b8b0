"""ORACLE (test infrastructure, not product code) for SURVEY 8f rank 1b: the 2-channel -> PNG / JSON writer.

CPU restatement of `CityscapesVps.converter_2ch_track_core` and the file layout of `inference_panoptic_video`
(reference tools/dataset/cityscapes_vps.py:26-160).  The reference colours every segment with panopticapi's
`IdGenerator.get_color` (panopticapi is a third-party dependency that is NOT vendored in /root/reference; its generator draws
random colours), so segment ids are not reproducible even by the reference itself; everything VPQ consumes is invariant to
them.  What IS determined: `get_color` returns the category's one fixed colour for a stuff category and a fresh colour per
request for a thing category, and the colour is the segment id (:131-138).  This restatement therefore uses the deterministic
ids 1000 * semantic + 1 (stuff) / 1000 * semantic + track + 1 (thing), 0 = VOID, and is pinned against the reference's own
function run with a stand-in generator of exactly that behaviour, modulo a bijection of ids, on frames produced by the
reference's own get_unified_pan_result (tests/golden/make_writer_golden.py, tests/test_writer_cpu.py).

Kept from the reference: keys 1000 * semantic + track channel are walked in ascending order (:104-112); VOID (semantic 255)
pixels stay 0 (:113-114); every key of a stuff category maps to the same segment -- native stuff pixels carry their pan value
in the track channel, thing regions demoted to stuff carry 0 (get_unified_pan_result :185-207) -- whose dict entry is
overwritten per key, so its bbox is the LAST key's while its area is re-counted from the merged image (:131-150); a thing
keeps its id across the frames of a clip (:116-121); bbox = [x_min, y_min, x_max - x_min, y_max - y_min] (no +1), iscrowd = 0;
frames are sampled [(labeled_fid // lambda)::lambda] before writing (:35)."""
import numpy as np


def id2rgb(ids):
    ids = np.asarray(ids).astype(np.uint32)
    return np.stack([ids % 256, (ids // 256) % 256, ids // 65536], axis=-1).astype(np.uint8)


def rgb2id(rgb):
    rgb = np.asarray(rgb).astype(np.uint32)
    return rgb[..., 0] + 256 * rgb[..., 1] + 65536 * rgb[..., 2]


def convert_frame(pan_2ch, num_stuff=11):
    """one frame of converter_2ch_track_core: returns (segments_info list, id map uint32 [H,W])"""
    p = np.asarray(pan_2ch).astype(np.uint32)
    sem, trk = p[..., 0], p[..., 2]
    key = 1000 * sem + trk
    ids = np.where(sem == 255, 0, 1000 * sem + np.where(sem < num_stuff, 0, trk) + 1).astype(np.uint32)
    info = {}
    for k in np.unique(key).tolist():                       # ascending keys; a later key of the same segment overwrites
        if k // 1000 == 255:
            continue
        m = key == k
        i = int(ids[m][0])
        ys, xs = np.nonzero(m)
        x, y = int(xs.min()), int(ys.min())
        info[i] = {"category_id": int(k // 1000), "iscrowd": 0, "id": i,
                   "bbox": [x, y, int(xs.max()) - x, int(ys.max()) - y], "area": 0}
    for i, a in zip(*np.unique(ids, return_counts=True)):   # areas re-counted from the merged id image
        if int(i) != 0:
            info[int(i)]["area"] = int(a)
    return list(info.values()), ids


def sample_frames(items, labeled_fid=20, lambda_=5):
    return items[(labeled_fid // lambda_)::lambda_]
