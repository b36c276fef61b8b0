"""2-channel -> PNG / JSON writer: `converter_2ch_track_core` + the file layout of
`inference_panoptic_video` (reference tools/dataset/cityscapes_vps.py:26-160), frame by frame.

The device part is two tested ops -- `vps_pan2ch_ids` (segment key per pixel) and the sort + run-length table of
`vps_tube_confusion` (segment areas), both through `vps_b200.vpq.segments_from_pan2ch`; bounding boxes, PNG encoding and the
JSON are host work (they end in files).  Segment ids: the reference uses the colours of panopticapi's `IdGenerator` (one fixed
colour per stuff category, random per thing key; everything downstream is invariant to the values); here id =
1000 * semantic + 1 for stuff, 1000 * semantic + track + 1 for things, colour = id2rgb(id).  One reference quirk is kept:
the bbox of a stuff segment that merges several keys is the bbox of its LAST key (segm_info[colour] is overwritten per key,
cityscapes_vps.py:131-138) -- `add_frame_ids` takes it from `bbox_ids` when given."""
import json
import os

import numpy as np


def id2rgb(ids):
    ids = np.asarray(ids).astype(np.uint32)
    return np.stack([ids % 256, (ids // 256) % 256, ids // 65536], axis=-1).astype(np.uint8)


def _clean_name(name):
    # inference_panoptic_video.save_image (:69)
    return name.replace('_leftImg8bit', '').replace('_newImg8bit', '').replace('jpg', 'png').replace('jpeg', 'png')


def last_key_ids(pan_2ch, num_stuff=11):
    """id map in which a stuff segment keeps only the pixels of its LARGEST key 1000 * semantic + track channel: the reference
    overwrites segm_info[colour] for every key of a category in ascending key order (cityscapes_vps.py:112-138), so the bbox
    that survives is the last key's, while the area is re-counted from the merged PNG."""
    p = np.asarray(pan_2ch).astype(np.uint32)
    sem, trk = p[..., 0], p[..., 2]
    ids = np.where(sem == 255, 0, 1000 * sem + np.where(sem < num_stuff, 0, trk) + 1).astype(np.uint32)
    out = ids.copy()
    for c in np.unique(sem[sem < num_stuff]).tolist():
        m = sem == c
        out[m & (trk != trk[m].max())] = 0
    return out


class PanWriter:
    """Feed the unified 3-channel results of a clip in order; sampled frames ([(labeled_fid // lambda_)::lambda_], :35) are
    converted and written to <output_dir>/pan_2ch/ and <output_dir>/pan_pred/; `finish()` writes pred.json."""

    def __init__(self, output_dir=None, labeled_fid=20, lambda_=5, sample=True):
        self.output_dir = output_dir
        self.start, self.step = (labeled_fid // lambda_, lambda_) if sample else (0, 1)
        self.index = 0
        self.annotations, self.names = [], []

    def _sampled(self):
        i = self.index
        self.index += 1
        return i >= self.start and (i - self.start) % self.step == 0

    def add_frame(self, name, pan_2ch):
        """pan_2ch: uint8 CUDA tensor [H,W,3] (vps_b200.postproc.PanUnifier).  Returns the frame's annotation or None if the
        frame is not a sampled one."""
        if not self._sampled():
            return None
        from .vpq import segments_from_pan2ch
        ids, segs = segments_from_pan2ch(pan_2ch)
        p2 = pan_2ch.cpu().numpy()
        return self.add_frame_ids(name, ids.cpu().numpy().astype(np.uint32), segs, p2, _counted=True, bbox_ids=last_key_ids(p2))

    def add_frame_ids(self, name, ids, segs, pan_2ch=None, _counted=False, bbox_ids=None):
        """host part: ids [H,W] uint32 (0 = VOID), segs = [{id, category_id, iscrowd, area}] (any order); bbox_ids: id map
        restricted to the pixels the reference takes a segment's bbox from (`last_key_ids`), default = ids"""
        if not _counted and not self._sampled():
            return None
        from scipy import ndimage
        ids = np.asarray(ids)
        if bbox_ids is None and pan_2ch is not None:
            bbox_ids = last_key_ids(pan_2ch)
        box_src = ids if bbox_ids is None else np.asarray(bbox_ids)
        uniq = sorted(s["id"] for s in segs)
        # dense relabelling so that find_objects does not scan 19000 empty labels
        lut = np.zeros(int(ids.max()) + 1, dtype=np.int32)
        lut[np.asarray(uniq, dtype=np.int64)] = np.arange(1, len(uniq) + 1, dtype=np.int32)
        boxes = ndimage.find_objects(lut[box_src])
        by_id = {s["id"]: s for s in segs}
        info = []
        for rank, i in enumerate(uniq):
            sl = boxes[rank]
            y, x = sl[0].start, sl[1].start
            s = by_id[i]
            info.append({"category_id": int(s["category_id"]), "iscrowd": 0, "id": int(i),
                         "bbox": [int(x), int(y), int(sl[1].stop - 1 - x), int(sl[0].stop - 1 - y)], "area": int(s["area"])})
        ann = {"segments_info": info}
        self.annotations.append(ann)
        self.names.append(name)
        if self.output_dir is not None:
            from PIL import Image
            fn = _clean_name(name)
            for sub, img in (("pan_pred", id2rgb(ids)), ("pan_2ch", pan_2ch)):
                if img is None:
                    continue
                path = os.path.join(self.output_dir, sub, fn)
                os.makedirs(os.path.dirname(path), exist_ok=True)
                Image.fromarray(np.ascontiguousarray(img)).save(path)
        return ann

    def finish(self):
        pred_json = {"annotations": self.annotations}
        if self.output_dir is not None:
            os.makedirs(self.output_dir, exist_ok=True)
            with open(os.path.join(self.output_dir, "pred.json"), "w") as f:
                json.dump(pred_json, f)
        return pred_json
