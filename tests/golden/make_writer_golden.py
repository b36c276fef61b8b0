"""Golden vectors for SURVEY 8f rank 1b (2-channel -> PNG/JSON converter): runs the REFERENCE's own
`converter_2ch_track_core` (/root/reference/tools/dataset/cityscapes_vps.py:96-158, imported unmodified) on seeded synthetic
semantic / panoptic maps chained through the reference's own `get_unified_pan_result` (:162-226).  panopticapi (the colour
generator) is not vendored in the reference: a stand-in `IdGenerator` with panopticapi's contract (the category's fixed
colour for a stuff category, a fresh colour per request for a thing category) makes the run reproducible; ids are therefore
compared modulo a bijection.  Output: tests/golden/writer_frames.npz + .json.
Run in the build container only:  python tests/golden/make_writer_golden.py"""
import json
import os
import sys
import types

import numpy as np

from make_unify_golden import import_reference


class StandInIdGenerator:
    """same interface as panopticapi.utils.IdGenerator.get_color: stuff -> the category's own colour, thing -> a new colour"""

    def __init__(self, categories):
        self.categories = categories
        self.n = 0

    def get_color(self, cat_id):
        if self.categories[int(cat_id)]["isthing"] == 0:
            return [int(cat_id) + 1, 0, 1]
        self.n += 1
        return [self.n % 256, self.n // 256, 2]


def rgb2id(color):
    c = np.asarray(color).astype(np.uint32)
    return int(c[0] + 256 * c[1] + 65536 * c[2])


def synth_clip(Cvps, rng, nfr, H, W):
    """frames as the path produces them: seeded semantic / panoptic maps through the REFERENCE's own get_unified_pan_result
    (so stuff pixels carry their pan value in the track channel, demoted thing regions carry 0, small stuff becomes VOID
    with a leftover track value, duplicate track ids are re-numbered), plus a VOID band."""
    from make_unify_golden import synth_frame
    segs, pans, clss, objs = [], [], [], []
    for f in range(nfr):
        seg, pan, cls, obj = synth_frame(rng, H, W, 9, with_255=(f == 2))
        obj = (obj % 5) + 1 if f % 2 else obj            # some track ids recur across frames
        segs.append(seg); pans.append(pan); clss.append(cls); objs.append(obj)
    names = ["f%d" % i for i in range(nfr)]
    out = Cvps.get_unified_pan_result(None, segs, pans, clss, obj_ids=objs, stuff_area_limit=300, names=names)
    return [out[n] for n in names]


def main():
    utils = types.ModuleType("panopticapi.utils")
    utils.IdGenerator, utils.rgb2id = StandInIdGenerator, rgb2id
    sys.modules["panopticapi"] = types.ModuleType("panopticapi")
    sys.modules["panopticapi.utils"] = utils
    Cvps = import_reference()
    categories = {i: {"id": i, "isthing": 1 if i >= 11 else 0} for i in range(19)}
    rng = np.random.default_rng(99)
    frames = synth_clip(Cvps, rng, 6, 96, 160)
    ann, pans = Cvps.converter_2ch_track_core(None, 0, frames, StandInIdGenerator(categories))
    out = {"nframes": np.int64(len(frames))}
    for i, (fr, pf) in enumerate(zip(frames, pans)):
        out["in%d" % i], out["png%d" % i] = fr, pf
    here = os.path.dirname(os.path.abspath(__file__))
    np.savez_compressed(os.path.join(here, "writer_frames.npz"), **out)
    json.dump(ann, open(os.path.join(here, "writer_frames.json"), "w"))
    print("wrote writer_frames.npz/.json", len(ann), "frames")


if __name__ == "__main__":
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    main()
