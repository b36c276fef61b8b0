"""Golden vectors for SURVEY 8f rank 1b (2-channel -> PNG/JSON converter): runs the REFERENCE's own
`converter_2ch_track_core` (/root/reference/tools/dataset/cityscapes_vps.py:96-158, imported unmodified) on seeded synthetic
3-channel frames.  panopticapi (the colour generator) is not vendored in the reference: a stand-in `IdGenerator` hands out
distinct colours (fixed colour per stuff category, fresh colour per thing request) so the run is reproducible; ids are
therefore compared modulo a bijection.  Output: tests/golden/writer_frames.npz + .json.
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


def synth_clip(rng, nfr, H, W):
    frames = []
    base = rng.integers(0, 11, size=((H + 15) // 16, (W + 15) // 16)).repeat(16, 0).repeat(16, 1)[:H, :W]
    inst = [(int(rng.integers(11, 19)), int(rng.integers(6, H // 2)), int(rng.integers(6, W // 2)), int(rng.integers(0, H // 2)),
             int(rng.integers(0, W // 2)), int(rng.integers(1, 200))) for _ in range(7)]
    for f in range(nfr):
        p = np.zeros((H, W, 3), np.uint8)
        p[..., 0] = base
        p[:5, :, 0] = 255                                              # VOID band
        for rank, (c, h, w, y, x, trk) in enumerate(inst):
            if (f + rank) % 4 == 3:
                continue                                               # instance absent in this frame
            yy, xx = min(y + f, H - h), min(x + 2 * f, W - w)
            p[yy:yy + h, xx:xx + w, 0] = c
            p[yy:yy + h, xx:xx + w, 1] = rank + 1
            p[yy:yy + h, xx:xx + w, 2] = trk
        frames.append(p)
    return frames


def main():
    utils = types.ModuleType("panopticapi.utils")
    utils.IdGenerator, utils.rgb2id = StandInIdGenerator, rgb2id
    sys.modules["panopticapi"] = types.ModuleType("panopticapi")
    sys.modules["panopticapi.utils"] = utils
    Cvps = import_reference()
    categories = {i: {"id": i, "isthing": 1 if i >= 11 else 0} for i in range(19)}
    rng = np.random.default_rng(99)
    frames = synth_clip(rng, 6, 64, 96)
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
