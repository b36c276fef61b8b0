"""Golden vectors for SURVEY 8f rank 1 (get_unified_pan_result): runs the REFERENCE's own function
(/root/reference/tools/dataset/cityscapes_vps.py:162-226, imported unmodified with stubs for the packages this image lacks)
on seeded synthetic label maps and stores inputs + outputs in tests/golden/unify_pan.npz.
Run in the build container only (the GPU box has no /root/reference):  python tests/golden/make_unify_golden.py"""
import collections
import collections.abc
import os
import sys
import types

import numpy as np

REF = "/root/reference"


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def import_reference():
    collections.Sequence = collections.abc.Sequence            # base_dataset.py:19 (python < 3.10 spelling)

    class EasyDict(dict):
        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError:
                raise AttributeError(k)

        def __setattr__(self, k, v):
            self[k] = v
    _stub("easydict", EasyDict=EasyDict)
    _stub("pycocotools")
    _stub("pycocotools.coco", COCO=object)
    _stub("pycocotools.mask")
    sys.path.insert(0, REF)
    from tools.config.config import config
    config.dataset.num_classes = 9                               # configs/cityscapes/test_cityscapes_1gpu.yaml:7-8
    config.dataset.num_seg_classes = 19
    from tools.dataset.cityscapes_vps import CityscapesVps
    return CityscapesVps


def synth_frame(rng, H, W, k, with_255=False):
    """semantic map, panoptic map (stuff 0..10, instance j -> 11 + j), thing class per instance (1..8), track ids"""
    seg = rng.integers(0, 11, size=((H + 31) // 32, (W + 31) // 32)).repeat(32, 0).repeat(32, 1)[:H, :W].astype(np.uint8)
    pan = seg.copy()
    cls = rng.integers(1, 9, size=k)
    for j in range(k):
        h, w = int(rng.integers(8, H // 2)), int(rng.integers(8, W // 2))
        y, x = int(rng.integers(0, H - h)), int(rng.integers(0, W - w))
        if j % 5 == 4:
            continue                                             # an instance that never wins a pixel (absent id)
        pan[y:y + h, x:x + w] = 11 + j
        mode = j % 4                                             # how the semantic head sees the region
        if mode == 0:
            seg[y:y + h, x:x + w] = 10 + cls[j]                  # agrees with the instance class
        elif mode == 1:
            seg[y:y + h, x:x + w] = rng.integers(0, 11)          # one stuff class dominates -> region becomes stuff
        elif mode == 2:
            seg[y:y + h, x:x + w] = rng.integers(0, 19, size=(h, w))     # no majority
        else:
            seg[y:y + h // 2, x:x + w] = 10 + (cls[j] % 8) + 1   # a different thing class wins the vote
    if with_255:
        pan[:4, :7] = 255
    obj = rng.integers(0, 40, size=k)
    obj[k // 2] = obj[0]                                         # redundant track ids (cityscapes_vps.py:171-181)
    if k > 6:
        obj[k - 1] = obj[1]
        obj[3] = obj[1]
    return seg, pan, cls.astype(np.int64), obj.astype(np.int64)


def main():
    Cvps = import_reference()
    rng = np.random.default_rng(2024)
    H, W = 256, 512
    ks = [12, 7, 1, 20, 0]
    segs, pans, clss, objs = [], [], [], []
    for i, k in enumerate(ks):
        if k == 0:                                                # the dummy detection (mask_roi.py:136-142): class 0
            seg = rng.integers(0, 11, size=(H // 32, W // 32)).repeat(32, 0).repeat(32, 1).astype(np.uint8)
            pan = seg.copy()
            pan[40:60, 100:180] = 11                              # pixels where the all-zero instance channel won
            cls, obj = np.zeros(1, np.int64), np.zeros(1, np.int64)
        else:
            seg, pan, cls, obj = synth_frame(rng, H, W, k, with_255=(i == 1))
        segs.append(seg); pans.append(pan); clss.append(cls); objs.append(obj)
    names = ["f%d" % i for i in range(len(ks))]
    out = Cvps.get_unified_pan_result(None, [s.copy() for s in segs], [p.copy() for p in pans], [c.copy() for c in clss],
                                      obj_ids=[o.copy() for o in objs], names=names)
    out_noid = Cvps.get_unified_pan_result(None, [s.copy() for s in segs], [p.copy() for p in pans], [c.copy() for c in clss],
                                           obj_ids=None, names=names)
    d = {}
    for i, n in enumerate(names):
        d["seg%d" % i], d["pan%d" % i], d["cls%d" % i], d["obj%d" % i] = segs[i], pans[i], clss[i], objs[i]
        d["out%d" % i], d["out_noid%d" % i] = out[n], out_noid[n]
    d["nframes"] = np.int64(len(ks))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "unify_pan.npz")
    np.savez_compressed(path, **d)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
