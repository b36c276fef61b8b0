"""ORACLE (test infrastructure, not product code) for SURVEY 8f rank 1: the step right after the FuseTrack hot path.

CPU restatement of `CityscapesVps.get_unified_pan_result` (reference tools/dataset/cityscapes_vps.py:162-226), which
turns one frame's label maps into the 3-channel (semantic, instance rank, track id) image VPQ is evaluated on.
Pinned bit-exactly against the reference's own function by tests/golden/unify_pan.npz
(tests/golden/make_unify_golden.py imports the reference and generates it; tests/test_unify_cpu.py checks this file).

Semantics kept from the reference, including its quirks:
  * instance id in `pan` = id_last_stuff + 1 + j  (j = index into cls_ind); `idx` = rank of the id among the ids PRESENT
    in the frame (ascending) -- `pan_ins` gets idx + 1 and the track id is looked up with idx, not j  (:192, :199-201);
  * the majority vote uses np.unique: ties go to the smallest class id (:196-197);
  * a region whose vote disagrees becomes stuff only if the winner has >= half of the pixels and is a stuff class (:203);
  * stuff classes (<= id_last_stuff) covering fewer than stuff_area_limit pixels of the FINAL semantic channel become 255
    (:214-219);
  * all three channels start as copies of `pan` (uint8): pixels of stuff keep their stuff id in the track-id channel, and
    values written into the uint8 maps wrap modulo 256 (:185-187, :201);
  * duplicate track ids inside a frame are re-numbered from a counter that starts at 100 and runs across frames; the
    last occurrence keeps the id (:166, :171-181)."""
import numpy as np

NUM_SEG_CLASSES, NUM_CLASSES = 19, 9            # configs/cityscapes/test_cityscapes_1gpu.yaml:7-8
ID_LAST_STUFF = NUM_SEG_CLASSES - NUM_CLASSES   # 10


def dedup_track_ids(obj_id, max_oid):
    """cityscapes_vps.py:171-181.  Returns (ids, new max_oid).  The reference patches a REVERSED copy of the array, so
    the LAST occurrence of a repeated id (in original order) keeps it and the earlier ones get max_oid, max_oid+1, ...
    walking backwards; repeated values are processed in ascending order of the value."""
    obj_id = np.asarray(obj_id).copy()
    vals, cnt = np.unique(obj_id, return_counts=True)
    out = obj_id.copy()
    for v in vals[cnt > 1]:
        pos = np.nonzero(obj_id == v)[0]
        for p in pos[-2::-1]:
            out[p] = max_oid
            max_oid += 1
    return out, max_oid


def unify_frame(seg, pan, cls_ind, obj_id, stuff_area_limit=4 * 64 * 64):
    """One frame.  seg, pan: uint8 [H,W]; cls_ind: thing class (1-based) per instance j; obj_id: de-duplicated track ids
    or None.  Returns uint8 [H,W,3]."""
    seg = np.asarray(seg)
    pan = np.asarray(pan).astype(np.uint8)
    out_seg, out_ins, out_obj = pan.copy(), pan.copy(), pan.copy()
    out_ins[pan <= ID_LAST_STUFF] = 0
    present = np.unique(pan[pan > ID_LAST_STUFF])
    for rank, pid in enumerate(present.tolist()):
        region = pan == pid
        if pid == 255:
            out_seg[region] = 255
            out_ins[region] = 0
            continue
        j = pid - ID_LAST_STUFF - 1
        hist = np.bincount(seg[region].ravel(), minlength=256)
        winner = int(np.argmax(hist))                      # first maximum = smallest class id on ties
        want = int(cls_ind[j]) + ID_LAST_STUFF
        to_stuff = winner != want and 2 * int(hist[winner]) >= int(hist.sum()) and winner <= ID_LAST_STUFF
        if to_stuff:
            out_seg[region] = winner
            out_ins[region] = 0
            out_obj[region] = 0
        else:
            out_seg[region] = np.uint8(want % 256)
            out_ins[region] = np.uint8((rank + 1) % 256)
            if obj_id is not None:
                out_obj[region] = np.uint8((int(obj_id[rank]) + 1) % 256)
    area = np.bincount(out_seg.ravel(), minlength=256)
    for c in range(ID_LAST_STUFF + 1):
        if 0 < area[c] < stuff_area_limit:
            out_seg[out_seg == c] = 255
    return np.stack([out_seg, out_ins, out_obj], axis=2)


def get_unified_pan_result(segs, pans, cls_inds, obj_ids=None, stuff_area_limit=4 * 64 * 64):
    """the reference's loop over frames (names dropped: returns a list)."""
    outs, max_oid = [], 100
    for i, (seg, pan, cls_ind) in enumerate(zip(segs, pans, cls_inds)):
        oid = None
        if obj_ids is not None and obj_ids[i] is not None:
            oid, max_oid = dedup_track_ids(obj_ids[i], max_oid)
        outs.append(unify_frame(seg, pan, cls_ind, oid, stuff_area_limit))
    return outs
