"""ORACLE (test infrastructure, not product code) for SURVEY 8f rank 2: the VPQ evaluator core.

CPU restatement of `vpq_compute_single_core` and `PQStat.pq_average` (reference tools/eval_vpq.py:73-203, 20-70): tube
matching of predicted against ground-truth video-panoptic segments over windows of `nframes` consecutive (sampled) frames.
Pinned against the reference's own functions by tests/golden/vpq_tubes.npz (tests/golden/make_vpq_golden.py imports
tools/eval_vpq.py and generates it; tests/test_vpq_cpu.py checks this file).

A frame is (gt_segments, pred_segments, gt_ids, pred_ids): segment lists of dicts {id, category_id, iscrowd, area} and the
[H,W] integer id maps (the reference decodes them from RGB PNGs as r + 256 g + 65536 b, eval_vpq.py:87-89).
Semantics kept: predicted areas are recounted from the id map (:104-113), ground-truth areas come from the annotation,
tube areas are sums over the window (:123-136), the confusion is np.unique over gt * 2^24 + pred (:139-145), a match needs
equal category, non-crowd gt and IoU > 0.5 with VOID (id 0) pixels removed from the union (:164-176); unmatched
predictions are ignored when more than half of their area lies on VOID or the crowd region of their category (:188-201)."""
import copy
from collections import defaultdict

import numpy as np

OFFSET = 256 * 256 * 256
VOID = 0


class CatStat:
    __slots__ = ("iou", "tp", "fp", "fn")

    def __init__(self):
        self.iou, self.tp, self.fp, self.fn = 0.0, 0, 0, 0


def merge_segments(seg_list):
    out = {}
    for el in seg_list:
        if el["id"] in out:
            out[el["id"]]["area"] += el["area"]
        else:
            out[el["id"]] = copy.deepcopy(el)
    return out


def tube_confusion(gt_ids, pred_ids):
    """np.unique over the packed (gt, pred) pairs of a tube: returns (pairs uint64 sorted, counts)."""
    packed = np.asarray(gt_ids).astype(np.uint64) * np.uint64(OFFSET) + np.asarray(pred_ids).astype(np.uint64)
    return np.unique(packed, return_counts=True)


def match_tube(stat, vid_gt, vid_pred, pairs, counts):
    """eval_vpq.py:147-201 on one tube, given its confusion; updates `stat` (dict category -> CatStat) in place."""
    conf = {}
    for lab, inter in zip(pairs.tolist(), counts.tolist()):
        conf[(lab // OFFSET, lab % OFFSET)] = inter
    gt_matched, pred_matched = set(), set()
    for (g, p), inter in conf.items():
        if g not in vid_gt or p not in vid_pred:
            continue
        if vid_gt[g]["iscrowd"] == 1 or vid_gt[g]["category_id"] != vid_pred[p]["category_id"]:
            continue
        union = vid_pred[p]["area"] + vid_gt[g]["area"] - inter - conf.get((VOID, p), 0)
        iou = inter / union
        assert iou <= 1.0
        if iou > 0.5:
            c = stat[vid_gt[g]["category_id"]]
            c.tp += 1
            c.iou += iou
            gt_matched.add(g)
            pred_matched.add(p)
    crowd = {}
    for g, info in vid_gt.items():
        if g in gt_matched:
            continue
        if info["iscrowd"] == 1:
            crowd[info["category_id"]] = g
            continue
        stat[info["category_id"]].fn += 1
    for p, info in vid_pred.items():
        if p in pred_matched:
            continue
        inter = conf.get((VOID, p), 0)
        if info["category_id"] in crowd:
            inter += conf.get((crowd[info["category_id"]], p), 0)
        if inter / info["area"] > 0.5:
            continue
        stat[info["category_id"]].fp += 1


def frame_segments(gt_segments, pred_segments, pred_ids, categories):
    """per-frame bookkeeping of eval_vpq.py:90-116 (predicted areas recounted from the map, sanity checks)"""
    gt_segms, pred_segms = merge_segments(gt_segments), merge_segments(pred_segments)
    left = set(el["id"] for el in pred_segments)
    labels, cnt = np.unique(np.asarray(pred_ids), return_counts=True)
    for lab, c in zip(labels.tolist(), cnt.tolist()):
        if lab not in pred_segms:
            if lab == VOID:
                continue
            raise KeyError("Segment with ID %d is presented in PNG and not presented in JSON." % lab)
        pred_segms[lab]["area"] = c
        left.remove(lab)
        if pred_segms[lab]["category_id"] not in categories:
            raise KeyError("Segment with ID %d has unknown category_id" % lab)
    if left:
        raise KeyError("segment IDs %s are presented in JSON and not presented in PNG." % sorted(left))
    return gt_segms, pred_segms


def accumulate_tube_areas(per_frame):
    vid = {}
    for segms in per_frame:
        for k, v in segms.items():
            if k not in vid:
                vid[k] = v
            else:
                vid[k]["area"] += v["area"]
    return vid


def vpq_compute_single_core(frames, categories, nframes=2):
    """frames: list of (gt_segments, pred_segments, gt_ids [H,W], pred_ids [H,W]).  Returns dict category -> CatStat."""
    stat = defaultdict(CatStat)
    for idx in range(0, len(frames) - nframes + 1):
        gts, preds, gl, pl = [], [], [], []
        for gt_segments, pred_segments, gt_ids, pred_ids in frames[idx:idx + nframes]:
            g, p = frame_segments(gt_segments, pred_segments, pred_ids, categories)
            gl.append(g); pl.append(p)
            gts.append(np.asarray(gt_ids)); preds.append(np.asarray(pred_ids))
        vid_gt, vid_pred = accumulate_tube_areas(gl), accumulate_tube_areas(pl)
        pairs, counts = tube_confusion(np.stack(gts), np.stack(preds))
        match_tube(stat, vid_gt, vid_pred, pairs, counts)
    return stat


def pq_average(stat, categories, isthing=None):
    """PQStat.pq_average (eval_vpq.py:44-70)"""
    pq = sq = rq = 0.0
    n = 0
    per_class = {}
    for label, info in categories.items():
        if isthing is not None and isthing != (info["isthing"] == 1):
            continue
        c = stat[label]
        if c.tp + c.fp + c.fn == 0:
            per_class[label] = {"pq": 0.0, "sq": 0.0, "rq": 0.0, "iou": 0.0, "tp": 0, "fp": 0, "fn": 0}
            continue
        n += 1
        pq_c = c.iou / (c.tp + 0.5 * c.fp + 0.5 * c.fn)
        sq_c = c.iou / c.tp if c.tp != 0 else 0
        rq_c = c.tp / (c.tp + 0.5 * c.fp + 0.5 * c.fn)
        per_class[label] = {"pq": pq_c, "sq": sq_c, "rq": rq_c, "iou": c.iou, "tp": c.tp, "fp": c.fp, "fn": c.fn}
        pq += pq_c; sq += sq_c; rq += rq_c
    return {"pq": pq / n, "sq": sq / n, "rq": rq / n, "n": n}, per_class


def segments_from_pan2ch(pan_2ch, num_stuff=11):
    """numpy counterpart of vps_b200.vpq.segments_from_pan2ch: the segmentation converter_2ch_track_core produces
    (tools/dataset/cityscapes_vps.py:96-140) -- one segment per stuff category (panopticapi's IdGenerator returns the
    category's fixed colour for every key of a stuff class), one per (thing category, track) key -- with deterministic ids
    instead of colours: returns (ids uint32 [H,W], segments)."""
    p = np.asarray(pan_2ch).astype(np.uint32)
    sem, trk = p[..., 0], p[..., 2]
    ids = np.where(sem == 255, 0, 1000 * sem + np.where(sem < num_stuff, 0, trk) + 1).astype(np.uint32)
    segs = []
    for i, a in zip(*np.unique(ids, return_counts=True)):
        if i == VOID:
            continue
        segs.append({"id": int(i), "category_id": int((int(i) - 1) // 1000), "iscrowd": 0, "area": int(a)})
    return ids, segs

