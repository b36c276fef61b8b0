"""Golden vectors for SURVEY 8f rank 2 (VPQ evaluator core): runs the REFERENCE's own `vpq_compute_single_core` and
`PQStat.pq_average` (/root/reference/tools/eval_vpq.py:73-203, 44-70, imported unmodified) on seeded synthetic video
tubes and stores inputs + per-class statistics in tests/golden/vpq_tubes.npz (+ .json for the segment lists).
Run in the build container only:  python tests/golden/make_vpq_golden.py"""
import importlib.util
import json
import os

import numpy as np

REF = "/root/reference/tools/eval_vpq.py"
CATEGORIES = {i: {"id": i, "isthing": 1 if i >= 11 else 0} for i in range(19)}


def synth_clip(rng, nfr, H, W, n_inst=6):
    """ground truth + a perturbed prediction.  ids: stuff segment of class c -> id c + 1 (0 = VOID); instance t of class c ->
    1000 * c + t + 1.  Returns list of (gt_segments, pred_segments, gt_ids, pred_ids)."""
    frames = []
    inst = [(int(rng.integers(11, 19)), int(rng.integers(10, H // 2)), int(rng.integers(10, W // 2)), int(rng.integers(0, H // 2)),
             int(rng.integers(0, W // 2)), int(rng.integers(-3, 4)), int(rng.integers(-3, 4))) for _ in range(n_inst)]
    base = rng.integers(0, 11, size=((H + 15) // 16, (W + 15) // 16)).repeat(16, 0).repeat(16, 1)[:H, :W]
    for f in range(nfr):
        gt = (base + 1).astype(np.uint32)
        gt[:6, :] = 0                                                 # a VOID band
        pr = gt.copy()
        pr[6:12, :] = 3                                               # prediction differs from gt on a band
        for t, (c, h, w, y, x, dy, dx) in enumerate(inst):
            yy, xx = min(max(y + dy * f, 0), H - h), min(max(x + dx * f, 0), W - w)
            gid = 1000 * c + t + 1
            gt[yy:yy + h, xx:xx + w] = gid
            if t % 3 == 0:
                pr[yy + 2:yy + h, xx + 1:xx + w] = gid                # good match
            elif t % 3 == 1:
                pr[yy:yy + h // 3, xx:xx + w // 3] = gid + 500        # poor overlap -> FP + FN
            # t % 3 == 2: missed -> FN
        pr[H - 8:, :20] = 1000 * 15 + 900                             # a hallucinated instance
        crowd_id = 1000 * 12 + 999
        gt[H - 20:H - 8, W - 30:] = crowd_id                          # a crowd region
        pr[H - 20:H - 8, W - 30:] = 1000 * 12 + 700                   # prediction inside the crowd region -> ignored

        def segs(m, is_gt):
            out = []
            ids, cnt = np.unique(m, return_counts=True)
            for i, a in zip(ids.tolist(), cnt.tolist()):
                if i == 0:
                    continue
                cat = i - 1 if i < 1000 else i // 1000
                out.append({"id": int(i), "category_id": int(cat), "iscrowd": int(is_gt and i == crowd_id), "area": int(a)})
            return out
        frames.append((segs(gt, True), segs(pr, False), gt, pr))
    return frames


def to_rgb(ids):
    ids = ids.astype(np.uint32)
    return np.stack([ids % 256, (ids // 256) % 256, ids // 65536], 2).astype(np.uint8)


def main():
    spec = importlib.util.spec_from_file_location("ref_eval_vpq", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    rng = np.random.default_rng(7)
    clips = [synth_clip(rng, 6, 96, 160), synth_clip(rng, 6, 64, 128, n_inst=9)]
    out, meta = {}, {"clips": []}
    for ci, frames in enumerate(clips):
        meta["clips"].append([{"gt": g, "pred": p} for g, p, _, _ in frames])
        for fi, (_, _, gt, pr) in enumerate(frames):
            out["c%d_f%d_gt" % (ci, fi)], out["c%d_f%d_pred" % (ci, fi)] = gt, pr
    for nframes in (1, 2, 3, 4):
        stat = ref.PQStat()
        for frames in clips:
            ref_set = [({"segments_info": g}, {"segments_info": p}, to_rgb(gt), to_rgb(pr), None) for g, p, gt, pr in frames]
            stat += ref.vpq_compute_single_core(ref_set, CATEGORIES, nframes=nframes)
        rows = []
        for c in range(19):
            s = stat[c]
            rows.append([s.iou, s.tp, s.fp, s.fn])
        out["stat_k%d" % nframes] = np.array(rows, dtype=np.float64)
        res = [stat.pq_average(CATEGORIES, isthing=t)[0] for t in (None, True, False)]
        out["avg_k%d" % nframes] = np.array([[r["pq"], r["sq"], r["rq"], r["n"]] for r in res], dtype=np.float64)
    here = os.path.dirname(os.path.abspath(__file__))
    np.savez_compressed(os.path.join(here, "vpq_tubes.npz"), **out)
    json.dump(meta, open(os.path.join(here, "vpq_tubes.json"), "w"))
    print("wrote vpq_tubes.npz / .json", os.path.getsize(os.path.join(here, "vpq_tubes.npz")))


if __name__ == "__main__":
    main()
