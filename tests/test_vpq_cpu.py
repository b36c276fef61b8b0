"""CPU: the oracle of the VPQ evaluator core (oracle/vpq.py) reproduces the reference's vpq_compute_single_core +
pq_average exactly on the golden tubes (tests/golden/make_vpq_golden.py generated them by importing
/root/reference/tools/eval_vpq.py)."""
import json
import os

import numpy as np

HERE = os.path.join(os.path.dirname(__file__), "golden")
CATEGORIES = {i: {"id": i, "isthing": 1 if i >= 11 else 0} for i in range(19)}


def load_clips():
    d = np.load(os.path.join(HERE, "vpq_tubes.npz"))
    meta = json.load(open(os.path.join(HERE, "vpq_tubes.json")))
    clips = []
    for ci, frames in enumerate(meta["clips"]):
        clips.append([(fr["gt"], fr["pred"], d["c%d_f%d_gt" % (ci, fi)], d["c%d_f%d_pred" % (ci, fi)]) for fi, fr in enumerate(frames)])
    return d, clips


def test_oracle_matches_reference_golden():
    from collections import defaultdict

    from oracle import vpq as V
    d, clips = load_clips()
    for nframes in (1, 2, 3, 4):
        stat = defaultdict(V.CatStat)
        for frames in clips:
            s = V.vpq_compute_single_core(frames, CATEGORIES, nframes=nframes)
            for c, v in s.items():
                stat[c].iou += v.iou; stat[c].tp += v.tp; stat[c].fp += v.fp; stat[c].fn += v.fn
        ref = d["stat_k%d" % nframes]
        for c in range(19):
            assert [stat[c].tp, stat[c].fp, stat[c].fn] == ref[c, 1:].astype(int).tolist(), (nframes, c)
            assert abs(stat[c].iou - ref[c, 0]) <= 1e-12, (nframes, c)
        for row, t in enumerate((None, True, False)):
            r, _ = V.pq_average(stat, CATEGORIES, isthing=t)
            assert np.allclose([r["pq"], r["sq"], r["rq"], r["n"]], d["avg_k%d" % nframes][row], rtol=0, atol=1e-12)
    assert d["stat_k2"][:, 1].sum() > 0 and d["stat_k2"][:, 2].sum() > 0 and d["stat_k2"][:, 3].sum() > 0     # TP, FP and FN all occur


def test_product_host_logic_matches_reference_golden():
    """vps_b200.vpq.VpqEvaluator's host side (window merging + the reference's matching logic) fed with numpy-made frame
    tables instead of the device confusion kernel: same statistics as the reference, bit for bit (no GPU needed)."""
    from collections import defaultdict

    from vps_b200 import vpq as P
    d, clips = load_clips()
    for nframes in (1, 2, 3, 4):
        stat = defaultdict(P.CatStat)
        for frames in clips:
            ev = P.VpqEvaluator(CATEGORIES)
            for gseg, pseg, gt, pr in frames:
                pairs, counts = np.unique(gt.astype(np.uint64) * np.uint64(P.OFFSET) + pr.astype(np.uint64), return_counts=True)
                ev.add_frame_table(gseg, pseg, pairs, counts)
            for c, v in ev.compute(nframes).items():
                stat[c] += v
        ref = d["stat_k%d" % nframes]
        for c in range(19):
            assert [stat[c].tp, stat[c].fp, stat[c].fn] == ref[c, 1:].astype(int).tolist(), (nframes, c)
            assert stat[c].iou == ref[c, 0], (nframes, c)
        for row, t in enumerate((None, True, False)):
            r, _ = P.pq_average(stat, CATEGORIES, isthing=t)
            assert [r["pq"], r["sq"], r["rq"], float(r["n"])] == d["avg_k%d" % nframes][row].tolist()
