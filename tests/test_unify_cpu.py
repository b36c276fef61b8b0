"""CPU: the oracle of the unified-pan-result step (oracle/unify.py) reproduces the reference's own
get_unified_pan_result bit for bit on the golden frames (tests/golden/make_unify_golden.py generated them by importing
/root/reference/tools/dataset/cityscapes_vps.py)."""
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(__file__), "golden", "unify_pan.npz")


def _load():
    d = np.load(GOLD)
    n = int(d["nframes"])
    return d, n


def test_oracle_matches_reference_golden():
    from oracle import unify as U
    d, n = _load()
    segs = [d["seg%d" % i] for i in range(n)]
    pans = [d["pan%d" % i] for i in range(n)]
    clss = [d["cls%d" % i] for i in range(n)]
    objs = [d["obj%d" % i] for i in range(n)]
    got = U.get_unified_pan_result(segs, pans, clss, objs)
    got_noid = U.get_unified_pan_result(segs, pans, clss, None)
    for i in range(n):
        assert np.array_equal(got[i], d["out%d" % i]), "frame %d (with track ids)" % i
        assert np.array_equal(got_noid[i], d["out_noid%d" % i]), "frame %d (no track ids)" % i
    # the golden set exercises every branch
    assert any((d["out%d" % i][..., 0] == 255).any() for i in range(n))           # small stuff areas / id 255
    assert any((d["out%d" % i][..., 1] > 0).any() for i in range(n))
    assert len(np.unique(d["obj0"])) < len(d["obj0"])                              # duplicate track ids present


def test_dedup_track_ids_counter_runs_across_frames():
    from oracle import unify as U
    a, m = U.dedup_track_ids(np.array([3, 5, 3, 3, 5]), 100)
    assert a.tolist() == [101, 102, 100, 3, 5] and m == 103       # last occurrence keeps the id, earlier ones walk backwards
    b, m = U.dedup_track_ids(np.array([7, 7]), m)
    assert b.tolist() == [103, 7] and m == 104


def test_product_dedup_matches_oracle():
    """PanUnifier's host-side duplicate-track-id bookkeeping (the only state of the unify row) equals the oracle's over a
    sequence of frames (no GPU needed)."""
    from oracle import unify as U
    from vps_b200.postproc import PanUnifier
    rng = np.random.default_rng(1)
    u, m = PanUnifier(), 100
    for _ in range(20):
        ids = rng.integers(0, 12, size=int(rng.integers(0, 15)))
        ref, m = U.dedup_track_ids(ids, m)
        assert u.dedup_track_ids(ids).tolist() == ref.tolist() and u.max_oid == m
