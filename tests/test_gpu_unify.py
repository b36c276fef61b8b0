"""GPU parity of the unified-pan-result step: vps_unify_pan via vps_b200.postproc.PanUnifier vs the
oracle (oracle/unify.py, itself pinned to the reference's function) and vs the golden frames produced by the reference."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "unify_pan.npz")


def _frames(rng, H, W, ks):
    from tests.golden.make_unify_golden import synth_frame
    return [synth_frame(rng, H, W, k) for k in ks]


def test_unify_matches_reference_golden(cuda):
    """the frames the reference's own get_unified_pan_result was run on; uint8 and int64 label maps, with / without ids"""
    from vps_b200.postproc import PanUnifier
    d = np.load(GOLD)
    n = int(d["nframes"])
    for dtype in (torch.uint8, torch.int64):
        u, u0 = PanUnifier(), PanUnifier()
        for i in range(n):
            seg = torch.from_numpy(d["seg%d" % i]).to(dtype).cuda()
            pan = torch.from_numpy(d["pan%d" % i]).to(dtype).cuda()
            got = u(seg[None], pan[None], d["cls%d" % i], d["obj%d" % i]).cpu().numpy()
            got0 = u0(seg, pan, torch.from_numpy(d["cls%d" % i]), None).cpu().numpy()
            assert np.array_equal(got, d["out%d" % i]), "frame %d" % i
            assert np.array_equal(got0, d["out_noid%d" % i]), "frame %d (no ids)" % i


@pytest.mark.parametrize("H,W", [(1024, 2048), (250, 333), (64, 64)])
def test_unify_matches_oracle_random(cuda, H, W):
    """full-size frames, odd sizes (pixel count not a multiple of 4), many instances, large track ids (uint8 wrap)"""
    from oracle import unify as U
    from vps_b200.postproc import PanUnifier
    rng = np.random.default_rng(H * 7 + W)
    ks = [40, 100, 3]
    frames = _frames(rng, H, W, ks)
    segs, pans, clss, objs = zip(*frames)
    objs = [o + 240 for o in objs]                         # track id + 1 wraps past 255
    ref = U.get_unified_pan_result(segs, pans, clss, objs, stuff_area_limit=min(4 * 64 * 64, H * W // 40))
    u = PanUnifier(stuff_area_limit=min(4 * 64 * 64, H * W // 40))
    for i in range(len(ks)):
        got = u(torch.from_numpy(segs[i]).cuda(), torch.from_numpy(pans[i]).cuda(), clss[i], objs[i]).cpu().numpy()
        assert np.array_equal(got, ref[i]), "frame %d" % i
    assert u.max_oid > 100


def test_unify_no_instances_and_cpu_inputs_fail(cuda):
    from oracle import unify as U
    from vps_b200.postproc import PanUnifier
    rng = np.random.default_rng(3)
    seg = rng.integers(0, 11, size=(128, 256)).astype(np.uint8)
    ref = U.unify_frame(seg, seg.copy(), np.zeros(0, np.int64), None)
    got = PanUnifier()(torch.from_numpy(seg).cuda(), torch.from_numpy(seg.copy()).cuda(), np.zeros(0, np.int64), None).cpu().numpy()
    assert np.array_equal(got, ref)
    with pytest.raises(RuntimeError):
        PanUnifier()(torch.from_numpy(seg), torch.from_numpy(seg), np.zeros(0, np.int64))


def test_unify_flags_instance_id_without_class(cuda):
    """a panoptic id beyond len(cls_ind): the reference raises IndexError (cityscapes_vps.py:197); PanUnifier.check() does too"""
    import numpy as np
    from vps_b200.postproc import PanUnifier
    H, W = 64, 96
    seg = torch.zeros(H, W, dtype=torch.uint8, device="cuda")
    pan = torch.zeros(H, W, dtype=torch.uint8, device="cuda")
    pan[10:20, 10:30] = 11 + 3                      # instance j = 3, but only 2 classes are given
    u = PanUnifier(stuff_area_limit=16)
    u(seg, pan, np.array([1, 2]), np.array([0, 1]))
    with pytest.raises(IndexError):
        u.check()
    pan[10:20, 10:30] = 11 + 1
    u(seg, pan, np.array([1, 2]), np.array([0, 1]))
    u.check()
