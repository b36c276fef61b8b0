"""GPU: the input stage -- Normalize + Pad(32) + ImageToTensor of a uint8 BGR frame on the device is
bit-identical to the reference pipeline's host arithmetic (oracle/pipeline.py restates mmcv 0.2.14), odd sizes included, and
feeding the detector from uint8 frames gives the same results as feeding it the host-normalised fp32 tensors."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("h,w", [(1024, 2048), (1000, 2048), (64, 128)])
def test_input_stage_bit_exact(cuda, h, w):
    from oracle import pipeline as OP
    from vps_b200.pipeline import CITYSCAPES_NORM, InputStage
    rng = np.random.default_rng(h + w)
    img = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    st = InputStage(img_scale=(max(h, w), min(h, w)))
    got, meta = st(torch.from_numpy(img))
    want = OP.prepare_frame(img, CITYSCAPES_NORM["mean"], CITYSCAPES_NORM["std"], True, 32)
    assert tuple(got.shape) == want.shape and meta["pad_shape"] == (want.shape[2], want.shape[3], 3) and meta["img_shape"] == (h, w, 3)
    assert np.array_equal(got.cpu().numpy(), want)
    with pytest.raises(NotImplementedError):
        InputStage()(torch.zeros(100, 300, 3, dtype=torch.uint8))


def test_detector_from_uint8_frames(cuda):
    from oracle import pipeline as OP
    from tests.e2e_util import build_models, meta
    from vps_b200.pipeline import CITYSCAPES_NORM, InputStage
    _, prod = build_models("C", 0, "tc32", "cuda:0")
    H, W = 128, 256
    rng = np.random.default_rng(5)
    a = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)
    b = np.roll(a, (2, 3), axis=(0, 1))
    st = InputStage(img_scale=(W, H))
    xa, _ = st(torch.from_numpy(a)); xb, _ = st(torch.from_numpy(b))
    prod.reset_tracker()
    r1 = prod.simple_test(xa, [meta(10001, H, W)], ref_img=[xb])
    fa = torch.from_numpy(OP.prepare_frame(a, CITYSCAPES_NORM["mean"], CITYSCAPES_NORM["std"])).cuda()
    fb = torch.from_numpy(OP.prepare_frame(b, CITYSCAPES_NORM["mean"], CITYSCAPES_NORM["std"])).cuda()
    prod.reset_tracker()
    r2 = prod.simple_test(fa, [meta(10001, H, W)], ref_img=[fb])
    assert torch.equal(r1[2]["panoptic_outputs"], r2[2]["panoptic_outputs"]) and torch.equal(r1[2]["fcn_outputs"], r2[2]["fcn_outputs"])


def test_clip_runner_from_uint8_frames(cuda):
    """ClipRunner(input_stage=...) fed with pinned uint8 frames == ClipRunner fed with the host-normalised fp32 tensors"""
    from oracle import pipeline as OP
    from tests.e2e_util import build_models, meta
    from vps_b200.pipeline import CITYSCAPES_NORM, InputStage
    from vps_b200.runner import ClipRunner
    _, prod = build_models("C", 0, "tc32", "cuda:0")
    H, W = 128, 256
    rng = np.random.default_rng(9)
    frames = [rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8) for _ in range(5)]
    pairs_u8 = [(torch.from_numpy(frames[t]).pin_memory(), torch.from_numpy(frames[max(t - 1, 0)]).pin_memory()) for t in range(5)]
    f32 = [torch.from_numpy(OP.prepare_frame(f, CITYSCAPES_NORM["mean"], CITYSCAPES_NORM["std"])) for f in frames]
    pairs_f32 = [(f32[t].pin_memory(), f32[max(t - 1, 0)].pin_memory()) for t in range(5)]
    metas = [meta(10001 + t, H, W) for t in range(5)]
    outs = []
    try:
        prod.label_dtype = torch.uint8
        for pairs, stage in ((pairs_f32, None), (pairs_u8, InputStage(img_scale=(W, H)))):
            prod.reset_tracker()
            res = [(r[2]["panoptic_outputs"].clone(), r[2]["fcn_outputs"].clone(), r[2]["panoptic_det_obj_ids"].cpu().clone())
                   for r in ClipRunner(prod, "cuda:0", input_stage=stage).run(pairs, metas)]
            outs.append(res)
    finally:
        prod.label_dtype = torch.int64
    for a, b in zip(*outs):
        assert all(torch.equal(x, y) for x, y in zip(a, b))
