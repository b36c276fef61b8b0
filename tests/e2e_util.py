"""Shared helpers for the end-to-end parity tests: build oracle + product with identical synthetic weights,
run both on the same seeded frame pair(s), and report per-stage differences."""
import numpy as np
import torch


def nchw(t):
    """product NHWC device tensor -> NCHW fp32 cpu"""
    return t.float().permute(0, 3, 1, 2).contiguous().cpu()


def make_pair(H, W, seed=1, shift=(2, 3)):
    g = torch.Generator().manual_seed(seed)
    img = torch.randn(1, 3, H, W, generator=g)
    ref = torch.roll(img, shifts=shift, dims=(2, 3)) + 0.05 * torch.randn(1, 3, H, W, generator=g)
    return img, ref


def build_models(kind="C", seed=0, precision="fp32", device="cuda:0"):
    from oracle.weights import make_model
    from vps_b200 import ConfigDict, build_detector, fusetrack_cfg
    oracle = make_model(kind, seed)
    cfg = fusetrack_cfg()
    prod = build_detector(ConfigDict(cfg["model"]), train_cfg=None, test_cfg=ConfigDict(cfg["test_cfg"]))
    prod.load_state_dict(oracle.state_dict(), strict=True)
    prod.precision = precision
    prod = prod.to(device)
    return oracle, prod


def meta(iid, H, W):
    return dict(filename="synthetic_city_%06d.png" % iid, iid=iid, img_shape=(H, W, 3), pad_shape=(H, W, 3),
                ori_shape=(H, W, 3), scale_factor=1.0)


def rel_err(a, b):
    a, b = a.float(), b.float()
    return float((a - b).abs().max() / max(1e-6, float(b.abs().max())))


def compare_frame(oracle, prod, img, ref, iid, device="cuda:0"):
    """Run one frame through both; returns dict stage -> error metrics."""
    H, W = img.shape[-2:]
    ot, pt = {}, {}
    o_res = oracle.simple_test(img, dict(iid=iid, img_shape=(H, W, 3)), ref, ot)
    p_res = prod.simple_test(img.to(device), [meta(iid, H, W)], ref_img=[ref.to(device)], taps=pt)
    torch.cuda.synchronize()
    rep = {}
    rep["flow_full"] = rel_err(nchw(pt["flow_full"]), ot["flow_full"])
    rep["flow"] = rel_err(nchw(pt["flow"]), ot["flow"])
    for i in range(5):
        rep["fpn%d" % i] = rel_err(nchw(pt["fpn"][i]), ot["fpn"][i])
        rep["fused%d" % i] = rel_err(nchw(pt["fused"][i]), ot["fused"][i])
    rep["bsf"] = rel_err(nchw(pt["bsf"]), ot["bsf"])
    rep["warp_bsf"] = rel_err(nchw(pt["warp_bsf"]), ot["warp_bsf"])
    rep["flow_fine"] = rel_err(nchw(pt["flow_fine"]), ot["flow_fine"])
    rep["refined"] = rel_err(nchw(pt["refined"]), ot["refined"])
    rep["fcn_score"] = rel_err(nchw(pt["fcn_score"]), ot["fcn_score"])
    rep["fcn_score_abs"] = float((nchw(pt["fcn_score"]) - ot["fcn_score"]).abs().max())
    rep["fcn_output_abs"] = float((nchw(pt["fcn_output"]) - ot["fcn_output"]).abs().max())
    # RPN
    A = 3
    for l in range(5):
        hd = nchw(pt["rpn_heads"][l])
        rep["rpn_cls%d" % l] = float((hd[:, :A] - ot["rpn_cls"][l]).abs().max())
        rep["rpn_reg%d" % l] = float((hd[:, A:5 * A] - ot["rpn_reg"][l]).abs().max())
    npp = int(pt["nprop"].item())
    pp = pt["proposals"][:npp].cpu()
    op = ot["proposals"]
    rep["n_proposals"] = (npp, op.shape[0])
    if npp == op.shape[0]:
        rep["proposals_abs"] = float((pp - op).abs().max())
        rep["cls_score_abs"] = float((pt["cls_score"][:npp].cpu() - ot["cls_score"]).abs().max())
        rep["bbox_pred_abs"] = float((pt["bbox_pred"][:npp].cpu() - ot["bbox_pred"]).abs().max())
    k_p, k_o = pt["det_rois"].shape[0], ot["det_rois"].shape[0]
    rep["n_det"] = (k_p, k_o)
    if k_p == k_o:
        rep["det_rois_abs"] = float((pt["det_rois"].cpu() - ot["det_rois"]).abs().max())
        rep["cls_idx_equal"] = bool((pt["cls_idx"].cpu().long() == ot["cls_idx"]).all())
        rep["cls_prob_abs"] = float((pt["cls_prob"].cpu() - ot["cls_prob"]).abs().max())
        rep["mask_logit_abs"] = float((pt["mask_logit"].cpu() - ot["mask_score"][:, 0]).abs().max())
        rep["obj_ids_equal"] = bool(np.array_equal(np.asarray(pt["det_obj_ids_all"]), ot["det_obj_ids_all"].numpy()))
        rep["keep_equal"] = bool(np.array_equal(np.asarray(pt["keep_inds"]), ot["keep_inds"].numpy()))
    po, oo = p_res[2]["panoptic_outputs"].cpu(), o_res[2]["panoptic_outputs"]
    rep["pano_agree"] = float((po == oo).float().mean())
    rep["sem_agree"] = float((p_res[2]["fcn_outputs"].cpu() == o_res[2]["fcn_outputs"]).float().mean())
    rep["ids_kept_equal"] = bool(np.array_equal(p_res[2]["panoptic_det_obj_ids"].cpu().numpy(),
                                                o_res[2]["panoptic_det_obj_ids"].numpy()))
    return rep, (o_res, ot), (p_res, pt)


def near_tie_report(prod_map, oracle_map, oracle_logits, tol, exclude=None):
    """Label maps of two correct fp32 implementations may differ where the argmax is decided by less than their logit
    error.  Returns (pixels that differ, differing pixels NOT explained by an oracle top-2 logit margin <= tol).
    `exclude`: optional bool map of pixels left out of the comparison."""
    pm, om = prod_map.reshape(-1).long(), oracle_map.reshape(-1).long()
    diff = pm != om
    if exclude is not None:
        diff = diff & ~exclude.reshape(-1)
    bad = diff.nonzero().flatten()
    if bad.numel() == 0:
        return 0, 0
    lg = oracle_logits.reshape(oracle_logits.shape[1], -1)[:, bad]          # [C, nbad]
    top2 = lg.topk(2, dim=0).values
    margin = top2[0] - top2[1]
    # the product's label must be one of the (near-)tied candidates
    prod_logit = lg.gather(0, pm[bad].clamp(max=lg.shape[0] - 1).view(1, -1))[0]
    explained = (margin <= tol) & ((top2[0] - prod_logit) <= tol)
    return int(bad.numel()), int((~explained).sum())
