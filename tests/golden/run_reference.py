"""Build the REFERENCE PanopticFuseTrack (its own python code, via ref_import stubs) with given weights."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def build_reference_detector(state_dict):
    from tests.golden.ref_import import REF, setup
    M = setup()
    from vps_b200.config import Config
    cfg = Config.fromfile(os.path.join(REF, "configs/cityscapes/fusetrack.py"))
    cfg.model["pretrained"] = None
    fsd = {k[len("flownet2."):]: v for k, v in state_dict.items() if k.startswith("flownet2.")}
    _load = torch.load
    torch.load = lambda *a, **k: {"state_dict": fsd}      # ctor reads work_dirs/flownet/...pth.tar (:100-106)
    _cd = torch.cuda.current_device
    torch.cuda.current_device = lambda: 0                  # ctor prints it with %d (:102-103)
    try:
        det = M.build_detector(cfg.model, train_cfg=None, test_cfg=cfg.test_cfg)
    finally:
        torch.load = _load
        torch.cuda.current_device = _cd
    det.load_state_dict(state_dict, strict=True)
    det.eval()
    return det
