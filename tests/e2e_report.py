"""Diagnostic (not a pytest): per-stage oracle-vs-product differences on a small synthetic clip.
Usage: python tests/e2e_report.py [fp32|bf16] [H W]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from tests.e2e_util import build_models, compare_frame, make_pair  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
H, W = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (128, 256)
oracle, prod = build_models("C", 0, prec)
img, ref = make_pair(H, W)
for iid, (a, b) in ((10001, (img, ref)), (10002, (ref, img))):
    t = time.time()
    rep, _, _ = compare_frame(oracle, prod, a, b, iid)
    print("frame iid", iid, "(%.1fs)" % (time.time() - t))
    for k, v in rep.items():
        print("  %-16s %s" % (k, v))
