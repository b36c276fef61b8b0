"""ncu driver: warm up, then run ONE full-size FuseTrack step between cudaProfilerStart/Stop.
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/one_step.py [tc32|bf16|fp32] [H W]
(numbers printed under ncu are not bench values)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import build_product, meta, synth_pairs  # noqa: E402

args = [a for a in sys.argv[1:] if a not in ("tc32", "bf16", "fp32")]
prec = ([a for a in sys.argv[1:] if a in ("tc32", "bf16", "fp32")] or ["tc32"])[0]
H, W = (int(args[0]), int(args[1])) if len(args) > 1 else (1024, 2048)
dev = torch.device("cuda:0")
det = build_product(prec, dev)
pairs = [(a.to(dev), b.to(dev)) for a, b in synth_pairs(2, H, W)]
for i in range(3):
    det.simple_test(pairs[i % 2][0], [meta(10001 + i, H, W)], ref_img=[pairs[i % 2][1]])
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
det.simple_test(pairs[1][0], [meta(10004, H, W)], ref_img=[pairs[1][1]])
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("one step done")
