"""ncu driver: warm up, then run ONE full-size FuseTrack step between cudaProfilerStart/Stop.
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/one_step.py
(numbers printed under ncu are not bench values)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import build_product, meta, synth_pairs  # noqa: E402

H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1024, 2048)
dev = torch.device("cuda:0")
det = build_product("bf16", dev)
pairs = [(a.to(dev), b.to(dev)) for a, b in synth_pairs(2, H, W)]
for i in range(3):
    det.simple_test(pairs[i % 2][0], [meta(10001 + i, H, W)], ref_img=[pairs[i % 2][1]])
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
det.simple_test(pairs[1][0], [meta(10004, H, W)], ref_img=[pairs[1][1]])
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("one step done")
