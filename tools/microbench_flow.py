"""BASELINE config 5: FlowNet2 correlation + resample2d microbench on 256-channel 128x256 feature maps.
Reports achieved HBM GB/s (algorithmic bytes / CUDA-event time) against MEASURED_PEAKS.json."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from vps_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
H, W, C = 128, 256, 256
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6650.0
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    evs = []
    for i in range(iters):
        flush.fill_(i)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in evs)
    return ms[len(ms) // 2]


res = {}
for dt, name in ((torch.bfloat16, "bf16"), (torch.float32, "f32")):
    esz = 2 if dt == torch.bfloat16 else 4
    f1 = torch.randn(1, H, W, C, generator=g).to(dev).to(dt)
    f2 = torch.randn(1, H, W, C, generator=g).to(dev).to(dt)
    out = torch.empty(1, H, W, 448, dtype=dt, device=dev)[..., :441]
    impls = ("tc", "simt") if dt == torch.bfloat16 else ("tc32", "simt")     # tc32: split fp16 planes, 3 tensor-core passes
    for impl in impls:
        ms = timeit(lambda: ops.correlation(f1, f2, out, 20, 20, 1, 2, impl=impl))
        by = (2 * H * W * C + H * W * 441) * esz
        res["correlation_%s_%s" % (name, impl)] = {"ms": ms, "GB/s": by / ms / 1e6, "frac_of_hbm_peak": by / ms / 1e6 / peak,
                                                  "TFLOP/s": 2 * H * W * C * 441 / ms / 1e9}
    flow = ((torch.rand(1, H, W, 2, generator=g) - 0.5) * 8).to(dev)
    o2 = torch.empty_like(f1)
    ms = timeit(lambda: ops.resample2d(f1, flow, o2))
    by = 2 * H * W * C * esz + H * W * 2 * 4
    res["resample2d_%s" % name] = {"ms": ms, "GB/s": by / ms / 1e6, "frac_of_hbm_peak": by / ms / 1e6 / peak}
print(json.dumps({"config": "BASELINE cfg5: 256ch 128x256, corr pad20 d20 s2=2 -> 441ch; resample2d bilinear", "hbm_peak_GBs": peak,
                  "results": res}, indent=1))
