"""GPU: the reference-side ctypes shims of INTEGRATION.md (integration/reference_shims.py -- raw C ABI, no vps_b200
import) produce what the native ops they replace produce: nms_cuda.nms, resample2d_cuda.forward, channelnorm_cuda.forward."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_nms_shim_matches_reference_contract(cuda):
    from integration import reference_shims as S
    from oracle import ops as O
    g = torch.Generator().manual_seed(3)
    for n in (1, 7, 300, 1000):
        xy = torch.rand(n, 2, generator=g) * 200
        wh = torch.rand(n, 2, generator=g) * 60 + 2
        dets = torch.cat([xy, xy + wh, torch.rand(n, 1, generator=g)], 1)
        got = S.nms_cuda_nms(dets.cuda(), 0.5)
        ref_dets, ref_inds = O.nms(dets, 0.5)
        assert got.dtype == torch.long and torch.equal(got.cpu(), ref_inds.sort()[0]), n
    assert S.nms_cuda_nms(torch.empty(0, 5, device="cuda"), 0.5).numel() == 0


def test_flow_op_shims(cuda):
    from integration import reference_shims as S
    from oracle import ops as O
    g = torch.Generator().manual_seed(4)
    x = torch.randn(1, 3, 24, 40, generator=g)
    flow = (torch.rand(1, 2, 24, 40, generator=g) - 0.5) * 12
    out = torch.empty(1, 3, 24, 40, device="cuda")
    S.resample2d_forward(x.cuda(), flow.cuda(), out)
    assert float((out.cpu() - O.resample2d(x, flow)).abs().max()) <= 1e-5
    o2 = torch.empty(1, 1, 24, 40, device="cuda")
    S.channelnorm_forward(x.cuda(), o2)
    assert float((o2.cpu() - O.channelnorm(x)).abs().max()) <= 1e-5
