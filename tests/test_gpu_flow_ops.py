"""GPU parity: correlation / resample2d / channelnorm kernels vs the oracle restatements."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _nhwc(t, dtype=torch.float32):
    return t.permute(0, 2, 3, 1).contiguous().to(dtype)


@pytest.mark.parametrize("cfg", [(20, 2, 64, 24, 40), (4, 1, 48, 20, 36), (20, 2, 256, 16, 70), (4, 1, 256, 33, 31)])
def test_correlation(cuda, cfg):
    from oracle import ops as O
    from vps_b200 import ops
    md, s2, C, H, W = cfg
    g = torch.Generator().manual_seed(3)
    f1 = torch.randn(1, C, H, W, generator=g)
    f2 = torch.randn(1, C, H, W, generator=g)
    ref = O.correlation(f1, f2, md, 1, md, 1, s2)
    D = 2 * (md // s2) + 1
    out = torch.empty(1, H, W, D * D, device=cuda)
    ops.correlation(_nhwc(f1).to(cuda), _nhwc(f2).to(cuda), out, md, md, 1, s2)
    torch.cuda.synchronize()
    got = out.cpu().permute(0, 3, 1, 2)
    assert (got - ref).abs().max().item() <= 1e-5


def test_resample2d_and_channelnorm(cuda):
    from oracle import ops as O
    from vps_b200 import ops
    g = torch.Generator().manual_seed(5)
    src = torch.randn(2, 3, 40, 56, generator=g)
    flow = (torch.rand(2, 2, 40, 56, generator=g) - 0.5) * 30
    ref = O.resample2d(src, flow)
    out = torch.empty(2, 40, 56, 3, device=cuda)
    ops.resample2d(_nhwc(src).to(cuda), _nhwc(flow).to(cuda), out)
    nrm = torch.empty(2, 40, 56, 1, device=cuda)
    ops.channelnorm(_nhwc(src).to(cuda), nrm, b=out)
    torch.cuda.synchronize()
    assert (out.cpu().permute(0, 3, 1, 2) - ref).abs().max().item() <= 1e-5
    refn = O.channelnorm(src - ref)
    assert (nrm.cpu().permute(0, 3, 1, 2) - refn).abs().max().item() <= 1e-5


@pytest.mark.parametrize("cfg", [(20, 2, 256, 128, 256), (20, 2, 64, 37, 53), (4, 1, 256, 64, 96), (4, 1, 128, 21, 50),
                                 (20, 2, 192, 9, 11), (4, 1, 64, 3, 5)])
def test_correlation_tensor_core(cuda, cfg):
    """banded-GEMM correlation on tcgen05 vs the oracle on bf16-representable features (fp32 accumulation both)."""
    from oracle import ops as O
    from vps_b200 import ops
    md, s2, C, H, W = cfg
    g = torch.Generator().manual_seed(17)
    f1 = torch.randn(1, C, H, W, generator=g).bfloat16().float()
    f2 = torch.randn(1, C, H, W, generator=g).bfloat16().float()
    ref = torch.nn.functional.leaky_relu(O.correlation(f1, f2, md, 1, md, 1, s2), 0.1)
    D = 2 * (md // s2) + 1
    cs = (D * D + 7) // 8 * 8
    for odt, tol in ((torch.float32, 2e-5), (torch.bfloat16, 1e-2)):
        buf = torch.full((1, H, W, cs + 8), 7.0, dtype=odt, device=cuda)
        out = buf[..., 8:8 + D * D]
        ops.correlation(_nhwc(f1, torch.bfloat16).to(cuda), _nhwc(f2, torch.bfloat16).to(cuda), out, md, md, 1, s2,
                        act=ops.ACT_LRELU, slope=0.1, impl="tc")
        torch.cuda.synchronize()
        got = out.float().cpu().permute(0, 3, 1, 2)
        assert (got - ref).abs().max().item() <= tol * max(1.0, ref.abs().max().item()), (cfg, odt)
        assert (buf[..., :8] == 7.0).all() and (buf[..., 8 + D * D:] == 7.0).all()     # neighbours of the slice untouched
