"""The operand split of the parity precision, emulated on the CPU: v = hi + 2^-11 lo with hi = fp16(v), lo = fp16(2^11 (v - hi)),
product a.b ~= hi.hi + 2^-11 (hi.lo + lo.hi).  Documents the accuracy class the tensor-core kernels (conv_tc32.cu, the three-pass
correlation of corr_tc.cu) are built on: fp32-class, ~4 orders of magnitude below one bf16 pass."""
import torch


def _split(v):
    hi = v.half().float()
    lo = ((v - hi) * 2048.0).half().float()
    return hi, lo


def test_split_reconstructs_fp32_values():
    g = torch.Generator().manual_seed(0)
    v = torch.randn(1 << 16, generator=g) * torch.logspace(-6, 3, 1 << 16)
    hi, lo = _split(v)
    assert (v - hi).abs().max() <= (v.abs() * 2.0 ** -11).max()          # the residual is at most half an fp16 ulp: scaling cannot overflow
    rec = hi.double() + lo.double() / 2048.0
    rel = ((rec - v.double()).abs() / v.double().abs().clamp_min(1e-30))
    assert float(rel[v.abs() > 1e-4].max()) <= 2.0 ** -21                 # ~22 significant bits for normal-range values


def test_three_product_contraction_is_fp32_class():
    g = torch.Generator().manual_seed(1)
    a = torch.randn(64, 256, generator=g)
    b = torch.randn(256, 96, generator=g)
    ref = a.double() @ b.double()
    ah, al = _split(a)
    bh, bl = _split(b)
    got = (ah.double() @ bh.double()) + (ah.double() @ bl.double() + al.double() @ bh.double()) / 2048.0
    scale = float(ref.abs().max())
    err3 = float((got - ref).abs().max()) / scale
    err_bf16 = float((a.bfloat16().double() @ b.bfloat16().double() - ref).abs().max()) / scale
    err_fp32 = float(((a @ b).double() - ref).abs().max()) / scale
    assert err3 <= 3e-7, err3                       # the dropped lo.lo term and the fp16 rounding of lo: ~2^-22 per product
    assert err3 <= 4 * max(err_fp32, 1e-7)          # same class as an fp32 matmul
    assert err_bf16 >= 1000 * err3                  # one bf16 pass is 3-4 orders of magnitude away
