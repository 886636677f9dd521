"""The numerical scheme of the tensor-core dense layers, emulated on the CPU with torch.half (no GPU): fp16 (hi, lo) split
pairs with power-of-two operand scaling, three products per K-step (lo*hi + hi*lo + hi*hi), fp32-or-better accumulation.
Pins the error model DESIGN.md section 3 / 5.1 states: 22-bit operands when the scaled values stay clear of fp16's
subnormals, graceful ABSOLUTE degradation below, and why the scaling is needed."""
import math

import torch


def split(x, e):
    xs = x * 2.0 ** e
    hi = xs.half()
    lo = (xs - hi.float()).half()
    return hi, lo


def weight_exp(w):
    return 14 - math.ceil(math.log2(float(w.abs().max())))


def gemm3(a, ea, w, ew):
    ah, al = split(a, ea)
    wh, wl = split(w, ew)
    ah, al, wh, wl = (t.double() for t in (ah, al, wh, wl))
    return (al @ wh.t() + ah @ wl.t() + ah @ wh.t()) * 2.0 ** -(ea + ew)


def test_pair_represents_22_bits_inside_the_scaled_range():
    torch.manual_seed(0)
    x = torch.randn(4096) * torch.logspace(-2, 2, 4096)            # 1e-2 .. 1e2
    hi, lo = split(x, 6)
    assert torch.isfinite(hi.float()).all()
    err = ((hi.double() + lo.double()) * 2.0 ** -6 - x.double()).abs()
    full = x.abs() >= 2.0 ** -9                                    # |x| * 2^6 >= 2^-3: the lo part is a normal fp16 number
    assert int(full.sum()) > 3000 and float((err[full] / x.double().abs()[full]).max()) <= 2.0 ** -21
    assert float(err[~full].max()) <= 2.0 ** -25 * 2.0 ** -6       # below: absolute half-ulp of an fp16 subnormal, rescaled


def test_three_products_match_fp64_like_fp32_matmul_does():
    torch.manual_seed(1)
    a = torch.randn(512, 256).clamp_min(0)
    w = torch.randn(256, 256) / 16
    exact = a.double() @ w.double().t()
    err = float((gemm3(a, 6, w, weight_exp(w)) - exact).abs().max() / exact.abs().max())
    err32 = float(((a @ w.t()).double() - exact).abs().max() / exact.abs().max())
    assert err <= 2e-7 and err <= err32 * 2


def test_scaling_is_what_keeps_small_operands_precise():
    torch.manual_seed(2)
    a = torch.randn(256, 128) * 1e-3
    w = torch.randn(64, 128) / 11
    exact = a.double() @ w.double().t()
    ref = exact.abs().max()
    unscaled = float((gemm3(a, 0, w, 0) - exact).abs().max() / ref)
    scaled = float((gemm3(a, 6, w, weight_exp(w)) - exact).abs().max() / ref)
    assert unscaled > 1e-6                                        # lo parts fall into fp16 subnormals
    assert scaled <= 5e-7


def test_overflow_threshold_of_the_default_activation_exponent():
    assert math.isinf(float(torch.tensor(1100.0 * 2.0 ** 6).half()))      # what NFK_FLAG_F16_RANGE reports
    assert math.isfinite(float(torch.tensor(1000.0 * 2.0 ** 6).half()))
