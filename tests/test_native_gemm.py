"""Dense-layer kernels (FFMA and tcgen05 split-TF32) against an fp64 matmul: both must be fp32-accurate."""
import pytest
import torch

from conftest import rel_err
from nflows_b200 import kernels as K

pytestmark = pytest.mark.gpu

SHAPES = [(1000, 392, 256), (300, 256, 9016), (257, 32, 128), (129, 784, 784), (5, 256, 736), (128, 8, 16), (4096, 128, 736)]


def reference(x, w, b, r, relu_in, relu_out):
    x64 = x.double().clamp_min(0) if relu_in else x.double()
    y = x64 @ w.double().t() + b.double()
    if relu_out:
        y = y.clamp_min(0)
    return y + (r.double() if r is not None else 0)


@torch.no_grad()
@pytest.mark.parametrize("n,k,o", SHAPES)
def test_linear_simt(cuda_device, n, k, o):
    g = torch.Generator(device=cuda_device).manual_seed(n + k + o)
    x = torch.randn(n, k, device=cuda_device, generator=g)
    w = torch.randn(o, k, device=cuda_device, generator=g) / k ** 0.5
    b = torch.randn(o, device=cuda_device, generator=g)
    r = torch.randn(n, o, device=cuda_device, generator=g)
    for relu_in, relu_out, res in ((False, False, None), (True, True, None), (False, False, r)):
        y = K.linear(x, w, b, residual=res, relu_in=relu_in, relu_out=relu_out)
        assert rel_err(y, reference(x, w, b, res, relu_in, relu_out)) <= 8e-6


@torch.no_grad()
@pytest.mark.parametrize("n,k,o", SHAPES)
def test_linear_tf32x3(cuda_device, n, k, o):
    g = torch.Generator(device=cuda_device).manual_seed(n + k + o)
    x = torch.randn(n, k, device=cuda_device, generator=g)
    w = torch.randn(o, k, device=cuda_device, generator=g) / k ** 0.5
    b = torch.randn(o, device=cuda_device, generator=g)
    r = torch.randn(n, o, device=cuda_device, generator=g)
    assert K.tf32x3_supported(k, k, k)
    wp = K.split_tf32(w)
    assert torch.equal(wp[0] + wp[1], w)                       # the split is exact
    for relu_in, relu_out, res in ((False, False, None), (True, True, None), (False, False, r)):
        xp = K.split_tf32(x, relu=relu_in)
        y, pair = K.linear_tf32x3(xp, wp, b, residual=res, relu_out=relu_out, want_y=True, want_split=True, split_relu=True)
        want = reference(x, w, b, res, relu_in, relu_out)
        assert rel_err(y, want) <= 8e-6, (n, k, o, relu_in, relu_out)
        assert torch.equal(pair[0] + pair[1], y.clamp_min(0))
        # hi part is a TF32 number: low 13 mantissa bits are zero
        assert int((pair[0].view(torch.int32) & 0x1FFF).abs().max()) == 0
    # gather + split
    cols = torch.arange(0, k, 2, device=cuda_device, dtype=torch.int32)
    hi, lo = K.split_tf32(x, cols)
    assert torch.equal(hi + lo, x[:, ::2])


@torch.no_grad()
@pytest.mark.parametrize("n,k,o", SHAPES)
def test_linear_tf32x3_fp32_activation_split_on_chip(cuda_device, n, k, o):
    """nfk_linear_tf32x3_a32: A given as plain fp32 (here a strided column block of a wider tensor, as the coupling trunk
    reads the identity features), relu + split done in shared memory; must match the pre-split path bit for bit."""
    g = torch.Generator(device=cuda_device).manual_seed(7 * n + k + o)
    wide = torch.randn(n, 2 * k, device=cuda_device, generator=g)
    x = wide[:, :k]
    w = torch.randn(o, k, device=cuda_device, generator=g) / k ** 0.5
    b = torch.randn(o, device=cuda_device, generator=g)
    r = torch.randn(n, o, device=cuda_device, generator=g)
    wp = K.split_tf32(w)
    for relu_in, relu_out, res in ((False, False, None), (True, True, None), (True, False, r)):
        y, pair = K.linear_tf32x3(x, wp, b, residual=res, relu_in=relu_in, relu_out=relu_out, want_y=True, want_split=True)
        assert rel_err(y, reference(x, w, b, res, relu_in, relu_out)) <= 8e-6, (n, k, o, relu_in, relu_out)
        assert torch.equal(pair[0] + pair[1], y)
        y2, _ = K.linear_tf32x3(K.split_tf32(x.contiguous(), relu=relu_in), wp, b, residual=res, relu_out=relu_out)
        assert torch.equal(y, y2)
        assert torch.equal(wide[:, :k], x)                         # the in-place split happens in shared memory only


@torch.no_grad()
def test_linear_tf32x3_unsupported_shapes_are_rejected(cuda_device):
    assert not K.tf32x3_supported(3, 3, 3)
    x = torch.randn(10, 6, device=cuda_device)
    w = torch.randn(4, 6, device=cuda_device)
    with pytest.raises(RuntimeError):
        K.linear_tf32x3(K.split_tf32(x), K.split_tf32(w))
    # ... and the FFMA kernel takes them
    assert rel_err(K.linear(x, w), x.double() @ w.double().t()) <= 8e-6
