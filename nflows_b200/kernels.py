"""Tensor-level wrappers over the C ABI (include/nfk.h).  Each function enqueues exactly the kernels the
C entry point launches on the current CUDA stream; tensors are only used for their storage."""
import ctypes

import torch

from . import _native as N


#: bench.py sets this to a list to collect (tag, rows, start_event, end_event) around tagged launches
TIMELINE = None


class timed:
    """Brackets a launch with CUDA events on the current stream when bench.py asked for a timeline."""

    def __init__(self, tag, rows):
        self.tag, self.rows = tag, rows

    def __enter__(self):
        if TIMELINE is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *exc):
        if TIMELINE is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            TIMELINE.append((self.tag, self.rows, self.e0, e1))


def _rows2d(t, name):
    if not (t.is_cuda and t.dtype == torch.float32 and t.dim() == 2 and t.stride(1) == 1):
        raise ValueError("{} must be a 2-D float32 CUDA tensor with unit column stride".format(name))
    return t


def native_ok(t):
    """True when `t` is something the native path takes: CUDA, fp32, no autograd graph to build."""
    return t.is_cuda and t.dtype == torch.float32 and not (torch.is_grad_enabled() and t.requires_grad)


def new_flags(device):
    return torch.zeros(1, dtype=torch.int32, device=device)


def index_tensor(idx, device):
    return idx.to(device=device, dtype=torch.int32).contiguous()


def fill_(t, value):
    N.check(N.lib().nfk_fill(t.data_ptr(), float(value), t.numel(), N.stream()))
    return t


def add_const_(lad, c):
    N.check(N.lib().nfk_add_const(lad.data_ptr(), float(c), lad.numel(), N.stream()))
    return lad


def zeros_lad(x):
    return fill_(torch.empty(x.shape[0], dtype=torch.float32, device=x.device), 0.0)


def linear(x, weight, bias=None, residual=None, relu_in=False, relu_out=False, out=None):
    """out = post(pre(x) @ weight.T + bias) + residual; weight is [out_features, in_features] (nn.Linear layout)."""
    _rows2d(x, "x")
    _rows2d(weight, "weight")
    n, k = x.shape
    o = weight.shape[0]
    if weight.shape[1] != k:
        raise ValueError("weight is {}x{}, x has {} features".format(o, weight.shape[1], k))
    if out is None:
        out = torch.empty(n, o, dtype=torch.float32, device=x.device)
    _rows2d(out, "out")
    if bias is not None and not bias.is_contiguous():
        bias = bias.contiguous()
    N.check(N.lib().nfk_linear(x.data_ptr(), x.stride(0), weight.data_ptr(), weight.stride(0), N.ptr(bias),
                               N.ptr(residual), residual.stride(0) if residual is not None else 0, out.data_ptr(),
                               out.stride(0), n, k, o, int(relu_in), int(relu_out), N.stream()))
    return out


def gather_cols(x, cols_i32, out=None):
    _rows2d(x, "x")
    n = x.shape[0]
    c = cols_i32.numel()
    if out is None:
        out = torch.empty(n, c, dtype=torch.float32, device=x.device)
    N.check(N.lib().nfk_gather_cols(x.data_ptr(), x.stride(0), cols_i32.data_ptr(), c, out.data_ptr(), out.stride(0), n,
                                    N.stream()))
    return out


def actnorm(x, scale, shift, lad_accum, lad_const, inverse):
    _rows2d(x, "x")
    y = torch.empty_like(x, memory_format=torch.contiguous_format)
    N.check(N.lib().nfk_actnorm(x.data_ptr(), x.stride(0), scale.data_ptr(), shift.data_ptr(), y.data_ptr(), y.stride(0),
                                N.ptr(lad_accum), float(lad_const), x.shape[0], x.shape[1], int(inverse), N.stream()))
    return y


def rqs_rows(desc, inverse, x, params, t_cols, id_cols, lad_accum, flags, out=None):
    _rows2d(x, "x")
    _rows2d(params, "params")
    if not params.is_contiguous():
        raise ValueError("params must be contiguous")
    y = torch.empty_like(x, memory_format=torch.contiguous_format) if out is None else out
    N.check(N.lib().nfk_rqs_rows(ctypes.byref(desc), int(inverse), x.data_ptr(), x.stride(0), params.data_ptr(),
                                 N.ptr(t_cols), t_cols.numel(), N.ptr(id_cols), id_cols.numel(), y.data_ptr(), y.stride(0),
                                 N.ptr(lad_accum), x.shape[0], N.ptr(flags), N.stream()))
    return y


def rqs_elementwise(desc, inverse, x, uw, uh, ud, param_period=0, flags=None):
    """x: any shape (contiguous); uw/uh/ud: [..., K], [..., K], [..., K-1|K+1] contiguous on the last dim."""
    x = x.contiguous()
    uw, uh, ud = uw.contiguous(), uh.contiguous(), ud.contiguous()
    y = torch.empty_like(x)
    lad = torch.empty_like(x)
    N.check(N.lib().nfk_rqs_elementwise(ctypes.byref(desc), int(inverse), x.data_ptr(), uw.data_ptr(), uh.data_ptr(),
                                        ud.data_ptr(), uw.shape[-1], uh.shape[-1], ud.shape[-1], int(param_period),
                                        y.data_ptr(), lad.data_ptr(), x.numel(), N.ptr(flags), N.stream()))
    return y, lad


def affine_coupling_rows(x, params, mult, scale_activation, inverse, t_cols, id_cols, lad_accum, out=None):
    _rows2d(x, "x")
    if not params.is_contiguous():
        raise ValueError("params must be contiguous")
    y = torch.empty_like(x, memory_format=torch.contiguous_format) if out is None else out
    N.check(N.lib().nfk_affine_coupling_rows(x.data_ptr(), x.stride(0), params.data_ptr(), int(mult), int(scale_activation),
                                             int(inverse), N.ptr(t_cols), t_cols.numel(), N.ptr(id_cols), id_cols.numel(),
                                             y.data_ptr(), y.stride(0), N.ptr(lad_accum), x.shape[0], N.stream()))
    return y


def std_normal_log_prob(z, log_z, lad=None):
    _rows2d(z, "z")
    out = torch.empty(z.shape[0], dtype=torch.float32, device=z.device)
    N.check(N.lib().nfk_std_normal_log_prob(z.data_ptr(), z.stride(0), z.shape[1], float(log_z), N.ptr(lad), out.data_ptr(),
                                            z.shape[0], N.stream()))
    return out


def raise_for_flags(flags):
    """Mirror the reference's exceptions (rational_quadratic.py:81-82, :142).  One device->host read."""
    from .transforms.base import InputOutsideDomain
    v = int(flags.item())
    if v & 1:
        raise InputOutsideDomain()
    if v & 2:
        raise AssertionError("rational-quadratic spline inverse: negative discriminant")


# ---- split-TF32 tensor-core dense layers ------------------------------------------------------------------------------
def split_tf32(x, cols_i32=None, relu=False, copy_to=None):
    """(hi, lo) operand pair of pre(x[:, cols]) for `linear_tf32x3`; optionally copies x[:, cols] into copy_to[:, cols]."""
    _rows2d(x, "x")
    n = x.shape[0]
    c = x.shape[1] if cols_i32 is None else cols_i32.numel()
    hi = torch.empty(n, c, dtype=torch.float32, device=x.device)
    lo = torch.empty_like(hi)
    if TIMELINE is not None:
        with timed("split_%d" % c, n):
            N.check(N.lib().nfk_split_tf32(x.data_ptr(), x.stride(0), N.ptr(cols_i32), c, int(relu), hi.data_ptr(), lo.data_ptr(),
                                           hi.stride(0), N.ptr(copy_to), copy_to.stride(0) if copy_to is not None else 0, n,
                                           N.stream()))
        return hi, lo
    N.check(N.lib().nfk_split_tf32(x.data_ptr(), x.stride(0), N.ptr(cols_i32), c, int(relu), hi.data_ptr(), lo.data_ptr(),
                                   hi.stride(0), N.ptr(copy_to), copy_to.stride(0) if copy_to is not None else 0, n,
                                   N.stream()))
    return hi, lo


def tf32x3_supported(lda, ldw, in_features):
    return bool(N.load().nfk_linear_tf32x3_supported(int(lda), int(ldw), int(in_features)))


def linear_tf32x3(a, w_pair, bias=None, residual=None, relu_in=False, relu_out=False, want_y=True, want_split=False,
                  split_relu=False, y_out=None, pair_out=None):
    """tcgen05 dense layer.  `a` is either the (hi, lo) split pair of the activations or the plain fp32 activation tensor,
    which the kernel splits on chip (after relu when relu_in).  Returns (y or None, (y_hi, y_lo) or None); y_out /
    pair_out are caller-provided destinations (row slices of larger buffers)."""
    w_hi, w_lo = w_pair
    raw = not isinstance(a, tuple)
    a_hi = _rows2d(a, "a") if raw else a[0]
    if not raw and relu_in:
        raise ValueError("relu_in needs the fp32 activation, not a split pair")
    n, k = a_hi.shape
    o = w_hi.shape[0]
    dev = a_hi.device
    y = (y_out if y_out is not None else torch.empty(n, o, dtype=torch.float32, device=dev)) if want_y else None
    if want_split:
        pair = pair_out if pair_out is not None else (torch.empty(n, o, dtype=torch.float32, device=dev),
                                                      torch.empty(n, o, dtype=torch.float32, device=dev))
    else:
        pair = None
    if bias is not None and not bias.is_contiguous():
        bias = bias.contiguous()
    tail = (w_hi.data_ptr(), w_lo.data_ptr(), w_hi.stride(0), N.ptr(bias),
            N.ptr(residual), residual.stride(0) if residual is not None else 0, N.ptr(y), y.stride(0) if y is not None else 0,
            N.ptr(pair[0]) if pair else 0, N.ptr(pair[1]) if pair else 0, pair[0].stride(0) if pair else 0, int(relu_out),
            int(split_relu), n, k, o, N.stream())
    with timed("linear_%dx%d" % (k, o), n):
        if raw:
            N.check(N.lib().nfk_linear_tf32x3_a32(a_hi.data_ptr(), a_hi.stride(0), int(relu_in), *tail))
        else:
            N.check(N.lib().nfk_linear_tf32x3(a_hi.data_ptr(), a[1].data_ptr(), a_hi.stride(0), *tail))
    return y, pair


def rq_coupling_final_supported(num_bins, tails, hidden, lda):
    return bool(N.load().nfk_rq_coupling_final_supported(int(num_bins), 1 if tails == "linear" else 0, int(hidden), int(lda)))


def rq_coupling_final_padded_params(num_bins, tails):
    return int(N.load().nfk_rq_coupling_final_padded_params(int(num_bins), 1 if tails == "linear" else 0))


def rq_coupling_final(desc, inverse, a, wp_pair, bias_packed, x, t_cols, y, lad_accum, flags, relu_in=False):
    """Fused final conditioner layer + RQ spline + scatter + log|det| (one tcgen05 kernel).  `a`: the (hi, lo) pair of the
    hidden activation, or the fp32 activation itself (split on chip, after relu when relu_in).  y may be x."""
    tail = (wp_pair[0].data_ptr(), wp_pair[1].data_ptr(), wp_pair[0].stride(0), bias_packed.data_ptr())
    rest = (x.data_ptr(), x.stride(0), t_cols.data_ptr(), t_cols.numel(), y.data_ptr(), y.stride(0), N.ptr(lad_accum),
            x.shape[0], N.ptr(flags), N.stream())
    if isinstance(a, tuple):
        if relu_in:
            raise ValueError("relu_in needs the fp32 activation, not a split pair")
        a_hi, a_lo = a
        N.check(N.lib().nfk_rq_coupling_final_tf32x3(ctypes.byref(desc), int(inverse), a_hi.data_ptr(), a_lo.data_ptr(),
                                                     a_hi.stride(0), *tail, a_hi.shape[1], *rest))
    else:
        _rows2d(a, "a")
        N.check(N.lib().nfk_rq_coupling_final_tf32x3_a32(ctypes.byref(desc), int(inverse), a.data_ptr(), a.stride(0),
                                                         int(relu_in), *tail, a.shape[1], *rest))
    return y
