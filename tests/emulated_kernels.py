"""CPU stand-ins for the Python kernel wrappers of nflows_b200.kernels, for tests of the HOST logic of the native chain
(column layouts, row blocking, context / image chains) where there is no GPU.

Each stand-in restates the CONTRACT of the wrapper it replaces (include/nfk.h) with torch CPU ops, including the fp16 split-pair
operand format, so a flow can be pushed through `_native_apply` on CPU tensors.  Test infrastructure only: nothing in the
package imports this, and the kernels themselves are checked on hardware by the `-m gpu` tests."""
import contextlib
import math

import torch
import torch.nn.functional as F

from nflows_b200 import kernels as K
from nflows_b200.transforms import splines


def _pair(x, exp, relu=False, out=None):
    v = (F.relu(x) if relu else x).double() * 2.0 ** exp
    hi = v.to(torch.float16)
    lo = (v - hi.double()).to(torch.float16)
    if out is None:
        return K.Pair16(hi, lo, exp)
    out.hi.copy_(hi)
    out.lo.copy_(lo)
    return out


def _value(pair):
    return ((pair.hi.double() + pair.lo.double()) * 2.0 ** -pair.exp)


def _spline(desc, x, params, inverse):
    """x: [n, d_t]; params: [n, d_t, M] with the reference's (widths, heights, derivatives) order."""
    k = desc.num_bins
    w, h, d = params[..., :k] / desc.wh_divisor, params[..., k:2 * k] / desc.wh_divisor, params[..., 2 * k:]
    common = dict(inputs=x, unnormalized_widths=w, unnormalized_heights=h, unnormalized_derivatives=d, inverse=bool(inverse),
                  min_bin_width=desc.min_bin_width, min_bin_height=desc.min_bin_height, min_derivative=desc.min_derivative)
    if desc.linear_tails:
        return splines.unconstrained_rational_quadratic_spline(tails="linear", tail_bound=desc.right, **common)
    return splines.rational_quadratic_spline(left=desc.left, right=desc.right, bottom=desc.bottom, top=desc.top, **common)


def _cols(t_cols, width):
    if isinstance(t_cols, tuple):
        return torch.arange(t_cols[0], t_cols[0] + t_cols[1])
    return t_cols.long()


def install(monkeypatch):
    calls = {}

    def count(name):
        calls[name] = calls.get(name, 0) + 1

    def native_ok(t, context=None):
        return t.dtype == torch.float32 and not (torch.is_grad_enabled() and t.requires_grad)

    def linear(x, weight, bias=None, residual=None, relu_in=False, relu_out=False, out=None):
        count("linear")
        y = F.linear(F.relu(x) if relu_in else x, weight, bias)
        if relu_out:
            y = F.relu(y)
        if residual is not None:
            y = y + residual
        if out is not None:
            out.copy_(y)
            return out
        return y

    def gather_cols(x, cols, out=None):
        count("gather_cols")
        y = x[:, cols.long()]
        if out is not None:
            out.copy_(y)
            return out
        return y.contiguous()

    def actnorm(x, scale, shift, lad_accum, lad_const, inverse):
        count("actnorm")
        if lad_accum is not None:
            lad_accum += lad_const
        return (x - shift) / scale if inverse else x * scale + shift

    def rqs_rows(desc, inverse, x, params, t_cols, id_cols, lad_accum, flags, out=None):
        count("rqs_rows")
        y = torch.empty_like(x) if out is None else out
        t = t_cols.long()
        m = params.shape[1] // t.numel()
        yt, lad = _spline(desc, x[:, t], params.reshape(x.shape[0], t.numel(), m), inverse)
        y[:, id_cols.long()] = x[:, id_cols.long()]
        y[:, t] = yt
        lad_accum += lad.sum(dim=1)
        return y

    def std_normal_log_prob(z, log_z, lad=None):
        lp = -0.5 * (z * z).sum(dim=1) - log_z
        return lp + lad if lad is not None else lp

    def run_with_activation_rescale(fn):
        out, lad, _ = fn()
        return out, lad

    def weight_exp(w):
        amax = float(w.detach().abs().max())
        if not (amax > 0.0) or not math.isfinite(amax):
            return 0
        return max(-40, min(40, 14 - math.ceil(math.log2(amax))))

    def split_f16(x, exp, relu=False, out=None, flags=None):
        count("split_f16")
        if out is not None and out.exp != exp:
            raise ValueError("exponent mismatch")
        return _pair(x, exp, relu, out)

    def glu_skip(t, gate, skip=None, want_y=True, want_split=False, split_relu=False, split_exp=None, pair_out=None, flags=None):
        count("glu_skip")
        from nflows_b200 import config
        v = t * torch.sigmoid(gate)
        if skip is not None:
            v = v + skip
        pair = _pair(v, config.activation_exp if split_exp is None else split_exp, split_relu, pair_out) if want_split else None
        return (v if want_y else None), pair

    def nchw_to_rows(x):
        count("nchw_to_rows")
        b, c, h, w = x.shape
        return x.permute(0, 2, 3, 1).reshape(b * h * w, c).contiguous()

    def rows_to_nchw(rows, b, c, h, w):
        count("rows_to_nchw")
        return rows.reshape(b, h, w, c).permute(0, 3, 1, 2).contiguous()

    def squeeze_rows(rows, b, c, h, w, inverse=False):
        count("squeeze_rows")
        from nflows_b200.transforms.reshape import SqueezeTransform
        img = rows.reshape(b, h, w, c).permute(0, 3, 1, 2)
        out = (SqueezeTransform().inverse(img) if inverse else SqueezeTransform()(img))[0]
        return out.permute(0, 2, 3, 1).reshape(-1, out.shape[1]).contiguous(), tuple(out.shape[1:])

    def im2col3x3(pair, n_images, h, w):
        count("im2col3x3")
        n, c = pair.shape
        assert n == n_images * h * w

        def one(t):
            img = F.pad(t.reshape(n_images, h, w, c), (0, 0, 1, 1, 1, 1))
            taps = [img[:, ky:ky + h, kx:kx + w, :] for ky in range(3) for kx in range(3)]
            return torch.cat(taps, dim=-1).reshape(n, 9 * c).contiguous()
        return K.Pair16(one(pair.hi), one(pair.lo), pair.exp)

    def segment_sum_(values, out_accum, segment_len):
        count("segment_sum")
        out_accum += values.reshape(out_accum.numel(), segment_len).sum(dim=1)
        return out_accum

    def f16x3_supported(lda, ldw, k):
        return k >= 8 and k % 8 == 0 and lda % 8 == 0 and ldw % 8 == 0

    def linear_f16x3(a, w, bias=None, residual=None, relu_out=False, want_y=True, want_split=False, split_relu=False,
                     split_exp=None, split_cols=0, y_out=None, pair_out=None, flags=None, y_first_col=0):
        count("linear_f16x3")
        from nflows_b200 import config
        assert f16x3_supported(a.hi.stride(0), w.hi.stride(0), a.shape[1]) and a.shape[1] == w.shape[1], (a.shape, w.shape)
        v = _value(a) @ _value(w).t()
        if bias is not None:
            v = v + bias.double()
        if relu_out:
            v = F.relu(v)
        if residual is not None:
            v = v + residual.double()
        v = v.float()
        y = None
        if want_y:
            y = y_out if y_out is not None else torch.empty_like(v)
            y[:, y_first_col:] = v[:, y_first_col:]
            if y_first_col:
                y[:, :y_first_col] = float("nan")        # the kernel leaves these unwritten: nobody may read them
        pair = None
        if want_split:
            exp = config.activation_exp if split_exp is None else split_exp
            cols = split_cols or v.shape[1]
            pair = pair_out if pair_out is not None else K.Pair16.empty(v.shape[0], v.shape[1], exp, v.device)
            _pair(v[:, :cols], exp, split_relu, pair.cols(0, cols))
        return y, pair

    def rq_coupling_final_supported(num_bins, tails, hidden, lda):
        return num_bins in (4, 8, 10, 16) and hidden >= 8 and hidden % 8 == 0 and lda % 8 == 0

    def rq_coupling_final_padded_params(num_bins, tails):
        m = 3 * num_bins - 1 if tails == "linear" else 3 * num_bins + 1
        return (m + 7) // 8 * 8

    def rq_coupling_final(desc, inverse, a, wp, bias_packed, x, t_cols, y, lad_accum, flags, y_pair=None):
        count("rq_coupling_final")
        t = _cols(t_cols, x.shape[1])
        d_t = t.numel()
        mp = wp.shape[0] // d_t
        m = 3 * desc.num_bins - 1 if desc.linear_tails else 3 * desc.num_bins + 1
        params = (_value(a) @ _value(wp).t() + bias_packed.double()).float().reshape(x.shape[0], d_t, mp)[:, :, :m]
        yt, lad = _spline(desc, x[:, t], params, inverse)
        lad_accum += lad.sum(dim=1)
        if y_pair is not None:
            _pair(yt, y_pair.exp, False, K.Pair16(y_pair.hi[:, t], y_pair.lo[:, t], y_pair.exp))
            y_pair.hi[:, t], y_pair.lo[:, t] = _pair(yt, y_pair.exp).hi, _pair(yt, y_pair.exp).lo
            return None
        y[:, t] = yt
        return y

    def rq_coupling_step_supported(num_bins, tails, hidden, in_features, num_square_layers):
        # the predicate of nfk_rq_coupling_step_supported (csrc/nfk_coupling_step_tc.cu)
        return (num_bins in (4, 8, 10, 16) and 32 <= hidden <= 256 and hidden % 32 == 0 and in_features >= 8 and in_features % 8 == 0
                and 0 <= num_square_layers < 9)

    def rq_coupling_step(plan, a, desc=None, inverse=False, wp=None, bias_packed=None, x=None, t_cols=None, y=None, lad_accum=None,
                         flags=None, y_pair=None, h_pair=None):
        """The layer recursion of include/nfk.h (nfk_rq_coupling_step_f16x3) on the operands a dense.StepPlan packs: flag bit 0
        relu on (acc + bias), bit 1 add the current skip tensor, bit 2 the fp32 result becomes the skip tensor, bit 3 the next
        consumer sees relu(.); every hidden activation goes through the fp16 pair at the plan's exponent."""
        count("rq_coupling_step" if h_pair is None else "trunk_step")
        hdim = plan.hidden
        cur, skip = _value(a), None
        for l, f in enumerate(plan.layer_flags):
            if l == 0:
                w = _value(plan.w0)
            else:
                blk = slice((l - 1) * hdim, l * hdim)
                w = _value(K.Pair16(plan.wt_hi[blk], plan.wt_lo[blk], int(plan.wt_exps_c[l - 1])))
            v = cur @ w.t() + plan.bias[l * hdim:(l + 1) * hdim].double()
            if f & 1:
                v = F.relu(v)
            if f & 2:
                v = v + skip
            v = v.float().double()                      # the kernel's sums are fp32
            if f & 4:
                skip = v
            cur = _value(_pair(v.float(), plan.act_exp, relu=bool(f & 8)))
        if h_pair is not None:
            _pair(cur.float(), plan.act_exp, False, h_pair)
            return None
        return rq_coupling_final(desc, inverse, _pair(cur.float(), plan.act_exp), wp, bias_packed, x, t_cols, y, lad_accum, flags,
                                 y_pair=y_pair)

    def affine_coupling_rows(x, params, mult, scale_activation, inverse, t_cols, id_cols, lad_accum, out=None):
        count("affine_coupling_rows")
        y = torch.empty_like(x) if out is None else out
        t, d_t = t_cols.long(), t_cols.numel()
        y[:, id_cols.long()] = x[:, id_cols.long()]
        shift = params[:, :d_t]
        if mult == 2:
            raw = params[:, d_t:]
            scale = torch.sigmoid(raw + 2) + 1e-3 if scale_activation == 0 else (F.softplus(raw) + 1e-3).clamp(0, 3)
            y[:, t] = (x[:, t] - shift) / scale if inverse else x[:, t] * scale + shift
            lad_accum += -torch.log(scale).sum(dim=1) if inverse else torch.log(scale).sum(dim=1)
        else:
            y[:, t] = x[:, t] - shift if inverse else x[:, t] + shift
        return y

    def affine_coupling_final(a, w, bias, x, t_cols, mult, scale_activation, inverse, y, lad_accum, flags=None):
        count("affine_coupling_final")
        t = _cols(t_cols, x.shape[1])
        params = (_value(a) @ _value(w).t() + bias.double()).float()
        xt = x[:, t]
        if mult == 2:
            shift, raw = params[:, 0::2], params[:, 1::2]
            scale = torch.sigmoid(raw + 2) + 1e-3 if scale_activation == 0 else (F.softplus(raw) + 1e-3).clamp(0, 3)
            lad = torch.log(scale).sum(dim=1)
            y[:, t] = (xt - shift) / scale if inverse else xt * scale + shift
            if lad_accum is not None:
                lad_accum += -lad if inverse else lad
        else:
            y[:, t] = xt - params if inverse else xt + params
        return y

    # the functional spline API keeps its torch formulation (it is what the stand-ins of the spline kernels call)
    monkeypatch.setattr(splines.rational_quadratic, "_use_native", lambda inputs, *params: False)
    for name, fn in dict(
            native_ok=native_ok, on_device_of=lambda t: contextlib.nullcontext(), warn_eager_cuda=lambda *a, **k: None,
            new_flags=lambda device: torch.zeros(1, dtype=torch.int32), index_tensor=lambda idx, device: idx.to(torch.int32),
            fill_=lambda t, v: t.fill_(v), add_const_=lambda lad, c: lad.add_(c),
            zeros_lad=lambda x: torch.zeros(x.shape[0]), linear=linear, gather_cols=gather_cols, actnorm=actnorm, rqs_rows=rqs_rows,
            std_normal_log_prob=std_normal_log_prob, raise_for_flags=lambda flags: None,
            run_with_activation_rescale=run_with_activation_rescale, weight_exp=weight_exp, split_f16=split_f16, glu_skip=glu_skip,
            nchw_to_rows=nchw_to_rows, rows_to_nchw=rows_to_nchw, squeeze_rows=squeeze_rows, im2col3x3=im2col3x3,
            segment_sum_=segment_sum_, f16x3_supported=f16x3_supported, linear_f16x3=linear_f16x3,
            rq_coupling_final_supported=rq_coupling_final_supported, rq_coupling_final_padded_params=rq_coupling_final_padded_params,
            rq_coupling_final=rq_coupling_final, affine_coupling_final=affine_coupling_final, affine_coupling_rows=affine_coupling_rows,
            rq_coupling_step_supported=rq_coupling_step_supported, rq_coupling_step=rq_coupling_step).items():
        monkeypatch.setattr(K, name, fn)
    return calls
