"""Masked autoregressive transforms (reference nflows/transforms/autoregressive.py:24-62, 65-116, 404-495).

Forward is one MADE pass + an elementwise map: ONE launch of the coupling-step kernel (masked conditioner + spline).  The
inverse is inherently sequential: feature i of the output needs the conditioner evaluated on outputs 1..i-1, so -- like the
reference (:43-52) -- it runs D passes; here pass i is one launch of the same kernel on the degree-sorted SUB-network feature
i can see (hidden units of degree <= i are a prefix once sorted) with the final layer of feature i alone."""
import numpy as np
import torch
from torch.nn import functional as F

from .. import _native as N
from .. import config
from .. import dense as D
from .. import kernels as K
from . import made as made_module
from . import splines
from .base import Transform, params_frozen


class AutoregressiveTransform(Transform):
    """outputs_i = f(inputs_i; params_i(inputs_<i)).  Subclasses define the elementwise map and its parameter count."""

    def __init__(self, autoregressive_net):
        super().__init__()
        self.autoregressive_net = autoregressive_net

    def _eager(self, inputs, context, inverse):
        if not inverse:
            return self._elementwise_forward(inputs, self.autoregressive_net(inputs, context))
        outputs = torch.zeros_like(inputs)
        logabsdet = None
        for _ in range(int(np.prod(inputs.shape[1:]))):
            outputs, logabsdet = self._elementwise_inverse(inputs, self.autoregressive_net(outputs, context))
        return outputs, logabsdet

    def _output_dim_multiplier(self):
        raise NotImplementedError()

    def _elementwise_forward(self, inputs, autoregressive_params):
        raise NotImplementedError()

    def _elementwise_inverse(self, inputs, autoregressive_params):
        raise NotImplementedError()


class MaskedAffineAutoregressiveTransform(AutoregressiveTransform):
    """MAF layer: y_i = scale_i * x_i + shift_i with (unconstrained scale, shift) = MADE(x)[i, :] (reference :65-116)."""

    def __init__(self, features, hidden_features, context_features=None, num_blocks=2, use_residual_blocks=True,
                 random_mask=False, activation=F.relu, dropout_probability=0.0, use_batch_norm=False):
        self.features = features
        net = made_module.MADE(features=features, hidden_features=hidden_features, context_features=context_features,
                               num_blocks=num_blocks, output_multiplier=self._output_dim_multiplier(),
                               use_residual_blocks=use_residual_blocks, random_mask=random_mask, activation=activation,
                               dropout_probability=dropout_probability, use_batch_norm=use_batch_norm)
        self._epsilon = 1e-3
        super().__init__(net)

    def _output_dim_multiplier(self):
        return 2

    def _scale_shift(self, params):
        params = params.view(-1, self.features, self._output_dim_multiplier())
        return torch.sigmoid(params[..., 0] + 2.0) + self._epsilon, params[..., 1]

    def _elementwise_forward(self, inputs, autoregressive_params):
        scale, shift = self._scale_shift(autoregressive_params)
        return scale * inputs + shift, torch.sum(torch.log(scale), dim=1)

    def _elementwise_inverse(self, inputs, autoregressive_params):
        scale, shift = self._scale_shift(autoregressive_params)
        return (inputs - shift) / scale, -torch.sum(torch.log(scale), dim=1)


class MaskedPiecewiseRationalQuadraticAutoregressiveTransform(AutoregressiveTransform):
    """Autoregressive neural spline layer: feature i goes through an RQ spline whose 3K-1 / 3K+1 parameters are rows
    i*M .. i*M+M-1 of the MADE output (feature-major, like the coupling layout)."""

    def __init__(self, features, hidden_features, context_features=None, num_bins=10, tails=None, tail_bound=1.0,
                 num_blocks=2, use_residual_blocks=True, random_mask=False, activation=F.relu, dropout_probability=0.0,
                 use_batch_norm=False, min_bin_width=splines.rational_quadratic.DEFAULT_MIN_BIN_WIDTH,
                 min_bin_height=splines.rational_quadratic.DEFAULT_MIN_BIN_HEIGHT,
                 min_derivative=splines.rational_quadratic.DEFAULT_MIN_DERIVATIVE):
        self.num_bins = num_bins
        self.min_bin_width = min_bin_width
        self.min_bin_height = min_bin_height
        self.min_derivative = min_derivative
        self.tails = tails
        self.tail_bound = tail_bound
        self.features = features
        net = made_module.MADE(features=features, hidden_features=hidden_features, context_features=context_features,
                               num_blocks=num_blocks, output_multiplier=self._output_dim_multiplier(),
                               use_residual_blocks=use_residual_blocks, random_mask=random_mask, activation=activation,
                               dropout_probability=dropout_probability, use_batch_norm=use_batch_norm)
        super().__init__(net)
        self._col_cache = None

    def _output_dim_multiplier(self):
        if self.tails == "linear":
            return self.num_bins * 3 - 1
        if self.tails is None:
            return self.num_bins * 3 + 1
        raise ValueError

    def _softmax_divisor(self):
        net = self.autoregressive_net
        return float(np.sqrt(net.hidden_features)) if hasattr(net, "hidden_features") else None

    def _elementwise(self, inputs, autoregressive_params, inverse=False):
        b, d = inputs.shape[0], inputs.shape[1]
        params = autoregressive_params.view(b, d, self._output_dim_multiplier())
        k = self.num_bins
        widths, heights, derivatives = params[..., :k], params[..., k:2 * k], params[..., 2 * k:]
        divisor = self._softmax_divisor()
        if divisor is not None:
            widths, heights = widths / divisor, heights / divisor
        common = dict(inputs=inputs, unnormalized_widths=widths, unnormalized_heights=heights,
                      unnormalized_derivatives=derivatives, inverse=inverse, min_bin_width=self.min_bin_width,
                      min_bin_height=self.min_bin_height, min_derivative=self.min_derivative)
        if self.tails is None:
            outputs, logabsdet = splines.rational_quadratic_spline(**common)
        elif self.tails == "linear":
            outputs, logabsdet = splines.unconstrained_rational_quadratic_spline(tails=self.tails, tail_bound=self.tail_bound,
                                                                               **common)
        else:
            raise ValueError
        return outputs, torch.sum(logabsdet.reshape(b, -1), dim=1)

    def _elementwise_forward(self, inputs, autoregressive_params):
        return self._elementwise(inputs, autoregressive_params)

    def _elementwise_inverse(self, inputs, autoregressive_params):
        return self._elementwise(inputs, autoregressive_params, inverse=True)

    # ---- native ----------------------------------------------------------------------------------------------------
    def _native_ready(self, inputs, context):
        return (K.native_ok(inputs, context) and inputs.dim() == 2 and context is None and params_frozen(self)
                and self.num_bins <= 64 and self.autoregressive_net.dense_chain(None) is not None)

    def _spline_desc(self):
        if self.min_bin_width * self.num_bins > 1.0:
            raise ValueError("Minimal bin width too large for the number of bins")
        if self.min_bin_height * self.num_bins > 1.0:
            raise ValueError("Minimal bin height too large for the number of bins")
        divisor = self._softmax_divisor()
        return N.spline_desc(self.num_bins, self.tails, self.tail_bound, 0.0, 1.0, 0.0, 1.0, self.min_bin_width,
                             self.min_bin_height, self.min_derivative, False, 1.0 if divisor is None else divisor)

    def _cols(self, device):
        if self._col_cache is None or self._col_cache[0] != str(device):
            self._col_cache = (str(device), torch.arange(self.features, device=device, dtype=torch.int32),
                               torch.zeros(0, device=device, dtype=torch.int32))
        return self._col_cache[1], self._col_cache[2]

    def _native_pass(self, conditioner_input, spline_input, lad, flags, inverse):
        """outputs[:, i] = spline_i(spline_input[:, i]; MADE(conditioner_input)[i]) for all i, one launch sequence."""
        chain = self.autoregressive_net.dense_chain(None)
        all_cols, no_cols = self._cols(spline_input.device)
        use_tc = D.chain_uses_tc(chain, self.features)
        outputs = torch.empty_like(spline_input, memory_format=torch.contiguous_format)
        desc = self._spline_desc()
        weight, bias = chain[-1][0], chain[-1][1]
        hidden = weight.shape[1]
        fused = (use_tc and config.fuse_coupling and bias is not None
                 and K.rq_coupling_final_supported(self.num_bins, self.tails, hidden, hidden))
        state = D.run_trunk(chain, conditioner_input, None, use_tc, flags=flags)
        if fused:
            m = self._output_dim_multiplier()
            mp = K.rq_coupling_final_padded_params(self.num_bins, self.tails)
            wp_pair, bias_packed = D.pack_final_spline(weight, bias, self.features, m, mp)
            K.rq_coupling_final(desc, inverse, state.pair, wp_pair, bias_packed, spline_input, (0, self.features), outputs, lad,
                                flags)
        else:
            params = D.run_last(chain, state, 0, spline_input.shape[0], use_tc, flags=flags)
            K.rqs_rows(desc, inverse, spline_input, params, all_cols, no_cols, lad, flags, out=outputs)
        return outputs

    # ---- the whole conditioner + spline as ONE launch per pass (nfk_rq_coupling_step_f16x3) ---------------------------------
    def _step_ready(self, chain):
        if not (config.coupling_step_kernel and config.fuse_coupling) or chain[-1][1] is None or self.features % 8:
            return False
        if D.plan_step_kernel(chain) is None or not D.chain_uses_tc(chain, self.features):
            return False
        hidden = chain[-1][0].shape[1]
        return K.rq_coupling_step_supported(self.num_bins, self.tails, hidden, self.features, len(chain) - 2)

    def _native_forward_step(self, chain, inputs, lad, flags):
        m, mp = self._output_dim_multiplier(), K.rq_coupling_final_padded_params(self.num_bins, self.tails)
        wp_pair, bias_packed = D.pack_final_spline(chain[-1][0], chain[-1][1], self.features, m, mp)
        outputs = torch.empty_like(inputs, memory_format=torch.contiguous_format)
        a_pair = K.split_f16(inputs, D.act_exp(), flags=flags)
        K.rq_coupling_step(D.step_plan(chain), a_pair, self._spline_desc(), False, wp_pair, bias_packed, inputs,
                           (0, self.features), outputs, lad, flags)
        return outputs

    def _sorted_subnets(self, chain):
        """Degree-sorted copies of the MADE weights for the inverse.  Feature i (degree i + 1) only sees hidden units of degree
        <= i; with the hidden units sorted by degree (one permutation for every hidden layer: the residual blocks keep degrees
        per index) those are a PREFIX, so pass i runs the sub-network of the first H_i units (rounded up to 32) and the final
        layer of feature i alone -- the total work of the D passes is ~1/8 of D full passes.  Cached per parameter version."""
        net = self.autoregressive_net
        key = tuple((l[0].data_ptr(), l[0]._version, l[1].data_ptr(), l[1]._version) for l in chain) + (D.act_exp(), D.cache_epoch())
        hit = getattr(self, "_subnet_cache", None)
        if hit is not None and hit[0] == key:
            return hit[1]
        deg = net.initial_layer.degrees.to(chain[0][0].device)
        for block in net.blocks:
            if not torch.equal(block.degrees.to(deg.device), deg):
                return None
        perm = torch.argsort(deg, stable=True)
        sorted_deg = deg[perm].cpu()
        hidden = deg.numel()
        body = []
        for li, (w, b, relu_in, relu_out, res) in enumerate(chain[:-1]):
            w = w.detach()
            w = w[perm] if li == 0 else w[perm][:, perm]
            body.append((w.contiguous(), b.detach()[perm].contiguous(), relu_in, relu_out, res))
        wf = chain[-1][0].detach()[:, perm].contiguous()
        m, mp = self._output_dim_multiplier(), K.rq_coupling_final_padded_params(self.num_bins, self.tails)
        wp_pair, bias_packed = D.pack_final_spline(wf, chain[-1][1].detach(), self.features, m, mp)
        flags_l = D.plan_step_kernel(body + [chain[-1]])
        plans, widths = {}, []
        for i in range(self.features):
            count = int((sorted_deg <= i).sum())
            h = min(hidden, max(32, (count + 31) // 32 * 32))
            widths.append(h)
            if h not in plans:
                sub = [((w[:h] if li == 0 else w[:h, :h]).contiguous(), b[:h].contiguous(), ri, ro, rs)
                       for li, (w, b, ri, ro, rs) in enumerate(body)]
                plans[h] = D.StepPlan(sub).set_flags(flags_l)
        out = (plans, widths, wp_pair, bias_packed, mp, (wf,))
        self._subnet_cache = (key, out)
        return out

    def _native_inverse_step(self, chain, inputs, lad, flags):
        sub = self._sorted_subnets(chain)
        if sub is None:
            return None
        plans, widths, wp_pair, bias_packed, mp, _ = sub
        n, d = inputs.shape
        desc = self._spline_desc()
        outputs = torch.zeros_like(inputs, memory_format=torch.contiguous_format)
        pair = K.Pair16(torch.zeros(n, d, dtype=torch.float16, device=inputs.device),
                        torch.zeros(n, d, dtype=torch.float16, device=inputs.device), D.act_exp())
        for i in range(d):
            h = widths[i]
            wp_i = K.Pair16(wp_pair.hi[i * mp:(i + 1) * mp, :h], wp_pair.lo[i * mp:(i + 1) * mp, :h], wp_pair.exp)
            K.rq_coupling_step(plans[h], pair, desc, True, wp_i, bias_packed[i * mp:(i + 1) * mp], inputs, (i, 1), outputs, lad,
                               flags)
            if i + 1 < d:
                K.split_f16(outputs[:, i:i + 1], pair.exp, out=pair.cols(i, i + 1), flags=flags)
        return outputs

    def _native_apply(self, inputs, lad, flags, inverse, context=None):
        if inputs.shape[1] != self.features:
            raise ValueError("Expected features = {}, got {}.".format(self.features, inputs.shape[1]))
        chain = self.autoregressive_net.dense_chain(None)
        step = self._step_ready(chain)
        if not inverse:
            if step:
                return self._native_forward_step(chain, inputs, lad, flags)
            return self._native_pass(inputs, inputs, lad, flags, False)
        if step:
            outputs = self._native_inverse_step(chain, inputs, lad, flags)
            if outputs is not None:
                return outputs
        outputs = K.fill_(torch.empty_like(inputs, memory_format=torch.contiguous_format), 0.0)
        for i in range(self.features):
            last = i == self.features - 1
            outputs = self._native_pass(outputs, inputs, lad if last else None, flags, True)
        return outputs
