"""Masked autoregressive transforms (reference nflows/transforms/autoregressive.py:24-62, 65-116, 404-495).

Forward is one MADE pass + an elementwise map.  The inverse is inherently sequential: feature i of the output needs the
conditioner evaluated on outputs 1..i-1, so -- like the reference (:43-52) -- it runs D passes; here every pass is the
tensor-core dense chain + ONE fused final-layer/spline kernel (or the HBM spline kernel), instead of ~700 ATen launches."""
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

    def _native_apply(self, inputs, lad, flags, inverse, context=None):
        if inputs.shape[1] != self.features:
            raise ValueError("Expected features = {}, got {}.".format(self.features, inputs.shape[1]))
        if not inverse:
            return self._native_pass(inputs, inputs, lad, flags, False)
        outputs = K.fill_(torch.empty_like(inputs, memory_format=torch.contiguous_format), 0.0)
        for i in range(self.features):
            last = i == self.features - 1
            outputs = self._native_pass(outputs, inputs, lad if last else None, flags, True)
        return outputs
