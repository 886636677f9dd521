"""StandardNormal base density (reference nflows/distributions/normal.py:11-50)."""
import numpy as np
import torch

from .. import kernels as K
from ..utils import torchutils
from .base import Distribution


class StandardNormal(Distribution):
    """Zero-mean, identity-covariance Gaussian over events of shape `shape`."""

    def __init__(self, shape):
        super().__init__()
        self._shape = torch.Size(shape)
        self.register_buffer("_log_z", torch.tensor(0.5 * np.prod(shape) * np.log(2 * np.pi), dtype=torch.float64),
                             persistent=False)
        self._log_z_host = float(0.5 * np.prod(shape) * np.log(2 * np.pi))

    def _check_shape(self, inputs):
        if inputs.shape[1:] != self._shape:
            raise ValueError("Expected input of shape {}, got {}".format(self._shape, inputs.shape[1:]))

    def _log_prob(self, inputs, context):
        return self._log_prob_plus(inputs, None)

    def _log_prob_plus(self, inputs, logabsdet):
        """log N(inputs) (+ logabsdet): with a native tensor this is one fused reduction kernel, which is how
        Flow._log_prob finishes (flows/base.py:49 in the reference adds the two afterwards)."""
        self._check_shape(inputs)
        if K.native_ok(inputs) and (logabsdet is None or K.native_ok(logabsdet)):
            flat = inputs.reshape(inputs.shape[0], int(np.prod(self._shape)))
            if flat.stride(-1) != 1:
                flat = flat.contiguous()
            with K.on_device_of(flat):
                return K.std_normal_log_prob(flat, self._log_z_host, logabsdet)
        neg_energy = -0.5 * torchutils.sum_except_batch(inputs ** 2, num_batch_dims=1)
        log_prob = neg_energy - self._log_z
        return log_prob if logabsdet is None else log_prob + logabsdet

    def _sample(self, num_samples, context):
        if context is None:
            return torch.randn(num_samples, *self._shape, device=self._log_z.device)
        context_size = context.shape[0]
        samples = torch.randn(context_size * num_samples, *self._shape, device=context.device)
        return torchutils.split_leading_dim(samples, [context_size, num_samples])

    def _mean(self, context):
        if context is None:
            return self._log_z.new_zeros(self._shape)
        return context.new_zeros(context.shape[0], *self._shape)


class ConditionalDiagonalNormal(Distribution):
    """Diagonal Gaussian whose means / log-stds are produced from the context by `context_encoder` (reference
    normal.py:53-132; the conditional base density of SURVEY.md section 8 row f4).  Torch path."""

    def __init__(self, shape, context_encoder=None):
        super().__init__()
        self._shape = torch.Size(shape)
        self._context_encoder = (lambda x: x) if context_encoder is None else context_encoder
        self.register_buffer("_log_z", torch.tensor(0.5 * np.prod(shape) * np.log(2 * np.pi), dtype=torch.float64),
                             persistent=False)

    def _compute_params(self, context):
        if context is None:
            raise ValueError("Context can't be None.")
        params = self._context_encoder(context)
        if params.shape[-1] % 2 != 0:
            raise RuntimeError("The context encoder must return a tensor whose last dimension is even.")
        if params.shape[0] != context.shape[0]:
            raise RuntimeError("The batch dimension of the parameters is inconsistent with the input.")
        half = params.shape[-1] // 2
        means = params[..., :half].reshape(params.shape[0], *self._shape)
        log_stds = params[..., half:].reshape(params.shape[0], *self._shape)
        return means, log_stds

    def _log_prob(self, inputs, context):
        if inputs.shape[1:] != self._shape:
            raise ValueError("Expected input of shape {}, got {}".format(self._shape, inputs.shape[1:]))
        means, log_stds = self._compute_params(context)
        assert means.shape == inputs.shape and log_stds.shape == inputs.shape
        z = (inputs - means) * torch.exp(-log_stds)
        log_prob = -0.5 * torchutils.sum_except_batch(z ** 2, num_batch_dims=1)
        log_prob -= torchutils.sum_except_batch(log_stds, num_batch_dims=1)
        log_prob -= self._log_z
        return log_prob

    def _sample(self, num_samples, context):
        means, log_stds = self._compute_params(context)
        means = torchutils.repeat_rows(means, num_samples)
        stds = torchutils.repeat_rows(torch.exp(log_stds), num_samples)
        noise = torch.randn(context.shape[0] * num_samples, *self._shape, device=means.device)
        return torchutils.split_leading_dim(means + stds * noise, [context.shape[0], num_samples])

    def _mean(self, context):
        return self._compute_params(context)[0]


class DiagonalNormal(Distribution):
    """Diagonal Gaussian with trainable mean and log-std (reference normal.py:135-180).  Torch path."""

    def __init__(self, shape):
        super().__init__()
        self._shape = torch.Size(shape)
        self.mean_ = torch.nn.Parameter(torch.zeros(shape).reshape(1, -1))
        self.log_std_ = torch.nn.Parameter(torch.zeros(shape).reshape(1, -1))
        self.register_buffer("_log_z", torch.tensor(0.5 * np.prod(shape) * np.log(2 * np.pi), dtype=torch.float64),
                             persistent=False)

    def _log_prob(self, inputs, context):
        if inputs.shape[1:] != self._shape:
            raise ValueError("Expected input of shape {}, got {}".format(self._shape, inputs.shape[1:]))
        z = (inputs - self.mean_) * torch.exp(-self.log_std_)
        log_prob = -0.5 * torchutils.sum_except_batch(z ** 2, num_batch_dims=1)
        log_prob -= torchutils.sum_except_batch(self.log_std_, num_batch_dims=1)
        log_prob -= self._log_z
        return log_prob

    def _sample(self, num_samples, context):
        raise NotImplementedError()

    def _mean(self, context):
        return self.mean
