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
