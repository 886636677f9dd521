"""Trivial transforms kept for API compatibility (reference nflows/transforms/standard.py:12-90); plain torch."""
import torch

from .base import Transform


class IdentityTransform(Transform):
    def forward(self, inputs, context=None):
        return inputs, inputs.new_zeros(inputs.shape[0])

    def inverse(self, inputs, context=None):
        return self(inputs, context)


class PointwiseAffineTransform(Transform):
    """y = scale * x + shift with broadcastable scale/shift; log|det| = sum(log|scale|) over event dims."""

    def __init__(self, shift=0.0, scale=1.0):
        super().__init__()
        shift, scale = map(torch.as_tensor, (shift, scale))
        if (scale == 0.0).any():
            raise ValueError("Scale must be non-zero.")
        self.register_buffer("_shift", shift)
        self.register_buffer("_scale", scale)

    @property
    def _log_abs_scale(self):
        return torch.log(torch.abs(self._scale))

    def _batch_logabsdet(self, batch_shape):
        if self._log_abs_scale.numel() > 1:
            return self._log_abs_scale.expand(batch_shape).sum()
        return self._log_abs_scale * torch.Size(batch_shape).numel()

    def forward(self, inputs, context=None):
        batch_size, *batch_shape = inputs.size()
        outputs = inputs * self._scale + self._shift
        return outputs, self._batch_logabsdet(batch_shape).expand(batch_size)

    def inverse(self, inputs, context=None):
        batch_size, *batch_shape = inputs.size()
        outputs = (inputs - self._shift) / self._scale
        return outputs, -self._batch_logabsdet(batch_shape).expand(batch_size)


class AffineTransform(PointwiseAffineTransform):
    """Deprecated alias kept by the reference (standard.py:70-86); `None` means "default"."""

    def __init__(self, shift=0.0, scale=1.0):
        import warnings
        warnings.warn("Use PointwiseAffineTransform", DeprecationWarning)
        super().__init__(0.0 if shift is None else shift, 1.0 if scale is None else scale)


AffineScalarTransform = AffineTransform
