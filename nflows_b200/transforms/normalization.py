"""ActNorm (reference nflows/transforms/normalization.py:144-218)."""
import torch
from torch import nn

from .. import kernels as K
from ..utils import typechecks as check
from .base import Transform, params_frozen


class ActNorm(Transform):
    """Per-feature (2-D inputs) or per-channel (NCHW inputs) affine normalisation, y = exp(log_scale) * x + shift,
    with the data-dependent initialisation of Glow on the first training batch."""

    def __init__(self, features):
        if not check.is_positive_int(features):
            raise TypeError("Number of features must be a positive integer.")
        super().__init__()
        self.register_buffer("initialized", torch.tensor(False, dtype=torch.bool))
        self.log_scale = nn.Parameter(torch.zeros(features))
        self.shift = nn.Parameter(torch.zeros(features))

    @property
    def scale(self):
        return torch.exp(self.log_scale)

    def _needs_init(self):
        return self.training and not bool(self.initialized)

    def _native_ready(self, inputs, context):
        return K.native_ok(inputs) and inputs.dim() == 2 and params_frozen(self) and not self._needs_init()

    def _native_apply(self, inputs, lad, flags, inverse, context=None):
        total = float(torch.sum(self.log_scale.detach()))
        return K.actnorm(inputs, self.scale.detach().contiguous(), self.shift.detach().contiguous(), lad,
                         -total if inverse else total, inverse)

    def _eager(self, inputs, context, inverse):
        if inputs.dim() not in (2, 4):
            raise ValueError("Expecting inputs to be a 2D or a 4D tensor.")
        if not inverse and self._needs_init():
            self._initialize(inputs)
        view = (1, -1, 1, 1) if inputs.dim() == 4 else (1, -1)
        scale, shift = self.scale.view(view), self.shift.view(view)
        sites = inputs.shape[2] * inputs.shape[3] if inputs.dim() == 4 else 1
        total = sites * torch.sum(self.log_scale)
        if inverse:
            outputs = (inputs - shift) / scale
            return outputs, -total * outputs.new_ones(inputs.shape[0])
        outputs = scale * inputs + shift
        return outputs, total * outputs.new_ones(inputs.shape[0])

    def _initialize(self, inputs):
        """Sets log_scale/shift so the outputs of this batch have zero mean and unit variance per feature
        (reference :206-218: std over the batch, mean of inputs/std)."""
        if inputs.dim() == 4:
            inputs = inputs.permute(0, 2, 3, 1).reshape(-1, inputs.shape[1])
        with torch.no_grad():
            std = inputs.std(dim=0)
            mu = (inputs / std).mean(dim=0)
            self.log_scale.data = -torch.log(std)
            self.shift.data = -mu
            self.initialized.data = torch.tensor(True, dtype=torch.bool)
