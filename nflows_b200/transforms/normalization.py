"""ActNorm and BatchNorm transforms (reference nflows/transforms/normalization.py:72-218)."""
import numpy as np
import torch
from torch import nn
from torch.nn import functional as F

from .. import kernels as K
from ..utils import typechecks as check
from .base import InverseNotAvailable, Transform, params_frozen


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
        return K.native_ok(inputs, context) and inputs.dim() == 2 and params_frozen(self) and not self._needs_init()

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


class BatchNorm(Transform):
    """Batch normalisation as an invertible transform on feature vectors (reference :72-141): batch statistics in training
    mode (running statistics updated with `momentum`), running statistics in eval mode, where it is a fixed per-feature
    affine map and therefore invertible.  Torch path only: it is not on the hot path."""

    def __init__(self, features, eps=1e-5, momentum=0.1, affine=True):
        if not check.is_positive_int(features):
            raise TypeError("Number of features must be a positive integer.")
        super().__init__()
        self.momentum = momentum
        self.eps = eps
        self.unconstrained_weight = nn.Parameter(np.log(np.exp(1 - eps) - 1) * torch.ones(features))
        self.bias = nn.Parameter(torch.zeros(features))
        self.register_buffer("running_mean", torch.zeros(features))
        self.register_buffer("running_var", torch.zeros(features))

    @property
    def weight(self):
        return F.softplus(self.unconstrained_weight) + self.eps

    def forward(self, inputs, context=None):
        if inputs.dim() != 2:
            raise ValueError("Expected 2-dim inputs, got inputs of shape: {}".format(inputs.shape))
        if self.training:
            mean, var = inputs.mean(0), inputs.var(0)
            self.running_mean.mul_(1 - self.momentum).add_(mean.detach() * self.momentum)
            self.running_var.mul_(1 - self.momentum).add_(var.detach() * self.momentum)
        else:
            mean, var = self.running_mean, self.running_var
        outputs = self.weight * ((inputs - mean) / torch.sqrt(var + self.eps)) + self.bias
        logabsdet = torch.sum(torch.log(self.weight) - 0.5 * torch.log(var + self.eps))
        return outputs, logabsdet * inputs.new_ones(inputs.shape[0])

    def inverse(self, inputs, context=None):
        if self.training:
            raise InverseNotAvailable("Batch norm inverse is only available in eval mode, not in training mode.")
        if inputs.dim() != 2:
            raise ValueError("Expected 2-dim inputs, got inputs of shape: {}".format(inputs.shape))
        outputs = torch.sqrt(self.running_var + self.eps) * ((inputs - self.bias) / self.weight) + self.running_mean
        logabsdet = torch.sum(-torch.log(self.weight) + 0.5 * torch.log(self.running_var + self.eps))
        return outputs, logabsdet * inputs.new_ones(inputs.shape[0])
