"""Base class of linear transforms with the eval-mode cache protocol (reference nflows/transforms/linear.py:14-126)."""
import torch
from torch import nn
from torch.nn import functional as F

from ..utils import typechecks as check
from .base import Transform


class LinearCache:
    """Weight, inverse weight and log|det| of a linear transform, filled lazily in eval mode."""

    def __init__(self):
        self.invalidate()

    def invalidate(self):
        self.weight = None
        self.inverse = None
        self.logabsdet = None


class Linear(Transform):
    """Abstract linear transform y = W x + bias.  Subclasses provide weight(), weight_inverse(), logabsdet(),
    forward_no_cache() and inverse_no_cache().  In eval mode with ``using_cache`` the dense weight (or its
    inverse) and the log-determinant are computed once and reused until ``train()`` is called."""

    def __init__(self, features, using_cache=False):
        if not check.is_positive_int(features):
            raise TypeError("Number of features must be a positive integer.")
        super().__init__()
        self.features = features
        self.bias = nn.Parameter(torch.zeros(features))
        self.using_cache = using_cache
        self.cache = LinearCache()

    def _cached_mode(self):
        return (not self.training) and self.using_cache

    def _eager(self, inputs, context, inverse):
        if not self._cached_mode():
            return self.inverse_no_cache(inputs) if inverse else self.forward_no_cache(inputs)
        if inverse:
            self._check_inverse_cache()
            outputs = F.linear(inputs - self.bias, self.cache.inverse)
            return outputs, (-self.cache.logabsdet) * outputs.new_ones(outputs.shape[0])
        self._check_forward_cache()
        outputs = F.linear(inputs, self.cache.weight, self.bias)
        return outputs, self.cache.logabsdet * outputs.new_ones(outputs.shape[0])

    def _check_forward_cache(self):
        if self.cache.weight is None and self.cache.logabsdet is None:
            self.cache.weight, self.cache.logabsdet = self.weight_and_logabsdet()
        elif self.cache.weight is None:
            self.cache.weight = self.weight()
        elif self.cache.logabsdet is None:
            self.cache.logabsdet = self.logabsdet()

    def _check_inverse_cache(self):
        if self.cache.inverse is None and self.cache.logabsdet is None:
            self.cache.inverse, self.cache.logabsdet = self.weight_inverse_and_logabsdet()
        elif self.cache.inverse is None:
            self.cache.inverse = self.weight_inverse()
        elif self.cache.logabsdet is None:
            self.cache.logabsdet = self.logabsdet()

    def train(self, mode=True):
        if mode:
            self.cache.invalidate()
        return super().train(mode)

    def use_cache(self, mode=True):
        if not check.is_bool(mode):
            raise TypeError("Mode must be boolean.")
        self.using_cache = mode

    def weight_and_logabsdet(self):
        return self.weight(), self.logabsdet()

    def weight_inverse_and_logabsdet(self):
        return self.weight_inverse(), self.logabsdet()

    def forward_no_cache(self, inputs):
        raise NotImplementedError()

    def inverse_no_cache(self, inputs):
        raise NotImplementedError()

    def weight(self):
        raise NotImplementedError()

    def weight_inverse(self):
        raise NotImplementedError()

    def logabsdet(self):
        raise NotImplementedError()
