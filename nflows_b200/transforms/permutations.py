"""Permutation transforms (reference nflows/transforms/permutations.py:9-63).  Indexing is bit-exact."""
import torch

from .. import kernels as K
from ..utils import typechecks as check
from .base import Transform


class Permutation(Transform):
    """outputs = inputs.index_select(dim, permutation); inverse uses argsort(permutation)."""

    def __init__(self, permutation, dim=1):
        if permutation.ndimension() != 1:
            raise ValueError("Permutation must be a 1D tensor.")
        if not check.is_positive_int(dim):
            raise ValueError("dim must be a positive integer.")
        super().__init__()
        self._dim = dim
        self.register_buffer("_permutation", permutation)
        self._idx_cache = {}

    @property
    def _inverse_permutation(self):
        return torch.argsort(self._permutation)

    def _index_i32(self, inverse, device):
        key = (str(device), self._permutation.data_ptr(), self._permutation._version)
        hit = self._idx_cache.get(inverse)
        if hit is None or hit[0] != key:
            perm = self._inverse_permutation if inverse else self._permutation
            hit = (key, K.index_tensor(perm, device))
            self._idx_cache[inverse] = hit
        return hit[1]

    def _check(self, inputs):
        if self._dim >= inputs.ndimension():
            raise ValueError("No dimension {} in inputs.".format(self._dim))
        if inputs.shape[self._dim] != len(self._permutation):
            raise ValueError("Dimension {} in inputs must be of size {}.".format(self._dim, len(self._permutation)))

    def _native_ready(self, inputs, context):
        return K.native_ok(inputs, context) and inputs.dim() == 2 and self._dim == 1

    def _native_apply(self, inputs, lad, flags, inverse, context=None):
        self._check(inputs)
        return K.gather_cols(inputs, self._index_i32(inverse, inputs.device))

    def _eager(self, inputs, context, inverse):
        self._check(inputs)
        perm = self._inverse_permutation if inverse else self._permutation
        return torch.index_select(inputs, self._dim, perm), inputs.new_zeros(inputs.shape[0])


class RandomPermutation(Permutation):
    """A fixed permutation drawn with torch.randperm at construction."""

    def __init__(self, features, dim=1):
        if not check.is_positive_int(features):
            raise ValueError("Number of features must be a positive integer.")
        super().__init__(torch.randperm(features), dim)


class ReversePermutation(Permutation):
    """Reverses the order of the features."""

    def __init__(self, features, dim=1):
        if not check.is_positive_int(features):
            raise ValueError("Number of features must be a positive integer.")
        super().__init__(torch.arange(features - 1, -1, -1), dim)
