"""SqueezeTransform (reference nflows/transforms/reshape.py:7-68): space-to-depth with a 2x2 (factor x factor) window,
the RealNVP "squeeze".  Pure index shuffling (log|det| = 0)."""

from ..utils import typechecks as check
from .base import Transform


class SqueezeTransform(Transform):
    """[B, C, H, W] -> [B, C*f*f, H/f, W/f]; channel c*f*f + i*f + j holds pixel (i, j) of each f x f window of channel c."""

    def __init__(self, factor=2):
        super().__init__()
        if not check.is_int(factor) or factor <= 1:
            raise ValueError("Factor must be an integer > 1.")
        self.factor = factor

    def get_output_shape(self, c, h, w):
        return (c * self.factor * self.factor, h // self.factor, w // self.factor)

    def forward(self, inputs, context=None):
        if inputs.dim() != 4:
            raise ValueError("Expecting inputs with 4 dimensions")
        b, c, h, w = inputs.size()
        f = self.factor
        if h % f != 0 or w % f != 0:
            raise ValueError("Input image size not compatible with the factor.")
        x = inputs.view(b, c, h // f, f, w // f, f).permute(0, 1, 3, 5, 2, 4).contiguous()
        return x.view(b, c * f * f, h // f, w // f), inputs.new_zeros(b)

    def inverse(self, inputs, context=None):
        if inputs.dim() != 4:
            raise ValueError("Expecting inputs with 4 dimensions")
        b, c, h, w = inputs.size()
        f = self.factor
        if c < 4 or c % 4 != 0:
            raise ValueError("Invalid number of channel dimensions.")
        x = inputs.view(b, c // (f * f), f, f, h, w).permute(0, 1, 4, 2, 5, 3).contiguous()
        return x.view(b, c // (f * f), h * f, w * f), inputs.new_zeros(b)
