"""OneByOneConvolution (reference nflows/transforms/conv.py:6-47): Glow's invertible 1x1 convolution = a fixed random
channel permutation followed by an LU-parameterised linear map applied at every pixel."""
from .lu import LULinear
from .permutations import RandomPermutation


class OneByOneConvolution(LULinear):
    """On [B, C, H, W]: every pixel's channel vector goes through the same `LULinear(C)`.  The pixel vectors are laid
    out as a [B*H*W, C] matrix, so CUDA fp32 inference runs the same folded tensor-core dense layer as `LULinear`;
    log|det| = H * W * sum(log diag U)."""

    def __init__(self, num_channels, using_cache=False, identity_init=True):
        super().__init__(num_channels, using_cache, identity_init)
        self.permutation = RandomPermutation(num_channels, dim=1)

    def _pixels(self, inputs, inverse):
        b, c, h, w = inputs.shape
        flat = inputs.permute(0, 2, 3, 1).reshape(b * h * w, c)
        out, lad = (super().inverse(flat) if inverse else super().forward(flat))
        out = out.reshape(b, h, w, c).permute(0, 3, 1, 2)
        return out, lad.reshape(b, h, w).sum(dim=(1, 2))

    def forward(self, inputs, context=None):
        if inputs.dim() != 4:
            raise ValueError("Inputs must be a 4D tensor.")
        inputs, _ = self.permutation(inputs)
        return self._pixels(inputs, inverse=False)

    def inverse(self, inputs, context=None):
        if inputs.dim() != 4:
            raise ValueError("Inputs must be a 4D tensor.")
        outputs, lad = self._pixels(inputs, inverse=True)
        outputs, _ = self.permutation.inverse(outputs)
        return outputs, lad
