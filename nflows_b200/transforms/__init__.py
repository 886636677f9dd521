from .autoregressive import (AutoregressiveTransform, MaskedAffineAutoregressiveTransform,
                             MaskedPiecewiseRationalQuadraticAutoregressiveTransform)
from .base import (CompositeTransform, InputOutsideDomain, InverseNotAvailable, InverseTransform,
                   MultiscaleCompositeTransform, Transform)
from .coupling import (AdditiveCouplingTransform, AffineCouplingTransform, CouplingTransform,
                       PiecewiseCouplingTransform, PiecewiseRationalQuadraticCouplingTransform)
from .linear import Linear
from .lu import LULinear
from .normalization import ActNorm, BatchNorm
from .conv import OneByOneConvolution
from .nonlinearities import PiecewiseRationalQuadraticCDF
from .reshape import SqueezeTransform
from .permutations import Permutation, RandomPermutation, ReversePermutation
from .standard import AffineScalarTransform, AffineTransform, IdentityTransform, PointwiseAffineTransform
from . import splines
