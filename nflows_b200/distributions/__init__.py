from .base import Distribution, NoMeanException
from .normal import ConditionalDiagonalNormal, DiagonalNormal, StandardNormal
