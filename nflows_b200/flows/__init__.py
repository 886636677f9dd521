from .base import Flow
from .realnvp import SimpleRealNVP
from . import realnvp
