from .mlp import MLP
from .resnet import ConvResidualNet, ResidualNet
