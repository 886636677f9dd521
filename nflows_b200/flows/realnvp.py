"""SimpleRealNVP (reference nflows/flows/realnvp.py:17-71): a stack of affine (or additive) couplings on feature vectors
with an alternating +-1 mask; pure composition of the hot-path classes."""
import torch
from torch.nn import functional as F

from ..distributions.normal import StandardNormal
from ..nn import nets
from ..transforms.base import CompositeTransform
from ..transforms.coupling import AdditiveCouplingTransform, AffineCouplingTransform
from ..transforms.normalization import BatchNorm
from .base import Flow


class SimpleRealNVP(Flow):
    def __init__(self, features, hidden_features, num_layers, num_blocks_per_layer, use_volume_preserving=False,
                 activation=F.relu, dropout_probability=0.0, batch_norm_within_layers=False,
                 batch_norm_between_layers=False):
        coupling = AdditiveCouplingTransform if use_volume_preserving else AffineCouplingTransform
        mask = torch.ones(features)
        mask[::2] = -1

        def create_resnet(in_features, out_features):
            return nets.ResidualNet(in_features, out_features, hidden_features=hidden_features,
                                    num_blocks=num_blocks_per_layer, activation=activation,
                                    dropout_probability=dropout_probability, use_batch_norm=batch_norm_within_layers)

        layers = []
        for _ in range(num_layers):
            layers.append(coupling(mask=mask, transform_net_create_fn=create_resnet))
            mask *= -1
            if batch_norm_between_layers:
                layers.append(BatchNorm(features=features))
        super().__init__(transform=CompositeTransform(layers), distribution=StandardNormal([features]))
