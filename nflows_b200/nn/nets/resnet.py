"""Residual conditioner networks (reference nflows/nn/nets/resnet.py:9-205).

`ResidualNet` is the conditioner of the coupling transforms on the hot path.  When it is the plain relu / no
batch-norm / no context / no active dropout configuration, `dense_chain()` describes it as a list of dense
layers so the coupling can execute it with `nfk_linear` launches (bias, relu and the residual add are fused
into the GEMM epilogues)."""
import torch
from torch import nn
from torch.nn import functional as F
from torch.nn import init


class ResidualBlock(nn.Module):
    """Pre-activation residual block on feature vectors: x + W2 drop(act(bn(W1 act(bn(x)))))."""

    def __init__(self, features, context_features, activation=F.relu, dropout_probability=0.0, use_batch_norm=False,
                 zero_initialization=True):
        super().__init__()
        self.activation = activation
        self.use_batch_norm = use_batch_norm
        if use_batch_norm:
            self.batch_norm_layers = nn.ModuleList([nn.BatchNorm1d(features, eps=1e-3) for _ in range(2)])
        if context_features is not None:
            self.context_layer = nn.Linear(context_features, features)
        self.linear_layers = nn.ModuleList([nn.Linear(features, features) for _ in range(2)])
        self.dropout = nn.Dropout(p=dropout_probability)
        if zero_initialization:
            init.uniform_(self.linear_layers[-1].weight, -1e-3, 1e-3)
            init.uniform_(self.linear_layers[-1].bias, -1e-3, 1e-3)

    def forward(self, inputs, context=None):
        t = inputs
        if self.use_batch_norm:
            t = self.batch_norm_layers[0](t)
        t = self.linear_layers[0](self.activation(t))
        if self.use_batch_norm:
            t = self.batch_norm_layers[1](t)
        t = self.linear_layers[1](self.dropout(self.activation(t)))
        if context is not None:
            t = F.glu(torch.cat((t, self.context_layer(context)), dim=1), dim=1)
        return inputs + t


class ResidualNet(nn.Module):
    """Linear -> num_blocks residual blocks -> Linear, on feature vectors."""

    def __init__(self, in_features, out_features, hidden_features, context_features=None, num_blocks=2,
                 activation=F.relu, dropout_probability=0.0, use_batch_norm=False):
        super().__init__()
        self.hidden_features = hidden_features
        self.context_features = context_features
        first_in = in_features if context_features is None else in_features + context_features
        self.initial_layer = nn.Linear(first_in, hidden_features)
        self.blocks = nn.ModuleList([
            ResidualBlock(features=hidden_features, context_features=context_features, activation=activation,
                          dropout_probability=dropout_probability, use_batch_norm=use_batch_norm)
            for _ in range(num_blocks)
        ])
        self.final_layer = nn.Linear(hidden_features, out_features)

    def forward(self, inputs, context=None):
        t = self.initial_layer(inputs if context is None else torch.cat((inputs, context), dim=1))
        for block in self.blocks:
            t = block(t, context=context)
        return self.final_layer(t)

    def dense_chain(self, context=None):
        """[(weight, bias, relu_in, relu_out, residual)] or None when this net needs the generic torch path.
        residual: None, or "skip" = add the block input."""
        if context is not None or self.context_features is not None:
            return None
        chain = [(self.initial_layer.weight, self.initial_layer.bias, False, False, None)]
        for block in self.blocks:
            if block.use_batch_norm or block.activation is not F.relu:
                return None
            if block.dropout.p > 0.0 and block.training:
                return None
            l0, l1 = block.linear_layers
            chain.append((l0.weight, l0.bias, True, True, None))
            chain.append((l1.weight, l1.bias, False, False, "skip"))
        chain.append((self.final_layer.weight, self.final_layer.bias, False, False, None))
        return chain


class ConvResidualBlock(nn.Module):
    def __init__(self, channels, context_channels=None, activation=F.relu, dropout_probability=0.0, use_batch_norm=False,
                 zero_initialization=True):
        super().__init__()
        self.activation = activation
        if context_channels is not None:
            self.context_layer = nn.Conv2d(in_channels=context_channels, out_channels=channels, kernel_size=1, padding=0)
        self.use_batch_norm = use_batch_norm
        if use_batch_norm:
            self.batch_norm_layers = nn.ModuleList([nn.BatchNorm2d(channels, eps=1e-3) for _ in range(2)])
        self.conv_layers = nn.ModuleList([nn.Conv2d(channels, channels, kernel_size=3, padding=1) for _ in range(2)])
        self.dropout = nn.Dropout(p=dropout_probability)
        if zero_initialization:
            init.uniform_(self.conv_layers[-1].weight, -1e-3, 1e-3)
            init.uniform_(self.conv_layers[-1].bias, -1e-3, 1e-3)

    def forward(self, inputs, context=None):
        t = inputs
        if self.use_batch_norm:
            t = self.batch_norm_layers[0](t)
        t = self.conv_layers[0](self.activation(t))
        if self.use_batch_norm:
            t = self.batch_norm_layers[1](t)
        t = self.conv_layers[1](self.dropout(self.activation(t)))
        if context is not None:
            t = F.glu(torch.cat((t, self.context_layer(context)), dim=1), dim=1)
        return inputs + t


class ConvResidualNet(nn.Module):
    """1x1 conv -> residual 3x3 blocks -> 1x1 conv (image conditioner; torch/cuDNN path, off the hot path)."""

    def __init__(self, in_channels, out_channels, hidden_channels, context_channels=None, num_blocks=2,
                 activation=F.relu, dropout_probability=0.0, use_batch_norm=False):
        super().__init__()
        self.context_channels = context_channels
        self.hidden_channels = hidden_channels
        first_in = in_channels if context_channels is None else in_channels + context_channels
        self.initial_layer = nn.Conv2d(in_channels=first_in, out_channels=hidden_channels, kernel_size=1, padding=0)
        self.blocks = nn.ModuleList([
            ConvResidualBlock(channels=hidden_channels, context_channels=context_channels, activation=activation,
                              dropout_probability=dropout_probability, use_batch_norm=use_batch_norm)
            for _ in range(num_blocks)
        ])
        self.final_layer = nn.Conv2d(hidden_channels, out_channels, kernel_size=1, padding=0)

    def forward(self, inputs, context=None):
        t = self.initial_layer(inputs if context is None else torch.cat((inputs, context), dim=1))
        for block in self.blocks:
            t = block(t, context)
        return self.final_layer(t)
