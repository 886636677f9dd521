"""MADE: masked autoregressive network (reference nflows/transforms/made.py:17-283), the conditioner of the
autoregressive transforms.  Masks are fixed 0/1 buffers; a masked layer is a dense layer with weight * mask, so on the
native path it runs on the same tensor-core dense kernels as the coupling conditioners (`dense_chain`)."""
import torch
from torch import nn
from torch.nn import functional as F
from torch.nn import init

from ..utils import torchutils


def _get_input_degrees(in_features):
    """Degrees 1..D of the D inputs."""
    return torch.arange(1, in_features + 1)


class MaskedLinear(nn.Linear):
    """nn.Linear whose weight is multiplied by a fixed autoregressive mask.  Hidden units get degrees
    `arange(H) % (D-1) + 1` (or random ones) and may see inputs of degree <= their own; the output layer repeats each
    input degree `multiplier` times CONSECUTIVELY (row j*multiplier + k belongs to feature j) and may only see hidden
    units of strictly smaller degree."""

    def __init__(self, in_degrees, out_features, autoregressive_features, random_mask, is_output, bias=True):
        super().__init__(in_features=len(in_degrees), out_features=out_features, bias=bias)
        mask, degrees = self._get_mask_and_degrees(in_degrees=in_degrees, out_features=out_features,
                                                   autoregressive_features=autoregressive_features,
                                                   random_mask=random_mask, is_output=is_output)
        self.register_buffer("mask", mask)
        self.register_buffer("degrees", degrees)
        self._masked_cache = None

    @classmethod
    def _get_mask_and_degrees(cls, in_degrees, out_features, autoregressive_features, random_mask, is_output):
        if is_output:
            out_degrees = torchutils.tile(_get_input_degrees(autoregressive_features), out_features // autoregressive_features)
            mask = (out_degrees[..., None] > in_degrees).float()
        else:
            if random_mask:
                low = min(torch.min(in_degrees).item(), autoregressive_features - 1)
                out_degrees = torch.randint(low=low, high=autoregressive_features, size=[out_features], dtype=torch.long)
            else:
                max_ = max(1, autoregressive_features - 1)
                min_ = min(1, autoregressive_features - 1)
                out_degrees = torch.arange(out_features) % max_ + min_
            mask = (out_degrees[..., None] >= in_degrees).float()
        return mask, out_degrees

    def forward(self, x):
        return F.linear(x, self.weight * self.mask, self.bias)

    def masked_weight(self):
        """weight * mask as a tensor object that stays the same until the weight changes (so split-operand caches hit)."""
        from .. import config
        sig = (self.weight.data_ptr(), self.weight._version, str(self.weight.device), config.cache_epoch)
        if self._masked_cache is None or self._masked_cache[0] != sig:
            self._masked_cache = (sig, (self.weight.detach() * self.mask).contiguous())
        return self._masked_cache[1]


class MaskedFeedforwardBlock(nn.Module):
    """bn? -> masked linear -> activation -> dropout (same width in and out)."""

    def __init__(self, in_degrees, autoregressive_features, context_features=None, random_mask=False, activation=F.relu,
                 dropout_probability=0.0, use_batch_norm=False):
        super().__init__()
        features = len(in_degrees)
        self.batch_norm = nn.BatchNorm1d(features, eps=1e-3) if use_batch_norm else None
        self.linear = MaskedLinear(in_degrees=in_degrees, out_features=features,
                                   autoregressive_features=autoregressive_features, random_mask=random_mask, is_output=False)
        self.degrees = self.linear.degrees
        self.activation = activation
        self.dropout = nn.Dropout(p=dropout_probability)

    def forward(self, inputs, context=None):
        t = self.batch_norm(inputs) if self.batch_norm else inputs
        return self.dropout(self.activation(self.linear(t)))


class MaskedResidualBlock(nn.Module):
    """Pre-activation residual block of two masked linears; degrees are preserved so the skip connection is legal."""

    def __init__(self, in_degrees, autoregressive_features, context_features=None, random_mask=False, activation=F.relu,
                 dropout_probability=0.0, use_batch_norm=False, zero_initialization=True):
        if random_mask:
            raise ValueError("Masked residual block can't be used with random masks.")
        super().__init__()
        features = len(in_degrees)
        if context_features is not None:
            self.context_layer = nn.Linear(context_features, features)
        self.use_batch_norm = use_batch_norm
        if use_batch_norm:
            self.batch_norm_layers = nn.ModuleList([nn.BatchNorm1d(features, eps=1e-3) for _ in range(2)])
        linear_0 = MaskedLinear(in_degrees=in_degrees, out_features=features, autoregressive_features=autoregressive_features,
                                random_mask=False, is_output=False)
        linear_1 = MaskedLinear(in_degrees=linear_0.degrees, out_features=features,
                                autoregressive_features=autoregressive_features, random_mask=False, is_output=False)
        self.linear_layers = nn.ModuleList([linear_0, linear_1])
        self.degrees = linear_1.degrees
        if torch.all(self.degrees >= in_degrees).item() != 1:
            raise RuntimeError("In a masked residual block, the output degrees can't be less than the corresponding input degrees.")
        self.activation = activation
        self.dropout = nn.Dropout(p=dropout_probability)
        if zero_initialization:
            init.uniform_(self.linear_layers[-1].weight, a=-1e-3, b=1e-3)
            init.uniform_(self.linear_layers[-1].bias, a=-1e-3, b=1e-3)

    def forward(self, inputs, context=None):
        t = inputs
        if self.use_batch_norm:
            t = self.batch_norm_layers[0](t)
        t = self.linear_layers[0](self.activation(t))
        if context is not None:
            t = t + self.context_layer(context)
        if self.use_batch_norm:
            t = self.batch_norm_layers[1](t)
        t = self.linear_layers[1](self.dropout(self.activation(t)))
        return inputs + t


class MADE(nn.Module):
    """Masked initial layer -> residual (default) or feedforward masked blocks -> masked output layer producing
    `output_multiplier` values per feature, feature-major.  NOTE: like the reference class it does NOT expose
    `hidden_features`, so spline transforms built on it do not rescale their softmax logits."""

    def __init__(self, features, hidden_features, context_features=None, num_blocks=2, output_multiplier=1,
                 use_residual_blocks=True, random_mask=False, activation=F.relu, dropout_probability=0.0,
                 use_batch_norm=False):
        if use_residual_blocks and random_mask:
            raise ValueError("Residual blocks can't be used with random masks.")
        super().__init__()
        self.initial_layer = MaskedLinear(in_degrees=_get_input_degrees(features), out_features=hidden_features,
                                          autoregressive_features=features, random_mask=random_mask, is_output=False)
        if context_features is not None:
            self.context_layer = nn.Linear(context_features, hidden_features)
        self.use_residual_blocks = use_residual_blocks
        self.activation = activation
        block_cls = MaskedResidualBlock if use_residual_blocks else MaskedFeedforwardBlock
        blocks = []
        degrees = self.initial_layer.degrees
        for _ in range(num_blocks):
            blocks.append(block_cls(in_degrees=degrees, autoregressive_features=features, context_features=context_features,
                                    random_mask=random_mask, activation=activation, dropout_probability=dropout_probability,
                                    use_batch_norm=use_batch_norm))
            degrees = blocks[-1].degrees
        self.blocks = nn.ModuleList(blocks)
        self.final_layer = MaskedLinear(in_degrees=degrees, out_features=features * output_multiplier,
                                        autoregressive_features=features, random_mask=random_mask, is_output=True)

    def forward(self, inputs, context=None):
        t = self.initial_layer(inputs)
        if context is not None:
            t = t + self.activation(self.context_layer(context))
        if not self.use_residual_blocks:
            t = self.activation(t)
        for block in self.blocks:
            t = block(t, context)
        return self.final_layer(t)

    def dense_chain(self, context=None):
        """[(weight*mask, bias, relu_in, relu_out, residual)] for the relu / residual / no-BN / no-context case."""
        if context is not None or hasattr(self, "context_layer") or not self.use_residual_blocks or self.activation is not F.relu:
            return None
        chain = [(self.initial_layer.masked_weight(), self.initial_layer.bias, False, False, None)]
        for block in self.blocks:
            if block.use_batch_norm or block.activation is not F.relu or (block.dropout.p > 0.0 and block.training):
                return None
            l0, l1 = block.linear_layers
            chain.append((l0.masked_weight(), l0.bias, True, True, None))
            chain.append((l1.masked_weight(), l1.bias, False, False, "skip"))
        chain.append((self.final_layer.masked_weight(), self.final_layer.bias, False, False, None))
        return chain
