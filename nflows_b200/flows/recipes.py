"""Constructors for the flows BASELINE.json names.  Pure composition of the public classes."""
from torch.nn import functional as F

from .. import transforms as T
from ..distributions.normal import StandardNormal
from ..nn.nets import ResidualNet
from ..utils import torchutils
from .base import Flow


def rq_nsf(features, hidden_features=256, num_layers=10, num_bins=8, tail_bound=3.0, num_blocks=2):
    """cfg 3: num_layers x [ActNorm, Composite[RandomPermutation, LULinear], RQ coupling (alternating mask,
    ResidualNet conditioner)], StandardNormal base.  Consumes the torch RNG in the same order as building the
    same stack from the reference's classes, so a seed reproduces the reference's weights."""
    steps = []
    for i in range(num_layers):
        steps.append(T.ActNorm(features))
        steps.append(T.CompositeTransform([T.RandomPermutation(features), T.LULinear(features, identity_init=True)]))
        steps.append(T.PiecewiseRationalQuadraticCouplingTransform(
            mask=torchutils.create_alternating_binary_mask(features, even=(i % 2 == 0)),
            transform_net_create_fn=lambda i_, o_: ResidualNet(i_, o_, hidden_features=hidden_features,
                                                               num_blocks=num_blocks, activation=F.relu),
            num_bins=num_bins, tails="linear", tail_bound=tail_bound))
    return Flow(T.CompositeTransform(steps), StandardNormal([features]))


def rq_coupling_layer(features=64, hidden_features=128, num_bins=8, tail_bound=3.0, num_blocks=2):
    """cfg 2: a single RQ coupling with an alternating mask."""
    return T.PiecewiseRationalQuadraticCouplingTransform(
        mask=torchutils.create_alternating_binary_mask(features),
        transform_net_create_fn=lambda i_, o_: ResidualNet(i_, o_, hidden_features=hidden_features, num_blocks=num_blocks),
        num_bins=num_bins, tails="linear", tail_bound=tail_bound)


def affine_flow_2d(hidden_features=8):
    """cfg 1: two affine couplings on 2-D data with complementary masks."""
    f = lambda i_, o_: ResidualNet(i_, o_, hidden_features=hidden_features)
    return Flow(T.CompositeTransform([T.AffineCouplingTransform(mask=[1, 0], transform_net_create_fn=f),
                                      T.AffineCouplingTransform(mask=[0, 1], transform_net_create_fn=f)]),
                StandardNormal([2]))


def perturb_(flow, seed=2):
    """The well-conditioned perturbation of SURVEY.md section 8d (makes ActNorm / LU / splines non-trivial)."""
    import numpy as np
    import torch
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in flow.named_parameters():
            leaf = name.split(".")[-1]
            if leaf in ("lower_entries", "upper_entries"):
                d = (1 + int(np.sqrt(1 + 8 * p.numel()))) // 2
                p.add_((0.1 / np.sqrt(d)) * torch.randn(p.shape, generator=g))
            elif leaf in ("log_scale", "shift", "unconstrained_upper_diag") or (leaf == "bias" and "transform_net" not in name):
                p.add_(0.1 * torch.randn(p.shape, generator=g))
            elif "final_layer" in name:
                p.mul_(3.0)
    return flow
