"""Helpers the hot path uses (reference nflows/utils/torchutils.py:19-52, 89-136).

These are shape/mask utilities executed on the host or as cheap torch views; the arithmetic versions that
sit on the hot path (searchsorted inside the spline, sum_except_batch of log-dets) are fused into the
CUDA kernels and only re-exposed here for API compatibility."""
import torch

from . import typechecks as check


def tile(x, n):
    """[a, b] -> [a]*n + [b]*n (reference :8-16); used for interleaved parameter layouts."""
    if not check.is_positive_int(n):
        raise TypeError("Argument 'n' must be a positive integer.")
    return x.reshape(-1).repeat_interleave(n)


def sum_except_batch(x, num_batch_dims=1):
    if not check.is_nonnegative_int(num_batch_dims):
        raise TypeError("Number of batch dimensions must be a non-negative integer.")
    dims = list(range(num_batch_dims, x.ndimension()))
    return torch.sum(x, dim=dims)


def split_leading_dim(x, shape):
    return torch.reshape(x, torch.Size(shape) + x.shape[1:])


def merge_leading_dims(x, num_dims):
    if not check.is_positive_int(num_dims):
        raise TypeError("Number of leading dims must be a positive integer.")
    if num_dims > x.dim():
        raise ValueError("Number of leading dims can't be greater than total number of dims.")
    return torch.reshape(x, torch.Size([-1]) + x.shape[num_dims:])


def repeat_rows(x, num_reps):
    if not check.is_positive_int(num_reps):
        raise TypeError("Number of repetitions must be a positive integer.")
    return x.repeat_interleave(num_reps, dim=0)


def tensor2numpy(x):
    return x.detach().cpu().numpy()


def logabsdet(x):
    return torch.slogdet(x)[1]


def get_num_parameters(model):
    return sum(p.numel() for p in model.parameters())


def create_alternating_binary_mask(features, even=True):
    mask = torch.zeros(features, dtype=torch.uint8)
    mask[(0 if even else 1)::2] = 1
    return mask


def create_mid_split_binary_mask(features):
    mask = torch.zeros(features, dtype=torch.uint8)
    mask[:(features + 1) // 2] = 1
    return mask


def create_random_binary_mask(features):
    mask = torch.zeros(features, dtype=torch.uint8)
    chosen = torch.multinomial(torch.ones(features), num_samples=(features + 1) // 2, replacement=False)
    mask[chosen] = 1
    return mask


def searchsorted(bin_locations, inputs, eps=1e-6):
    """Index of the bin each input falls in; like the reference (:134-136) the last location is bumped by
    eps IN PLACE so the right edge belongs to the last bin."""
    bin_locations[..., -1] += eps
    return torch.sum(inputs[..., None] >= bin_locations, dim=-1) - 1


def random_orthogonal(size):
    """Random orthogonal [size, size] matrix (Q factor of a Gaussian matrix)."""
    q, _ = torch.linalg.qr(torch.randn(size, size))
    return q


def cbrt(x):
    """Real cube root."""
    return torch.sign(x) * torch.exp(torch.log(torch.abs(x)) / 3.0)
