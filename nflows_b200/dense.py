"""Execution of dense-layer chains (conditioner networks, folded affine maps) on the native GEMM kernels.

Two kernels implement the same contract (fp32-equivalent products, fp32 accumulation):
  * "tc"   -- `nfk_linear_tf32x3`: tcgen05 tensor cores, operands carried as (hi, lo) TF32 split pairs; the default.
  * "simt" -- `nfk_linear`: FP32 FFMA pipe; used for shapes the TMA path cannot address (in_features not a multiple
              of 4) and selectable with NFLOWS_B200_GEMM=simt for A/B comparisons.
Both are CUDA kernels of this library; there is no PyTorch/cuBLAS path here."""
import os

import torch

from . import kernels as K

_SPLIT_CACHE = {}


def backend():
    return os.environ.get("NFLOWS_B200_GEMM", "tc")


def split_weight(weight):
    """(hi, lo) pair of a weight matrix, cached until the parameter is modified."""
    w = weight.detach()
    key = id(weight)
    sig = (w.data_ptr(), w._version, str(w.device), tuple(w.shape))
    hit = _SPLIT_CACHE.get(key)
    if hit is None or hit[0] != sig:
        if not w.is_contiguous():
            w = w.contiguous()
        hit = (sig, K.split_tf32(w), weight)
        _SPLIT_CACHE[key] = hit
        if len(_SPLIT_CACHE) > 4096:
            _SPLIT_CACHE.pop(next(iter(_SPLIT_CACHE)))
    return hit[1]


def chain_uses_tc(chain, first_in_features):
    if backend() != "tc":
        return False
    k = first_in_features
    for weight, _, _, _, _ in chain:
        if weight.shape[1] != k or not K.tf32x3_supported(k, weight.stride(0), k):
            return False
        k = weight.shape[0]
    return True


class ChainState:
    """Activation between two layers: fp32 tensor (`raw`) and/or the split pair a tensor-core layer consumes."""

    def __init__(self, raw=None, pair=None):
        self.raw, self.pair = raw, pair


def run_trunk(chain, x, id_cols, use_tc, want_pair=False, x_id=None):
    """All layers of `chain` but the last, on the identity columns of x.  Returns the ChainState feeding the last layer:
    the split pair of its (pre-activated) input when want_pair, else the fp32 tensor.
    x_id: the identity columns as a (possibly strided) view when they are contiguous in x -- no gather pass then.
    The tensor-core path walks row sub-blocks so that the intermediates of a sub-block are reused out of L2."""
    from . import config
    n = x.shape[0]
    step = max(128, int(config.trunk_block_rows))
    if use_tc and n > step and len(chain) > 1:
        width = chain[-2][0].shape[0]
        if want_pair:
            dst = (torch.empty(n, width, dtype=torch.float32, device=x.device),
                   torch.empty(n, width, dtype=torch.float32, device=x.device))
        else:
            dst = torch.empty(n, width, dtype=torch.float32, device=x.device)
        for r0 in range(0, n, step):
            r1 = min(n, r0 + step)
            _run_trunk_block(chain, x[r0:r1], id_cols, True,
                             (dst[0][r0:r1], dst[1][r0:r1]) if want_pair else dst[r0:r1], want_pair,
                             None if x_id is None else x_id[r0:r1])
        return ChainState(pair=dst) if want_pair else ChainState(raw=dst)
    return _run_trunk_block(chain, x, id_cols, use_tc, None, want_pair, x_id)


def _run_trunk_block(chain, x, id_cols, use_tc, last_out, want_pair, x_id=None):
    body = chain[:-1]
    last_relu_in = chain[-1][2]
    if use_tc:
        # every layer reads the fp32 activation its producer wrote and splits it on chip (nfk_linear_tf32x3_a32); only the
        # input of the LAST layer is materialised as a split pair when the fused coupling kernel (which re-reads it once per
        # column tile) consumes it
        if x_id is None:
            x_id = x if id_cols is None else K.gather_cols(x, id_cols)
        if not body:
            return ChainState(pair=K.split_tf32(x_id, relu=last_relu_in)) if want_pair else ChainState(raw=x_id)
        hidden, skip_src = x_id, None
        state = None
        for i, (weight, bias, relu_in, relu_out, residual) in enumerate(body):
            last = i == len(body) - 1
            res = skip_src if residual == "skip" else None
            b = bias.detach() if bias is not None else None
            if last and want_pair:
                _, pair = K.linear_tf32x3(hidden, split_weight(weight), b, residual=res, relu_in=relu_in, relu_out=relu_out,
                                          want_y=False, want_split=True, split_relu=last_relu_in, pair_out=last_out)
                state = ChainState(pair=pair)
            else:
                y, _ = K.linear_tf32x3(hidden, split_weight(weight), b, residual=res, relu_in=relu_in, relu_out=relu_out,
                                       want_y=True, y_out=last_out if last else None)
                state = ChainState(raw=y)
                if i + 2 < len(chain) and chain[i + 2][4] == "skip":
                    skip_src = y                       # input of the residual block that starts with the next layer
                hidden = y
        return state
    hidden = x if id_cols is None else K.gather_cols(x, id_cols)
    branch = None
    for weight, bias, relu_in, relu_out, residual in body:
        b = bias.detach() if bias is not None else None
        if residual == "skip":
            hidden = K.linear(branch, weight.detach(), b, residual=hidden, relu_in=relu_in, relu_out=relu_out, out=hidden)
        elif relu_in:   # first layer of a residual block: keep its input for the skip connection
            branch = K.linear(hidden, weight.detach(), b, relu_in=True, relu_out=relu_out)
        else:
            hidden = K.linear(hidden, weight.detach(), b, relu_in=relu_in, relu_out=relu_out)
    return ChainState(raw=hidden)


def run_last(chain, state, r0, r1, use_tc):
    """Last layer of the chain on rows [r0, r1) of the trunk output -> fp32 conditioner output."""
    weight, bias, relu_in, relu_out, _ = chain[-1]
    b = bias.detach() if bias is not None else None
    if use_tc:
        return K.linear_tf32x3(state.raw[r0:r1], split_weight(weight), b, relu_in=relu_in, relu_out=relu_out, want_y=True)[0]
    return K.linear(state.raw[r0:r1], weight.detach(), b, relu_in=relu_in, relu_out=relu_out)


def affine_map(x, weight, bias):
    """y = x @ weight.T + bias for a folded ActNorm/Permutation/LU run: one tensor-core GEMM reading x as it is."""
    from . import config
    n, k = x.shape
    if backend() == "tc" and K.tf32x3_supported(x.stride(0), weight.stride(0), k):
        w_pair = split_weight(weight)
        y = torch.empty(n, weight.shape[0], dtype=torch.float32, device=x.device)
        step = max(128, int(config.affine_block_rows))
        for r0 in range(0, n, step):
            r1 = min(n, r0 + step)
            K.linear_tf32x3(x[r0:r1], w_pair, bias, want_y=True, y_out=y[r0:r1])
        return y
    return K.linear(x, weight, bias)


_PACK_CACHE = {}


def pack_final_spline(weight, bias, d_t, m, mp):
    """Packed operands of the fused coupling kernel: rows regrouped to `mp` per transformed feature (zero padded), split
    into the (hi, lo) pair; bias packed the same way.  Cached until weight or bias is modified."""
    w, b = weight.detach(), bias.detach()
    key = id(weight)
    sig = (w.data_ptr(), w._version, b.data_ptr(), b._version, str(w.device), d_t, m, mp)
    hit = _PACK_CACHE.get(key)
    if hit is None or hit[0] != sig:
        k = w.shape[1]
        wp = w.new_zeros(d_t, mp, k)
        wp[:, :m, :] = w.reshape(d_t, m, k)
        fpt = 128 // mp                                   # features per thread in the fused kernel (FusedCfg::FPT)
        tiles = -(-d_t // (2 * fpt))
        bp = b.new_zeros(tiles * 2 * fpt, mp)             # padded to whole tiles: the kernel reads bias per tile column
        bp[:d_t, :m] = b.reshape(d_t, m)
        hit = (sig, K.split_tf32(wp.reshape(d_t * mp, k)), bp.reshape(-1).contiguous(), weight)
        _PACK_CACHE[key] = hit
        if len(_PACK_CACHE) > 1024:
            _PACK_CACHE.pop(next(iter(_PACK_CACHE)))
    return hit[1], hit[2]
