"""nflows_b200: B200-native (sm_100a) implementation of the nflows coupling-flow hot path.

Drop-in for `nflows.transforms.Transform / CompositeTransform`, the coupling / ActNorm / LULinear /
Permutation transforms, `Flow.log_prob / sample` and `StandardNormal` -- same class names, constructor
kwargs, state_dict keys and exceptions -- with CUDA fp32 inference executed by hand-written kernels in
libnfk_sm100.so (C ABI: include/nfk.h)."""
__version__ = "0.1.0"


class _Config:
    #: read the device flag word after each public call and raise the reference's exceptions
    #: (InputOutsideDomain / AssertionError).  Costs one device->host sync per call.
    check_domain = True
    #: upper bound (MiB) for the conditioner-output chunk a coupling keeps in flight; sized to stay in B200's L2
    param_chunk_mib = 64
    #: run the final conditioner layer and the spline as ONE tensor-core kernel when an instance exists
    fuse_coupling = True
    #: hand the fused kernel the (hi, lo) split pair of the hidden activation (written by the last trunk layer) instead of
    #: the fp32 activation it then splits on chip; kept for A/B measurements
    fused_pair_input = False
    #: rows per sub-block of a dense-layer chain: intermediates of a sub-block (split pairs, hidden activations) stay
    #: resident in the 126 MB L2 between consecutive kernels instead of round-tripping through HBM
    #: (measured r1: sub-blocks of 8-16 K rows are SLOWER -- 1-wave launches pay prologue/launch overhead -- so the default
    #: keeps whole 256 K-row blocks; the knob stays for a future persistent / graph-captured executor)
    trunk_block_rows = 1 << 18
    affine_block_rows = 1 << 18


config = _Config()
