"""Batch sharding of Flow.log_prob / Flow.sample across the GPUs of one box, and host-streamed evaluation.

Every sample of the path is independent (no batch statistics in eval mode), so the path shards with NO data-path
collective: one process per GPU (torchrun), a full replica of the weights per rank, rank r owns rows
[lo_r, hi_r).  The only exchange is the optional all-gather of the per-sample log-probs at the end (4 bytes per row,
NCCL over NVLink; latency-bound).  The reference has no distributed code at all."""
import os

import torch
import torch.distributed as dist


def shard_bounds(n_rows, world_size, rank):
    """Contiguous near-equal shards: the first n_rows % world_size ranks get one extra row."""
    base, extra = divmod(int(n_rows), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _world(group):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def gather_rows(local, n_rows_total, group=None):
    """All-gather per-row results of ragged contiguous shards back into global row order on every rank."""
    world, rank = _world(group)
    if world == 1:
        return local
    sizes = [shard_bounds(n_rows_total, world, r) for r in range(world)]
    counts = [hi - lo for lo, hi in sizes]
    if len(set(counts)) == 1:
        out = local.new_empty((n_rows_total,) + tuple(local.shape[1:]))
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    width = max(counts)
    padded = local.new_zeros((width,) + tuple(local.shape[1:]))
    padded[: local.shape[0]] = local
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)


@torch.no_grad()
def log_prob_sharded(flow, inputs, context=None, group=None, gather=True):
    """`inputs` (and `context`) are the FULL batch, replicated or addressable on every rank; each rank evaluates its
    shard with its own replica of `flow` and (optionally) all ranks receive the full [n_rows] result."""
    world, rank = _world(group)
    lo, hi = shard_bounds(inputs.shape[0], world, rank)
    local = flow.log_prob(inputs[lo:hi], None if context is None else context[lo:hi])
    return gather_rows(local, inputs.shape[0], group) if gather else local


@torch.no_grad()
def log_prob_streamed(flow, host_inputs, device, chunk_rows=1 << 16, out=None):
    """Flow.log_prob of a (pinned) HOST tensor: chunks are copied on a side stream while the previous chunk is being
    evaluated, so the host->device transfer hides behind the kernels.  Returns the [n_rows] result on `device`."""
    from . import config
    from . import kernels as K
    n = host_inputs.shape[0]
    result = out if out is not None else torch.empty(n, dtype=torch.float32, device=device)
    if n == 0:
        return result
    # One flag read for the whole call instead of one per chunk (each read would drain the GPU while the host prepares the next
    # chunk's launches); a range flag repeats the call with a smaller activation exponent, like a single Flow.log_prob does.
    saved = config.activation_exp
    try:
        while True:
            with K.deferred_flags(torch.device(device)) as deferred:
                _stream_chunks(flow, host_inputs, device, chunk_rows, result)
            v = deferred.value() if config.check_domain else 0
            if (v & 4) and config.auto_activation_exp and config.activation_exp > -24:
                config.activation_exp = max(-24, config.activation_exp - 5)
                continue
            K.raise_for_flag_value(v)
            return result
    finally:
        config.activation_exp = saved


_STAGING = {}


def _staging(device, rows, cols):
    """Two persistent device staging buffers per (device, chunk shape) and one copy stream per device: allocating a fresh buffer
    per chunk made the caching allocator fall back to cudaMalloc (a device synchronisation) whenever a block recorded on the
    other stream was not yet reusable -- end-to-end times jumped between 220 and 280 ms per 2^20 rows with the chunking."""
    key = (str(device), int(rows), int(cols))
    hit = _STAGING.get(key)
    if hit is None:
        if len(_STAGING) > 8:
            _STAGING.clear()
        hit = ([torch.empty(rows, cols, dtype=torch.float32, device=device) for _ in range(2)], torch.cuda.Stream(device=device),
               [None, None])          # buffers, copy stream, "last consumer has finished" event per buffer (kept across calls)
        _STAGING[key] = hit
    return hit


def _stream_chunks(flow, host_inputs, device, chunk_rows, result):
    n = host_inputs.shape[0]
    staging, copy_stream, ready = _staging(device, min(chunk_rows, n), host_inputs.shape[1])
    compute = torch.cuda.current_stream(device)

    # chunk boundaries: a short first chunk (its copy is the only one nothing hides) and a half one, then full chunks
    bounds, lo = [], 0
    ramp = os.environ.get("NFLOWS_B200_STREAM_RAMP", "1") != "0"
    for size in (max(4096, chunk_rows // 8), max(4096, chunk_rows // 2)):
        if ramp and n - lo > chunk_rows:
            bounds.append((lo, lo + size))
            lo += size
    while lo < n:
        bounds.append((lo, min(n, lo + chunk_rows)))
        lo = bounds[-1][1]

    def stage(i, k):
        lo, hi = bounds[k]
        with torch.cuda.stream(copy_stream):
            if ready[i] is not None:
                copy_stream.wait_event(ready[i])          # previous consumer of this buffer has finished
            staging[i][:hi - lo].copy_(host_inputs[lo:hi], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return ev

    pending = (0, stage(0, 0))
    for k, (lo, hi) in enumerate(bounds):
        i, ev = pending
        if k + 1 < len(bounds):
            j = (k + 1) & 1
            pending = (j, stage(j, k + 1))
        compute.wait_event(ev)
        result[lo:hi] = flow.log_prob(staging[i][:hi - lo])
        done = torch.cuda.Event()
        done.record(compute)
        ready[i] = done
