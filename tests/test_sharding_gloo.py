"""Multi-process row sharding on CPU (gloo, world_size 2): shard bounds, ragged gather order, equality with the
single-process result.  The flow runs on its device-agnostic torch path here; the GPU kernels are covered by -m gpu."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nflows_b200 import sharding
from nflows_b200.flows import recipes


def test_shard_bounds_cover_rows_exactly():
    for n in (0, 1, 7, 8, 1000, 1 << 20):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_rows, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        flow = recipes.perturb_(recipes.rq_nsf(features=12, hidden_features=16, num_layers=2).eval())
        x = torch.randn(n_rows, 12, generator=torch.Generator().manual_seed(1))
        full = sharding.log_prob_sharded(flow, x)
        local = sharding.log_prob_sharded(flow, x, gather=False)
        torch.save({"full": full, "local": local, "bounds": sharding.shard_bounds(n_rows, world, rank)},
                   os.path.join(out_dir, "rank%d.pt" % rank))
    finally:
        dist.destroy_process_group()


def test_log_prob_sharded_matches_single_process(tmp_path):
    for n_rows in (64, 37):     # equal shards (all_gather_into_tensor) and ragged shards (padded all_gather)
        port = _free_port()
        mp.spawn(_worker, args=(2, port, n_rows, str(tmp_path)), nprocs=2, join=True)
        torch.manual_seed(0)
        flow = recipes.perturb_(recipes.rq_nsf(features=12, hidden_features=16, num_layers=2).eval())
        x = torch.randn(n_rows, 12, generator=torch.Generator().manual_seed(1))
        with torch.no_grad():
            want = flow.log_prob(x)
        for rank in range(2):
            got = torch.load(os.path.join(str(tmp_path), "rank%d.pt" % rank))
            lo, hi = got["bounds"]
            assert torch.allclose(got["full"], want, rtol=0, atol=1e-5)
            assert torch.allclose(got["local"], want[lo:hi], rtol=0, atol=1e-5)
