"""EXPERIMENTAL persistent trunk kernel (nfk_residual_trunk_f16x3).  Host planning is tested on the CPU; the GPU comparison
runs only with NFLOWS_B200_TRUNK_KERNEL=1 (the kernel had no hardware time in round 1 -- see DESIGN.md section 8)."""
import os

import pytest
import torch

from conftest import rel_err
from nflows_b200 import dense as D
from nflows_b200.nn.nets import MLP, ResidualNet


def test_plan_follows_the_residual_block_structure(monkeypatch):
    monkeypatch.setattr("nflows_b200.kernels.residual_trunk_supported", lambda h, n, lda, ldw: h % 32 == 0 and h <= 256 and n <= 8)
    net = ResidualNet(24, 48, hidden_features=64, num_blocks=2).eval()
    flags = D.plan_trunk_kernel(net.dense_chain(None))
    # block = [relu -> W -> relu] [W + skip]; the first block's output is the second block's skip tensor
    assert flags == [1, 2 | 4 | 8, 1, 2]
    assert D.plan_trunk_kernel(ResidualNet(24, 48, hidden_features=40, num_blocks=2).eval().dense_chain(None)) is None   # 40 % 32
    assert D.plan_trunk_kernel(MLP([24], [48], [64]).eval().dense_chain(None)) is None                                   # too short
    mlp = MLP([24], [48], [64, 64, 64, 64]).eval()
    assert D.plan_trunk_kernel(mlp.dense_chain(None)) == [1, 1, 1]      # relu outputs, no skips, consumers take them as they are


@pytest.mark.gpu
@pytest.mark.skipif(os.environ.get("NFLOWS_B200_TRUNK_KERNEL", "0") != "1", reason="experimental kernel: set NFLOWS_B200_TRUNK_KERNEL=1")
@torch.no_grad()
@pytest.mark.parametrize("hidden,blocks,rows", [(256, 2, 1000), (64, 1, 130), (128, 3, 4096)])
def test_trunk_kernel_matches_the_layer_by_layer_path(cuda_device, hidden, blocks, rows, monkeypatch):
    torch.manual_seed(hidden + blocks)
    net = ResidualNet(40, 16, hidden_features=hidden, num_blocks=blocks).eval()
    for p in net.parameters():
        p.add_(torch.randn_like(p) * 0.05)
    x = torch.randn(rows, 40)
    # fp64 value of the trunk output (input of the final layer)
    h = net.double().initial_layer(x.double())
    for block in net.blocks:
        h = block(h)
    net = net.float().to(cuda_device)
    chain = net.dense_chain(None)
    xd = x.to(cuda_device)
    got = D.run_trunk(chain, xd, None, True).pair.float().cpu()
    monkeypatch.setenv("NFLOWS_B200_TRUNK_KERNEL", "0")
    base = D.run_trunk(chain, xd, None, True).pair.float().cpu()
    assert rel_err(got, h) <= 1e-5 and rel_err(base, h) <= 1e-5
    assert rel_err(got, base) <= 2e-6
