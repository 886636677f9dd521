"""Host-side logic of the native path that needs no GPU: how conditioner networks are described to the kernels."""
import torch


def test_context_chain_description_matches_the_gated_resnet():
    """dense.Chain of a context-conditioned ResidualNet (reference nn/nets/resnet.py:36-100), interpreted with
    torch ops on the CPU, reproduces the module's own forward: the tokens "ctx_init" / "glu_skip", the zero-padded context
    weights and the row slicing are what the native runner (dense._run_trunk_block) executes."""
    import torch.nn.functional as F
    from nflows_b200 import dense as D
    from nflows_b200.nn.nets import ResidualNet
    torch.manual_seed(3)
    net = ResidualNet(8, 12, hidden_features=32, context_features=6, num_blocks=2).eval()
    with torch.no_grad():
        for p in net.parameters():
            p.add_(torch.randn_like(p) * 0.1)
        x, c = torch.randn(50, 8), torch.randn(50, 6)
        assert net.dense_chain(None) is None and ResidualNet(8, 12, 32).dense_chain(c) is None      # context must match the net
        chain = net.dense_chain(c)
        assert isinstance(chain, D.Chain) and chain.ctx_pad == 8 and [l[4] for l in chain] == ["ctx_init", None, "glu_skip", None, "glu_skip", None]
        assert net.dense_chain(c)[0][0] is chain[0][0]          # persistent slices: the split-pair cache keys on identity

        def run(ch):
            ctx = F.pad(ch.context, (0, ch.ctx_pad - ch.context.shape[1]))
            h, skip, branch = x[10:30], None, None
            for i, (w, b, relu_in, relu_out, res) in enumerate(ch[:-1]):
                a = F.relu(h) if relu_in else h
                y = F.linear(a, w, b)
                if res == "ctx_init":
                    y = y + F.linear(ctx, ch.ctx_init[0], ch.ctx_init[1])
                    assert torch.equal(ch.ctx_init[0][:, :6], ch.ctx_init[2])
                elif res == "glu_skip":
                    y = skip + y * torch.sigmoid(F.linear(ctx, ch.ctx_gates[i][0], ch.ctx_gates[i][1]))
                if relu_out:
                    y = F.relu(y)
                if i + 2 < len(ch) and ch[i + 2][4] == "glu_skip":
                    skip = y
                h = y
            return F.linear(h, ch[-1][0], ch[-1][1])

        got = run(chain.rows(10, 30))
        assert torch.allclose(got, net(x[10:30], c[10:30]), atol=1e-5)
        net.initial_layer.weight.mul_(2.0)                      # in-place update: the slices are rebuilt
        assert not torch.equal(net.dense_chain(c)[0][0], chain[0][0])
