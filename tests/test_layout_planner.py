"""Host logic of the column-layout planner (transforms/fused_affine.py, CompositeTransform._native_apply): folded affine
operands for permuted input / output column orders, and the layout a fused coupling asks for.  CPU only."""
import numpy as np
import torch

from nflows_b200 import transforms as T
from nflows_b200.nn.nets import ResidualNet
from nflows_b200.transforms.fused_affine import AffineRun, Layout
from nflows_b200.utils import torchutils


def test_layout_index_maps_are_inverse_permutations():
    lay = Layout([3, 0, 2, 1])
    x = torch.arange(20.0).reshape(5, 4)
    phys = x[:, lay.cols("cpu").long()]
    assert torch.equal(phys[:, 0], x[:, 3]) and torch.equal(phys[:, 3], x[:, 1])
    assert torch.equal(phys[:, lay.cols("cpu", inverse=True).long()], x)


@torch.no_grad()
def test_affine_run_operands_absorb_input_and_output_layouts():
    torch.manual_seed(0)
    d = 12
    leaves = [T.ActNorm(d), T.RandomPermutation(d), T.LULinear(d, identity_init=False)]
    leaves[0].log_scale.normal_(0, 0.3); leaves[0].shift.normal_(); leaves[0].initialized.fill_(1)
    for p in leaves[2].parameters():
        p.add_(torch.randn_like(p) * 0.2)
    x = torch.randn(7, d, dtype=torch.float64)
    want = x
    for t in leaves:
        want = t.double()(want)[0]
    for t in leaves:
        t.float()
    rng = np.random.default_rng(1)
    lin, lout = Layout(rng.permutation(d)), Layout(rng.permutation(d))
    for inverse in (False, True):
        seq = [(t, inverse) for t in (reversed(leaves) if inverse else leaves)]
        run = AffineRun(seq, "cpu")
        if inverse:
            src, ref = want, x
        else:
            src, ref = x, want
        for li, lo in ((None, None), (lin, None), (None, lout), (lin, lout)):
            w, b = run.operands(li, lo)
            xin = src if li is None else src[:, torch.from_numpy(li.perm)]
            got = xin @ w.double().t() + b.double()
            exp = ref if lo is None else ref[:, torch.from_numpy(lo.perm)]
            assert float((got - exp).abs().max()) <= 5e-6, (inverse, li is not None, lo is not None)
        assert run.operands(lin, lout)[0] is run.operands(lin, lout)[0]          # cached per layout pair


def test_fused_coupling_asks_for_identity_first_layout_only_when_tma_addressable(monkeypatch):
    def make(d, hidden=16):
        return T.PiecewiseRationalQuadraticCouplingTransform(
            torchutils.create_alternating_binary_mask(d), lambda i, o: ResidualNet(i, o, hidden_features=hidden, num_blocks=1),
            num_bins=8, tails="linear", tail_bound=3.0).eval()
    t = make(32)
    # pretend the tensor is something the native path takes (no GPU here): only the layout decision is under test
    monkeypatch.setattr(type(t), "_native_ready", lambda self, inputs, context: True)
    monkeypatch.setattr("nflows_b200.kernels.f16x3_supported", lambda lda, ldw, k: k % 8 == 0 and lda % 8 == 0 and ldw % 8 == 0)
    monkeypatch.setattr("nflows_b200.kernels.rq_coupling_final_supported", lambda bins, tails, hidden, lda: hidden % 8 == 0)
    x = torch.zeros(4, 32)
    lay = t._native_layout(x, None)
    assert lay is not None and lay is t._native_layout(x, None)                  # cached object: identity-keyed downstream
    assert lay.perm.tolist() == t.identity_features.tolist() + t.transform_features.tolist()
    assert make(24)._native_layout(torch.zeros(4, 24), None) is None            # 12 identity features: not a multiple of 8
