"""The public classes, run on CPU (device-agnostic torch path), against the reference's golden outputs.
Checks constructor kwargs, state_dict key compatibility (load_state_dict of the REFERENCE's state_dict) and the
host-side composition logic.  The CUDA kernels are checked in test_native_parity.py (-m gpu)."""
import pytest
import torch

from conftest import load_golden, rel_err
from nflows_b200 import transforms as T
from nflows_b200.distributions import StandardNormal
from nflows_b200.flows import Flow
from nflows_b200.flows import recipes
from nflows_b200.nn.nets import ResidualNet
from nflows_b200.utils import torchutils

TOL = 2e-6


@torch.no_grad()
def test_cfg1_affine_flow_cpu():
    g = load_golden("cfg1_affine")
    flow = recipes.affine_flow_2d().eval()
    flow.load_state_dict(g["sd"])            # strict: same keys as the reference
    assert rel_err(flow.log_prob(g["x"]), g["log_prob"]) <= TOL
    z, lad = flow._transform(g["x"])
    assert rel_err(z, g["z"]) <= TOL and rel_err(lad, g["lad"]) <= TOL
    xr, ladr = flow._transform.inverse(g["z"])
    assert rel_err(xr, g["x_roundtrip"]) <= TOL and rel_err(ladr, g["lad_inverse"]) <= TOL
    assert flow.sample(5).shape == (5, 2)
    s, lp = flow.sample_and_log_prob(7)
    assert rel_err(lp, flow.log_prob(s)) <= 1e-5


@torch.no_grad()
def test_affine_variants_cpu():
    g = load_golden("affine_variants")
    f = lambda i, o: ResidualNet(i, o, hidden_features=16)
    mask = torchutils.create_mid_split_binary_mask(10)
    tg = T.AffineCouplingTransform(mask, f, scale_activation=T.AffineCouplingTransform.GENERAL_SCALE_ACTIVATION).eval()
    tg.load_state_dict(g["sd_general"])
    y, l = tg(g["x"])
    assert rel_err(y, g["y_general"]) <= TOL and rel_err(l, g["lad_general"]) <= TOL
    ta = T.AdditiveCouplingTransform(mask, f).eval()
    ta.load_state_dict(g["sd_additive"])
    y, l = ta(g["x"])
    assert rel_err(y, g["y_additive"]) <= TOL and torch.equal(l, torch.zeros_like(l))


@torch.no_grad()
def test_cfg2_rq_coupling_cpu():
    g = load_golden("cfg2_rq_coupling")
    t = recipes.rq_coupling_layer().eval()
    t.load_state_dict(g["sd"])
    y, l = t(g["x"])
    assert rel_err(y, g["y"]) <= TOL and rel_err(l, g["lad"]) <= 1e-5
    assert torch.equal(y[:, t.identity_features], g["x"][:, t.identity_features])
    xi, li = t.inverse(g["x"])
    assert rel_err(xi, g["xinv"]) <= TOL and rel_err(li, g["ladinv"]) <= 1e-5


@torch.no_grad()
def test_rq_constrained_cpu():
    g = load_golden("rq_coupling_constrained")
    t = T.PiecewiseRationalQuadraticCouplingTransform(
        mask=torchutils.create_mid_split_binary_mask(11),
        transform_net_create_fn=lambda i, o: ResidualNet(i, o, hidden_features=24, num_blocks=1),
        num_bins=5, tails=None, min_bin_width=2e-3, min_bin_height=3e-3, min_derivative=4e-3).eval()
    t.load_state_dict(g["sd"])
    y, l = t(g["x"])
    assert rel_err(y, g["y"]) <= TOL and rel_err(l, g["lad"]) <= 1e-5
    with pytest.raises(T.InputOutsideDomain):
        t(g["x"] + 1.0)


@torch.no_grad()
def test_linear_transforms_cpu():
    g = load_golden("linear_transforms")
    d = g["x"].shape[1]
    an, lu, pm = T.ActNorm(d).eval(), T.LULinear(d, identity_init=False).eval(), T.RandomPermutation(d).eval()
    an.load_state_dict(g["sd_actnorm"]); lu.load_state_dict(g["sd_lu"]); pm.load_state_dict(g["sd_perm"])
    for name, m in (("actnorm", an), ("lu", lu), ("perm", pm)):
        y, l = m(g["x"])
        assert rel_err(y, g[name + "_y"]) <= TOL and rel_err(l, g[name + "_lad"]) <= TOL
        y, l = m.inverse(g["x"])
        assert rel_err(y, g[name + "_xinv"]) <= 1e-5 and rel_err(l, g[name + "_ladinv"]) <= TOL
    assert torch.equal(pm(g["x"])[0], g["x"][:, g["sd_perm"]["_permutation"]])
    assert rel_err(lu.weight(), g["lu_weight"]) <= TOL
    assert rel_err(lu.weight_inverse(), g["lu_weight_inverse"]) <= 1e-5
    # cache protocol (reference linear.py:46-96)
    lu.use_cache(True)
    y1, _ = lu(g["x"])
    assert lu.cache.weight is not None and rel_err(y1, g["lu_y"]) <= 1e-5
    lu.train()
    assert lu.cache.weight is None


@torch.no_grad()
def test_nsf_small_cpu():
    g = load_golden("nsf_small")
    flow = recipes.rq_nsf(g["features"], g["hidden"], g["layers"]).eval()
    flow.load_state_dict(g["sd"])
    assert rel_err(flow.log_prob(g["x"]), g["log_prob"]) <= TOL
    xs, lads = flow._transform.inverse(g["noise"])
    assert rel_err(xs, g["sample"]) <= 1e-5 and rel_err(lads, g["lad_inverse"]) <= 1e-5
    assert rel_err(flow.transform_to_noise(g["x"]), g["z"]) <= TOL


@torch.no_grad()
def test_seeded_construction_matches_reference_weights():
    """Constructors consume the RNG like the reference's: the seed-only fixtures depend on it."""
    g = load_golden("cfg2_rq_coupling")
    torch.manual_seed(0)
    t = recipes.rq_coupling_layer()
    for k, v in g["sd"].items():
        assert torch.equal(t.state_dict()[k], v), k
    g = load_golden("nsf784_layer")
    torch.manual_seed(g["seed"])
    flow = recipes.perturb_(recipes.rq_nsf(g["features"], g["hidden"], g["layers"]).eval(), g["perturb_seed"])
    ck = float(sum(v.double().abs().sum() for v in flow.state_dict().values() if v.is_floating_point()))
    assert abs(ck - g["checksum"]) <= 1e-9 * abs(g["checksum"])
    assert rel_err(flow.log_prob(g["x"]), g["log_prob"]) <= TOL


@torch.no_grad()
def test_autoregressive_rq_cpu_seeded():
    g = load_golden("ar_rq")
    torch.manual_seed(g["seed"])
    ar = T.MaskedPiecewiseRationalQuadraticAutoregressiveTransform(features=64, hidden_features=256, num_bins=8, tails="linear",
                                                                   tail_bound=3.0, num_blocks=2).eval()
    for name, p in ar.named_parameters():
        if "final_layer" in name:
            p.mul_(g["final_scale"])
    y, lad = ar(g["x"])
    assert rel_err(y, g["y"]) <= TOL and rel_err(lad, g["lad"]) <= TOL
    g2 = load_golden("ar_rq_small")
    ar2 = T.MaskedPiecewiseRationalQuadraticAutoregressiveTransform(features=6, hidden_features=16, num_bins=4, tails=None,
                                                                    num_blocks=1).eval()
    ar2.load_state_dict(g2["sd"])
    xi, li = ar2.inverse(g2["x"])
    assert rel_err(xi, g2["xinv"]) <= TOL and rel_err(li, g2["ladinv"]) <= TOL


@torch.no_grad()
def test_next_rows_cpu_replay_of_reference_goldens():
    """tests/golden/next_rows.pt (reference outputs, oracle/make_golden.py next_rows) through the package's CPU path: the
    reference's state_dicts load with strict=True and the outputs agree."""
    import warnings

    from nflows_b200.flows import SimpleRealNVP
    from nflows_b200.nn.nets import MLP, ResidualNet
    from nflows_b200.utils import torchutils
    g = load_golden("next_rows")
    for key, channels in (("conv1x1", 3), ("conv1x1_c12", 12)):
        r = g[key]
        conv = T.OneByOneConvolution(channels, identity_init=False).eval()
        conv.load_state_dict(r["sd"], strict=True)
        y, lad = conv(r["x"])
        assert rel_err(y, r["y"]) <= TOL and rel_err(lad, r["lad"]) <= TOL
    for tails in (None, "linear"):
        r = g["rq_cdf_%s" % (tails or "none")]
        cdf = T.PiecewiseRationalQuadraticCDF(shape=[7], num_bins=6, tails=tails, tail_bound=2.0).eval()
        cdf.load_state_dict(r["sd"], strict=True)
        y, lad = cdf(r["x"])
        xi, li = cdf.inverse(r["inv_in"])
        assert rel_err(y, r["y"]) <= TOL and rel_err(lad, r["lad"]) <= TOL and rel_err(xi, r["xinv"]) <= TOL
    r = g["rq_coupling_unconditional"]
    t = T.PiecewiseRationalQuadraticCouplingTransform(
        torchutils.create_alternating_binary_mask(16), lambda i, o: ResidualNet(i, o, hidden_features=32, num_blocks=1),
        num_bins=8, tails="linear", tail_bound=3.0, apply_unconditional_transform=True).eval()
    t.load_state_dict(r["sd"], strict=True)
    y, lad = t(r["x"])
    assert rel_err(y, r["y"]) <= TOL and rel_err(lad, r["lad"]) <= TOL
    r = g["simple_realnvp"]
    flow = SimpleRealNVP(features=10, hidden_features=16, num_layers=3, num_blocks_per_layer=2).eval()
    flow.load_state_dict(r["sd"], strict=True)
    assert rel_err(flow.log_prob(r["x"]), r["log_prob"]) <= TOL
    r = g["rq_coupling_mlp"]
    t = T.PiecewiseRationalQuadraticCouplingTransform(
        torchutils.create_alternating_binary_mask(32), lambda i, o: MLP([i], [o], [64, 64, 64]),
        num_bins=8, tails="linear", tail_bound=3.0).eval()
    t.load_state_dict({k.replace("transform_net.mlp.", "transform_net."): v for k, v in r["sd"].items()}, strict=True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        y, lad = t(r["x"])
    assert rel_err(y, r["y"]) <= TOL and rel_err(lad, r["lad"]) <= TOL


@torch.no_grad()
def test_image_flow_cpu_replay_of_reference_golden():
    """tests/golden/image_rows.pt (BASELINE cfg 5 in small, reference outputs): recipes.glow_multiscale has the reference's module
    tree (strict state_dict load) and the package's CPU path reproduces log_prob, the noise and the inverse."""
    from nflows_b200.flows import recipes
    g = load_golden("image_rows")["glow_small"]
    flow = recipes.glow_multiscale(image_shape=(3, 16, 16), levels=3, steps=2, hidden_channels=32).eval()
    flow.load_state_dict(g["sd"], strict=True)
    assert rel_err(flow.log_prob(g["x"]), g["log_prob"]) <= TOL
    assert rel_err(flow.transform_to_noise(g["x"]), g["z"]) <= TOL
    xs, lad = flow._transform.inverse(g["noise"])
    assert rel_err(xs, g["sample"]) <= 1e-4 and rel_err(lad, g["lad_inv"]) <= 1e-4
