"""Known-answer tests that pin the CPU oracle to the reference's cited lines.
The reference ships no tests; these are the hand-derivable answers."""
import math
import os

import numpy as np
import pytest
import torch

import gccvae_oracle as O


def _cfg(kind, mu=None):
    if kind == "one-one":
        return dict(gate_type="fixed", gate_subtype="one-one", gating_reg=0.2)
    if kind == "inferred":
        return dict(gate_type="fixed", gate_subtype="inferred", mu_init=mu, gating_reg=0.2)
    return dict(gate_type="learnable", gate_subtype=None, mu_init=mu, gating_reg=0.2)


def test_param_counts():
    p = O.init_params(0)
    n = lambda pre: sum(v.numel() for k, v in p.items() if k.startswith(pre))
    assert n("enc.") == 729690 and n("dec.") == 276249 and n("cls.") == 342 and n("prior.") == 1296
    assert O.param_count(p) == 1007577  # + 324 for a learnable mu = 1,007,901


def test_invalid_gate_type():
    with pytest.raises(ValueError, match="Invalid gate type/subtype"):
        O.initialise_mu(dict(gate_type="fixed", gate_subtype="bogus"))


@pytest.mark.parametrize("T", [0.3, 1.0])
def test_one_one_gate_is_exact_identity(T):
    mu, trainable = O.initialise_mu(_cfg("one-one"))
    assert not trainable
    g = torch.Generator().manual_seed(7)
    for _ in range(5):
        c = O.sample_gating_parameter(mu, T, torch.rand(18, 18, generator=g), torch.rand(18, 18, generator=g))
        assert torch.equal(c, torch.eye(18))


def test_one_one_classifier_is_diagonal():
    p = O.init_params(1)
    z = torch.randn(5, 18)
    logits = O.classifier(p, z.unsqueeze(-1).repeat(1, 1, 18), torch.eye(18))
    torch.testing.assert_close(logits, z * torch.diagonal(p["cls.w"]) + p["cls.b"])


def test_log_py_constant():
    y = (torch.rand(6, 18) < 0.5).long()
    lp = O.bernoulli_probs_log_prob(torch.full((6, 18), 0.5), y).sum(-1)
    torch.testing.assert_close(lp, torch.full((6,), 18 * math.log(0.5)))
    assert abs(18 * math.log(0.5) + 12.476649) < 1e-6


def test_img_log_likelihood_known():
    x = torch.rand(3, 64, 64, 3)
    torch.testing.assert_close(O.img_log_likelihood(x, x), torch.full((3,), -12288 * math.log(2.0)), rtol=1e-6, atol=0)
    r = torch.rand(3, 64, 64, 3)
    want = -(x - r).abs().double().sum((1, 2, 3)) - 12288 * math.log(2.0)
    torch.testing.assert_close(O.img_log_likelihood(r, x).double(), want, rtol=1e-5, atol=0)


def test_kl_known():
    loc, sc = torch.rand(4, 45), torch.rand(4, 45) + 0.1
    assert O.get_gaussian_kl_div(loc, sc, loc, sc).abs().max() < 1e-5
    want = (0.5 * (loc ** 2 + sc ** 2 - 1) - torch.log(sc)).sum(-1)
    torch.testing.assert_close(O.get_gaussian_kl_div(loc, sc), want, rtol=1e-5, atol=1e-5)


def test_fresh_prior_independent_of_y():
    p = O.init_params(0)
    c = torch.rand(18, 18)
    for _ in range(2):
        y = (torch.rand(4, 18) < 0.5).long()
        loc, sc = O.cond_prior(p, y.unsqueeze(-1).repeat(1, 1, 18).double(), c)
        assert torch.equal(loc, torch.zeros(4, 18))
        torch.testing.assert_close(sc, torch.nn.functional.softplus(c.sum(1)).expand(4, 18))


def test_k1_same_noise_gives_unit_weight():
    p = O.init_params(0, trained_like=True)
    x, y, noise = O.make_inputs(3, k=1)
    noise["eps_k"] = noise["eps"].unsqueeze(0).clone()
    out = O.sup_loss(p, torch.eye(18), x, y, noise, _cfg("one-one"), 0.3)
    torch.testing.assert_close(out["log_qy_x"], out["log_qy_zc"], rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(out["w"], torch.ones(3), rtol=1e-5, atol=1e-5)


def test_gate_half():
    U = torch.rand(18, 18)
    c = O.sample_gating_parameter(torch.full((18, 18), 0.5), 1.0, U, U.clone())
    torch.testing.assert_close(c, torch.full((18, 18), 0.5))


def test_learned_mu_clip(golden_dir):
    mu = torch.from_numpy(np.load(os.path.join(golden_dir, "learned", "learned_gating_matrix_1.0_best.npy")))
    assert mu.dtype == torch.float32 and mu.min() < 0 and mu.max() > 1
    mu = mu.clone().requires_grad_(True)
    U1, U2 = torch.rand(18, 18), torch.rand(18, 18)
    c = O.sample_gating_parameter(mu, 1.0, U1, U2)
    assert torch.all(c[mu.detach() < 0] == 0) and torch.all(c[mu.detach() > 1] == 1)
    (c * torch.randn(18, 18)).sum().backward()
    outside = (mu.detach() < 0) | (mu.detach() > 1)
    assert outside.any() and torch.all(mu.grad[outside] == 0)


@pytest.mark.parametrize("frac", ["0.0", "0.1", "0.2", "0.5", "1.0"])
def test_gating_matrix_fixtures(golden_dir, frac):
    a = np.load(os.path.join(golden_dir, "data", "gating_matrix_{}.npy".format(frac)))
    assert a.dtype == np.float64 and a.shape == (18, 18)
    assert np.array_equal(np.diag(a), np.ones(18)) and np.allclose(a, a.T)
    mu, trainable = O.initialise_mu(_cfg("inferred", a))
    assert not trainable and mu.dtype == torch.float32
    assert np.array_equal(mu.numpy(), a.astype(np.float32))  # row=i=z, col=j=y, no transpose
    if frac == "0.0":
        assert np.all(a[~np.eye(18, dtype=bool)] == 0.5)


def test_one_one_offdiag_grads_exactly_zero():
    p = O.init_params(0, trained_like=True)
    x, y, noise = O.make_inputs(3, k=4)
    cfg = _cfg("one-one")
    mu, _ = O.initialise_mu(cfg)
    _, g = O.loss_and_grads(p, mu, x, y, noise, cfg, 0.3, True)
    off = ~torch.eye(18, dtype=torch.bool)
    assert g["mu"] is None
    for k in ("cls.w", "prior.loc_true", "prior.loc_false", "prior.scale_true", "prior.scale_false"):
        assert torch.all(g[k][off] == 0) and g[k][~off].abs().max() > 0


def test_unsup_labels_follow_uniforms():
    p = O.init_params(0, trained_like=True)
    x, _, noise = O.make_inputs(4, k=1)
    out = O.unsup_loss(p, torch.eye(18), x, noise, _cfg("one-one"), 0.3)
    assert out["y"].dtype == torch.int32
    assert torch.equal(out["y"].bool(), noise["U_y"] < torch.sigmoid(out["logits"]))


def test_keras_adam_first_step():
    opt = O.KerasAdam(1e-4)
    p = {"a": torch.tensor([1.0, -2.0])}
    opt.apply(p, {"a": torch.tensor([0.5, -0.25])})
    # t=1: m=(1-b1)g, v=(1-b2)g^2, lr_t = lr*sqrt(1-b2)/(1-b1)  ->  step ~= lr * sign(g)
    lr_t = 1e-4 * math.sqrt(1 - 0.999) / (1 - 0.9)
    g = torch.tensor([0.5, -0.25])
    want = torch.tensor([1.0, -2.0]) - lr_t * (0.1 * g) / (torch.sqrt(0.001 * g * g) + 1e-7)
    torch.testing.assert_close(p["a"], want)
