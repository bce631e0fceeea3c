"""GPU, BASELINE.json's full size (configs[1]: fixed-inferred gating, bf16, batch 1024, K = 100): properties that do
not need the oracle at that size - the loss is the stated combination of the per-image terms (gated_ccvae.py:225-230,
291-298), known answers hold for every image, and an image's terms do not depend on the batch it is computed in."""
import math

import pytest
import torch

import gccvae_oracle as O
from helpers import assert_close, cfg_for, make_learner

pytestmark = pytest.mark.gpu
B, K = 1024, 100
TERMS = ["post_locs", "post_scales", "z", "kl", "log_qy_zc", "log_qy_x", "w", "log_py", "log_pxz"]


@pytest.fixture(scope="module")
def setup():
    cfg = cfg_for("inferred", "0.2")
    lrn = make_learner(cfg, O.init_params(0, trained_like=True), precision="bf16")
    x, y, noise = O.make_inputs(B, k=K)
    return lrn, x, y, noise


def _terms(lrn):
    torch.cuda.synchronize()
    return {k: lrn.last[k].detach().double().cpu().clone() for k in TERMS if lrn.last.get(k) is not None}


def test_supervised_loss_is_the_stated_combination_of_the_per_image_terms(setup):
    lrn, x, y, noise = setup
    loss, c = lrn.sup_loss(x, y, noise=noise, k=K)
    t = _terms(lrn)
    assert all(bool(torch.isfinite(v).all()) for v in t.values())
    want = -(t["w"] * (t["log_pxz"] - t["kl"] - t["log_qy_zc"]) + t["log_py"] + t["log_qy_x"]).mean()
    assert abs(float(loss) - float(want)) <= 1e-5 * abs(float(want))
    assert_close(t["log_py"], torch.full((B,), 18 * math.log(0.5), dtype=torch.float64), 1e-6, "log p(y)")
    assert float(t["kl"].min()) >= 0.0 and float(t["post_scales"].min()) >= 1e-3
    assert float(t["log_pxz"].max()) <= -12288 * math.log(2.0) + 1e-3      # -|x - xhat|_1 - 12288 ln 2
    assert c.shape == (18, 18) and float(c.min()) >= 0.0 and float(c.max()) <= 1.0


def test_unsupervised_loss_is_the_stated_combination_of_the_per_image_terms(setup):
    lrn, x, y, noise = setup
    loss, _ = lrn.unsup_loss(x, noise=noise)
    t = _terms(lrn)
    want = -(t["log_pxz"] + t["log_py"] - t["kl"] - t["log_qy_zc"]).mean()
    assert abs(float(loss) - float(want)) <= 1e-5 * abs(float(want))


def test_an_images_terms_do_not_depend_on_its_batch(setup):
    lrn, x, y, noise = setup
    lrn.sup_loss(x, y, noise=noise, k=K)
    full = _terms(lrn)
    h = B // 4                                     # 256 images: other tile counts, grids and split factors
    part_noise = dict(eps=noise["eps"][:h], eps_k=noise["eps_k"][:, :h], U_y=noise["U_y"][:h], U1=noise["U1"], U2=noise["U2"])
    lrn.sup_loss(x[:h], y[:h], noise=part_noise, k=K)
    part = _terms(lrn)
    for k in TERMS:
        # per-image sums over pixels are accumulated with atomics (log_pxz): order-dependent in the last bits only
        assert_close(part[k], full[k][:h], 2e-6, k)
