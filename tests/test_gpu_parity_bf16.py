"""GPU parity of the bf16 tensor-core path against the fp64 CPU oracle.

Tolerance (BASELINE.json north_star): rtol 1e-2 in bf16 mode.
 * Forward quantities (every per-image loss term, the loss) are compared with the PLAIN fp64 oracle,
   element-wise relative to the tensor's max.
 * Gradients are compared, per parameter tensor in the relative L2 norm, with the fp64 oracle run
   with bf16 rounding at the same tensor boundaries as the kernels (gccvae_oracle.bf16_emulation):
   the Laplace likelihood's gradient is sign(x - xhat), discontinuous in the forward pass, so ANY
   1e-4 perturbation of xhat flips ~3e-4 of the pixel signs = 2*sqrt(3e-4) ~ 3.5% L2 noise on the
   logit gradient.  The deviation from the un-rounded oracle is also reported and bounded (15%)."""
import numpy as np
import pytest
import torch

import gccvae_oracle as O
from helpers import assert_close, cfg_for, make_learner, rel_err

pytestmark = pytest.mark.gpu
RTOL = 1e-2


def l2_rel(got, want):
    got = torch.as_tensor(got).detach().double().cpu()
    want = torch.as_tensor(want).detach().double().cpu()
    return float((got - want).norm() / want.norm().clamp_min(1e-30))


@pytest.mark.parametrize("mode,supervised,B,K", [
    ("one-one", True, 200, 100),
    ("inferred", True, 32, 100),
    ("inferred", False, 32, 100),
    ("learnable", True, 32, 100),
    ("learnable", False, 32, 100),
    ("learnable", True, 3, 5),
])
def test_bf16_step_matches_oracle(mode, supervised, B, K):
    cfg = cfg_for(mode, "0.2")
    T = cfg["gating_init_temp"]
    p32 = O.init_params(0, trained_like=True)
    p64 = {k: v.double() for k, v in p32.items()}
    mu32, _ = O.initialise_mu(cfg)
    x, y, noise = O.make_inputs(B, k=K)
    n64 = {k: v.double() for k, v in noise.items()}
    o64, g64 = O.loss_and_grads(p64, mu32.double(), x.double(), y, n64, cfg, T, supervised)
    with O.bf16_emulation():
        oem, gem = O.loss_and_grads(p64, mu32.double(), x.double(), y, n64, cfg, T, supervised)
    lrn = make_learner(cfg, p32, precision="bf16")
    loss, c = lrn.loss_and_grads(x, y, supervised, noise=noise, k=K)
    torch.cuda.synchronize()
    last = lrn.last
    if mode == "one-one":
        assert torch.equal(c.cpu(), torch.eye(18))
    report = {}
    for k in ["post_locs", "post_scales", "z", "kl", "log_pxz", "recon"]:
        report[k] = assert_close(last[k], o64[k], RTOL, k)
    if supervised:
        # label-dependent terms need identical labels; unsup labels are sampled from the bf16 logits and may
        # legitimately flip where U_y is within rounding of sigmoid(logit)
        for k in ["logits", "log_qy_zc", "log_qy_x", "w"]:
            report[k] = assert_close(last[k], o64[k], RTOL, k)
        report["loss"] = assert_close(loss, o64["loss"], RTOL, "loss")
    else:
        flips = int((last["y"].cpu() != o64["y"]).sum())
        assert flips <= max(1, B * 18 // 100), "too many sampled-label flips: {}".format(flips)
        if flips == 0:
            report["loss"] = assert_close(loss, o64["loss"], RTOL, "loss")
    assert_close(last["recon"], oem["recon"], 1e-3, "recon vs bf16-emulating oracle")
    worst, worst_plain = ("", 0.0), ("", 0.0)
    y_same = supervised or (int((last["y"].cpu() != oem["y"]).sum()) == 0 and int((oem["y"] != o64["y"]).sum()) == 0)
    for name in lrn.store.names():
        if name == "mu" and not lrn.model.mu_trainable:
            continue
        if not y_same and (name.startswith("prior") or name.startswith("cls") or name == "mu"):
            continue
        e = l2_rel(lrn.store.g(name), gem[name])
        ep = l2_rel(lrn.store.g(name), g64[name])
        worst = max(worst, (name, e), key=lambda t: t[1])
        worst_plain = max(worst_plain, (name, ep), key=lambda t: t[1])
        report["g:" + name] = e
    print("bf16 report:", {k: float("%.2e" % v) for k, v in report.items()})
    print("worst vs emulating oracle", worst, "| worst vs plain oracle", worst_plain)
    # residual: fp32-vs-fp64 accumulation moves a few bf16 roundings by one ulp, which still flips
    # ~1e-5..1e-4 of the sign(x - xhat) factors: 0.6% L2 on the encoder, up to 2% on the innermost
    # decoder layers (fc1 / conv1t see the least averaging); tiny batches average less still.
    # (B = 32: a single flipped sign among 393k pixels is already 0.3% of the logit gradient, so the bound is 5%)
    bound = (3 if B >= 64 else 5) * RTOL
    assert worst[1] <= bound, "gradient {} relative L2 error {:.3e} vs bf16-emulating oracle".format(*worst)
    assert worst_plain[1] <= 0.15, "gradient {} relative L2 error {:.3e} vs plain oracle".format(*worst_plain)


def test_bf16_train_step_runs_and_learns():
    """30 Adam steps on one fixed batch with fixed noise: the (deterministic) unsupervised loss must fall."""
    cfg = dict(cfg_for("learnable", "0.5"), lr=1e-3)
    lrn = make_learner(cfg, O.init_params(0), precision="bf16")
    x, y, noise = O.make_inputs(64, k=20)
    x = 0.8 + 0.1 * x          # learnable structure (uniform noise images cannot be fitted at all)
    first, _ = lrn.unsup_loss(x, noise=noise)
    for it in range(30):
        lrn.train_step(x, y, it % 2 == 0, noise=noise, k=20)
    last, _ = lrn.unsup_loss(x, noise=noise)
    torch.cuda.synchronize()
    assert torch.isfinite(lrn.store.flat).all()
    assert float(last) < float(first) - 30.0, (float(first), float(last))
