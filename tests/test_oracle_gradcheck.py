"""fp64 central finite differences of the oracle's own loss, for every parameter group incl. mu
(SURVEY.md §8c item 10).  The stop_gradient at gated_ccvae.py:280 is honoured by freezing the
detached z_c at its unperturbed value while differencing."""
import numpy as np
import pytest
import torch

import gccvae_oracle as O


def _setup(golden_dir, supervised):
    torch.manual_seed(0)
    mu0 = np.load(golden_dir + "/data/gating_matrix_0.5.npy")
    cfg = dict(gate_type="learnable", gate_subtype=None, mu_init=mu0, gating_reg=0.2)
    p = O.init_params(3, dtype=torch.float64, trained_like=True)
    mu, _ = O.initialise_mu(cfg, dtype=torch.float64)
    mu = (mu * 0.9 + 0.05)  # keep away from the clip corners so FD is two-sided
    x, y, noise = O.make_inputs(2, k=3, dtype=torch.float64)
    return cfg, p, mu, x, y, noise


@pytest.mark.parametrize("supervised", [True, False])
def test_fd_matches_autograd(golden_dir, supervised):
    cfg, p, mu, x, y, noise = _setup(golden_dir, supervised)
    T = 0.7
    out, grads = O.loss_and_grads(p, mu, x, y, noise, cfg, T, supervised)
    zc_frozen = out["z"][:, O.Z_STYLE:].clone()
    y_frozen = out.get("y")

    def loss_of(pp, mm):
        with torch.no_grad():
            if supervised:
                return float(O.sup_loss(pp, mm, x, y, noise, cfg, T, zc_detached_override=zc_frozen)["loss"])
            o = O.unsup_loss(pp, mm, x, noise, cfg, T)
            assert torch.equal(o["y"], y_frozen)  # sampled labels must not flip under the perturbation
            return float(o["loss"])

    rng = np.random.RandomState(0)
    h = 1e-6
    names = list(p.keys()) + ["mu"]
    worst = 0.0
    for name in names:
        base = mu if name == "mu" else p[name]
        g = grads[name]
        flat = base.reshape(-1)
        for idx in rng.choice(flat.numel(), size=min(4, flat.numel()), replace=False):
            old = float(flat[idx])
            vals = []
            for s in (+1, -1):
                flat[idx] = old + s * h
                vals.append(loss_of(p, mu))
            flat[idx] = old
            fd = (vals[0] - vals[1]) / (2 * h)
            an = float(g.reshape(-1)[idx])
            err = abs(fd - an) / (max(abs(fd), abs(an)) + 1e-2)  # loss ~1e4 in fp64: FD round-off ~3e-6 absolute at h=1e-6
            worst = max(worst, err)
            assert err < 2e-4, (name, int(idx), fd, an)
    assert worst < 2e-4
