"""Shared helpers for the parity tests: build a Learner whose parameters equal an oracle parameter
dict, and compare tensors with the tolerance stated by BASELINE.json's north_star."""
import os

import numpy as np
import torch

import gccvae_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def cfg_for(mode, frac="0.5"):
    """mode in {'one-one','inferred','learnable'} -> reference train_config (gated_ccvae.py:462-476)."""
    mu = np.load(os.path.join(GOLDEN, "data", "gating_matrix_{}.npy".format(frac)))
    base = dict(lr=1e-4, batch_size=256, init_temp=0.1, gating_reg=0.2, n_epochs=1)
    if mode == "one-one":
        return dict(base, gate_type="fixed", gate_subtype="one-one", gating_init_temp=0.3)
    if mode == "inferred":
        return dict(base, gate_type="fixed", gate_subtype="inferred", mu_init=mu, gating_init_temp=0.3)
    if mode == "learnable":
        return dict(base, gate_type="learnable", gate_subtype=None, mu_init=mu, gating_init_temp=1.0)
    raise KeyError(mode)


def make_learner(cfg, params, precision="fp32", **kw):
    import gccvae_b200 as G
    lrn = G.Learner((64, 64, 3), 45, 18, 18, 1000, 1.0, cfg, precision=precision, **kw)
    lrn.store.load_dict(params)
    return lrn


def rel_err(got, want):
    """max |got-want| / max|want|  (tensor-wise relative error; near-zero entries are judged
    against the tensor's scale, SURVEY.md §7 'fp32-mode 1e-5 on gradients')."""
    got = torch.as_tensor(got).detach().double().cpu()
    want = torch.as_tensor(want).detach().double().cpu()
    assert got.shape == want.shape, (got.shape, want.shape)
    denom = float(want.abs().max())
    if denom == 0.0:
        return float(got.abs().max())
    return float((got - want).abs().max()) / denom


def assert_close(got, want, rtol, what):
    e = rel_err(got, want)
    assert e <= rtol, "{}: rel err {:.3e} > {:.1e}".format(what, e, rtol)
    return e
