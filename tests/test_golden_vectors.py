"""Committed golden vectors (tests/golden/oracle_golden.npz, made by tests/golden/make_golden.py).
CPU: the oracle still reproduces them.  GPU: the fp32 CUDA path matches them (rtol 1e-5 of the tensor's scale)."""
import os

import numpy as np
import pytest
import torch

import gccvae_oracle as O
from helpers import GOLDEN, cfg_for, make_learner

import importlib.util

_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLDEN, "make_golden.py"))
MG = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(MG)
G = np.load(os.path.join(GOLDEN, "oracle_golden.npz"))


def _rel(got, want):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    d = np.abs(want).max()
    return float(np.abs(got - want).max() / (d if d > 0 else 1.0))


@pytest.mark.parametrize("case", MG.CASES, ids=[c[0] for c in MG.CASES])
def test_oracle_reproduces_golden(case):
    name, mode, frac, sup, B, K, T = case
    o, g = MG.run_case(mode, frac, sup, B, K, T)
    for t in MG.TERMS + (["log_qy_x", "w"] if sup else ["y"]):
        assert _rel(o[t].numpy(), G["{}/{}".format(name, t)]) < 1e-12, t
    for k, v in g.items():
        if v is None:
            continue
        idx = G["{}/grad_idx/{}".format(name, k)]
        assert _rel(v.reshape(-1).numpy()[idx], G["{}/grad_val/{}".format(name, k)]) < 1e-10, k


@pytest.mark.gpu
@pytest.mark.parametrize("case", MG.CASES, ids=[c[0] for c in MG.CASES])
def test_cuda_fp32_matches_golden(case):
    name, mode, frac, sup, B, K, T = case
    cfg = cfg_for(mode, frac)
    lrn = make_learner(cfg, O.init_params(0, trained_like=True))
    lrn.gating_sampler_temp = T
    x, y, noise = O.make_inputs(B, k=K)
    loss, c = lrn.loss_and_grads(x, y, sup, noise=noise, k=K)
    torch.cuda.synchronize()
    last = lrn.last
    assert _rel(loss.cpu().numpy(), G[name + "/loss"]) < 1e-5
    assert _rel(c.cpu().numpy(), G[name + "/c"]) < 2e-5
    for t in ["post_locs", "post_scales", "z", "logits", "log_qy_zc", "log_py", "kl", "log_pxz"]:
        assert _rel(last[t].cpu().numpy(), G["{}/{}".format(name, t)]) < 2e-5, t
    if sup:
        assert _rel(last["log_qy_x"].cpu().numpy(), G[name + "/log_qy_x"]) < 2e-5
        assert _rel(last["w"].cpu().numpy(), G[name + "/w"]) < 5e-5
    else:
        assert np.array_equal(last["y"].cpu().numpy(), G[name + "/y"])
    for k in lrn.store.names():
        key = "{}/grad_idx/{}".format(name, k)
        if key not in G.files:
            continue
        got = lrn.store.g(k).cpu().numpy().reshape(-1)
        # probes are judged against the tensor's own scale (its norm / sqrt(n))
        scale = float(G["{}/grad_norm/{}".format(name, k)]) / np.sqrt(got.size)
        err = np.abs(got[G[key]] - G["{}/grad_val/{}".format(name, k)]).max()
        assert err < 2e-4 * max(scale, 1e-30) + 1e-5 * np.abs(G["{}/grad_val/{}".format(name, k)]).max(), (k, err, scale)
