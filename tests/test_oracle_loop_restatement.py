"""A second, independent restatement of the latent stage of the supervised loss - scalar Python loops written straight
from the reference lines, no broadcasting, no tiling - against the vectorised oracle on ASYMMETRIC gates and kernels
(the shipped co-occurrence matrices are symmetric, so a transposed index would go unnoticed with them).
Index conventions under test: c[i, j] with i = z_c dimension, j = label;
classifier kernel W[i, j]; prior kernels K[j, i]; gated_ccvae.py:167-182, 234-300; networks.py:72-74, 83-86, 104-127."""
import math

import numpy as np
import torch

import gccvae_oracle as O

ZS, ZC, Y = 27, 18, 18


def _softplus(v):
    return max(v, 0.0) + math.log1p(math.exp(-abs(v)))


def _bern_lp(logit, y):          # TFP Bernoulli(logits).log_prob(y)
    return -_softplus(-logit) if y == 1 else -_softplus(logit)


def _loop_latent(p, c, loc, scale, eps, eps_k, y):
    """-> per image: log_qy_zc, log_qy_x, kl, w (all Python floats)."""
    B, K = loc.shape[0], eps_k.shape[0]
    W, bias = p["cls.w"], p["cls.b"]
    out = []
    for b in range(B):
        z = [loc[b, d] + scale[b, d] * eps[b, d] for d in range(ZS + ZC)]            # sample_normal, :90-93
        # classifier on the tiled z_c: logits[j] = sum_i z_c[i] * c[i, j] * W[i, j] + bias[j]
        def logits_of(zc):
            return [sum(zc[i] * c[i, j] * W[i, j] for i in range(ZC)) + bias[j] for j in range(Y)]
        lg = logits_of(z[ZS:])
        log_qy_zc = sum(_bern_lp(lg[j], int(y[b, j])) for j in range(Y))
        # K-sample estimate of log q(y|x), :167-182
        rows = []
        for k in range(K):
            zk = [loc[b, ZS + i] + scale[b, ZS + i] * eps_k[k, b, ZS + i] for i in range(ZC)]
            lk = logits_of(zk)
            rows.append(sum(_bern_lp(lk[j], int(y[b, j])) for j in range(Y)))
        m = max(rows)
        log_qy_x = m + math.log(sum(math.exp(r - m) for r in rows)) - math.log(float(K))
        # conditional prior on the tiled labels with c transposed: loc_p[i] = sum_j (y_j Kt[j,i] + (1-y_j) Kf[j,i]) c[i,j]
        kl = 0.0
        for d in range(ZS + ZC):
            if d < ZS:
                lp, sp = 0.0, 1.0
            else:
                i = d - ZS
                lp = sum((p["prior.loc_true"][j, i] if y[b, j] == 1 else p["prior.loc_false"][j, i]) * c[i, j] for j in range(Y))
                raw = sum((p["prior.scale_true"][j, i] if y[b, j] == 1 else p["prior.scale_false"][j, i]) * c[i, j] for j in range(Y))
                sp = min(max(_softplus(raw), 1e-3), 1e3)
            lq, sq = loc[b, d], scale[b, d]
            dls = math.log(sq) - math.log(sp)
            kl += 0.5 * ((lq - lp) / sp) ** 2 + 0.5 * math.expm1(2.0 * dls) - dls     # TFP kl(Normal, Normal)
        out.append((log_qy_zc, log_qy_x, kl, math.exp(log_qy_zc - log_qy_x)))
    return out


def test_loop_restatement_agrees_with_the_oracle_on_asymmetric_gates():
    torch.manual_seed(0)
    B, K = 3, 4
    p = {k: v.double() for k, v in O.init_params(0, trained_like=True).items()}
    g = torch.Generator().manual_seed(11)
    for name in ("cls.w", "prior.loc_true", "prior.loc_false", "prior.scale_true", "prior.scale_false"):
        p[name] = torch.randn(p[name].shape, generator=g, dtype=torch.float64)        # asymmetric, O(1)
    c = torch.rand(ZC, Y, generator=g, dtype=torch.float64)
    assert float((c - c.t()).abs().max()) > 0.1
    x, y, noise = O.make_inputs(B, k=K, dtype=torch.float64)
    loc, scale = O.encoder(p, x)
    # the oracle's own pieces, assembled as sup_loss does (gated_ccvae.py:237-289)
    z = O.sample_normal(loc, scale, noise["eps"])
    z_t = z[:, ZS:].unsqueeze(-1).repeat(1, 1, Y)
    lq = O.bernoulli_logits_log_prob(O.classifier(p, z_t, c), y).sum(-1)
    lqx = O.classifier_loss(p, x, y, c, noise["eps_k"])
    kl, _, _ = O._prior_and_kl(p, y, c, loc, scale)
    w = torch.exp(lq - lqx)
    pn = {k: v.numpy() for k, v in p.items()}
    got = _loop_latent(pn, c.numpy(), loc.detach().numpy(), scale.detach().numpy(), noise["eps"].numpy(),
                       noise["eps_k"].numpy(), y.numpy())
    for b in range(B):
        for name, mine, theirs in zip(("log_qy_zc", "log_qy_x", "kl", "w"), got[b], (lq[b], lqx[b], kl[b], w[b])):
            assert abs(mine - float(theirs)) <= 1e-9 * max(1.0, abs(mine)), (b, name, mine, float(theirs))
    # and a transposed gate IS a different model (the test would notice an i/j mix-up)
    lq_t = O.bernoulli_logits_log_prob(O.classifier(p, z_t, c.t().contiguous()), y).sum(-1)
    assert float((lq_t - lq).abs().max()) > 1e-3
