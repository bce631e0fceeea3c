"""CPU oracle for the Gated-CCVAE ELBO training step.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module.  The product package
(``semi-supervised-gated-lt-vae_b200/``) never does.

PARITY UNPINNED.  The reference is TensorFlow 2 / Keras 2.8 / TensorFlow-Probability
(gated_ccvae.py:7,15; networks.py:3-4; utils.py:3-6); none of those can be imported in the
build container and the reference ships no tests, seeds or golden vectors.
This file is therefore a plain PyTorch-CPU restatement that follows the reference line by
line, with every random draw turned into an explicit argument.  It is anchored by
  * the hand-derivable known answers of SURVEY.md §8(c) (tests/test_oracle_known_answers.py),
  * the reference's own gating-matrix files (tests/golden/data, tests/golden/learned),
  * fp64 finite differences of its own loss (tests/test_oracle_gradcheck.py).
The TFP closed forms (Bernoulli / Normal KL / Laplace) are the published definitions of
tensorflow_probability 0.16 (the release contemporaneous with Keras 2.8.0); TFP is not on
disk, so last-ulp op order inside those three functions is unverifiable.

Layouts are the reference's: NHWC activations, Conv2D kernels [kh,kw,Cin,Cout],
Conv2DTranspose kernels [kh,kw,Cout,Cin], Dense kernels [in,out] (Keras conventions,
SURVEY.md quirk 4).  Index convention of the gate: c[i, j], i = z_c index, j = label index
(gated_ccvae.py:54,193).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

Z_DIM = 45          # configs.py:10
Y_DIM = 18          # len(CELEBA_EASY_LABELS), utils_data.py:23-25
Z_CLASSIFY = 18     # gated_ccvae.py:517-518
Z_STYLE = Z_DIM - Z_CLASSIFY
IM_SHAPE = (64, 64, 3)  # gated_ccvae.py:481

# name -> shape, in the order tf.keras lists trainable_variables for
# encoder, decoder, classifier, cond_prior (gated_ccvae.py:30-38) and then mu.
PARAM_SHAPES = [
    # Encoder, networks.py:11-18
    ("enc.conv1.w", (4, 4, 3, 32)), ("enc.conv1.b", (32,)),
    ("enc.conv2.w", (4, 4, 32, 32)), ("enc.conv2.b", (32,)),
    ("enc.conv3.w", (4, 4, 32, 64)), ("enc.conv3.b", (64,)),
    ("enc.conv4.w", (4, 4, 64, 128)), ("enc.conv4.b", (128,)),
    ("enc.conv5.w", (4, 4, 128, 256)), ("enc.conv5.b", (256,)),
    ("enc.locs.w", (256, Z_DIM)), ("enc.locs.b", (Z_DIM,)),
    ("enc.std.w", (256, Z_DIM)), ("enc.std.b", (Z_DIM,)),
    # Decoder(hidden_dim=z_dim), networks.py:43-49 + gated_ccvae.py:34
    ("dec.fc1.w", (Z_DIM, Z_DIM)), ("dec.fc1.b", (Z_DIM,)),
    ("dec.conv1t.w", (4, 4, 128, Z_DIM)), ("dec.conv1t.b", (128,)),
    ("dec.conv2t.w", (4, 4, 64, 128)), ("dec.conv2t.b", (64,)),
    ("dec.conv3t.w", (4, 4, 32, 64)), ("dec.conv3t.b", (32,)),
    ("dec.conv4t.w", (4, 4, 32, 32)), ("dec.conv4t.b", (32,)),
    ("dec.conv5t.w", (4, 4, 3, 32)), ("dec.conv5t.b", (3,)),
    # Classifier, networks.py:69-70
    ("cls.w", (Z_CLASSIFY, Y_DIM)), ("cls.b", (Y_DIM,)),
    # Conditional_Prior, networks.py:113-116  (kernels are [Y, Zc])
    ("prior.loc_true", (Y_DIM, Z_CLASSIFY)), ("prior.loc_false", (Y_DIM, Z_CLASSIFY)),
    ("prior.scale_true", (Y_DIM, Z_CLASSIFY)), ("prior.scale_false", (Y_DIM, Z_CLASSIFY)),
]


def _glorot_uniform(shape, gen, dtype):
    # Keras VarianceScaling(scale=1, mode='fan_avg', distribution='uniform')
    if len(shape) == 2:
        fan_in, fan_out = shape
    else:
        rf = shape[0] * shape[1]
        fan_in, fan_out = shape[-2] * rf, shape[-1] * rf
    limit = math.sqrt(6.0 / (fan_in + fan_out))
    return ((torch.rand(shape, generator=gen, dtype=torch.float64) * 2 - 1) * limit).to(dtype)


def init_params(seed: int = 0, dtype=torch.float32, trained_like: bool = False) -> Dict[str, torch.Tensor]:
    """Keras-default initialisers.  ``trained_like=True`` perturbs the
    zero/one-initialised tensors (biases, prior kernels) so that no gradient route is
    trivially zero in parity tests."""
    gen = torch.Generator().manual_seed(seed)
    p = {}
    for name, shape in PARAM_SHAPES:
        if name.startswith("cls."):
            t = (torch.randn(shape, generator=gen, dtype=torch.float64) * 0.05).to(dtype)
        elif name in ("prior.loc_true", "prior.loc_false"):
            t = torch.zeros(shape, dtype=dtype)
        elif name in ("prior.scale_true", "prior.scale_false"):
            t = torch.ones(shape, dtype=dtype)
        elif name.endswith(".b"):
            t = torch.zeros(shape, dtype=dtype)
        else:
            t = _glorot_uniform(shape, gen, dtype)
        if trained_like:
            if name.endswith(".b") and not name.startswith("cls."):
                t = t + (torch.randn(shape, generator=gen, dtype=torch.float64) * 0.05).to(dtype)
            if name.startswith("prior."):
                t = t + (torch.randn(shape, generator=gen, dtype=torch.float64) * 0.3).to(dtype)
            if name.startswith("cls."):
                t = t * 8.0
        p[name] = t
    return p


def param_count(p) -> int:
    return sum(int(v.numel()) for v in p.values())


# --------------------------------------------------------------------------------------
# bf16 emulation.  The reference computes everything in fp32.  The B200 tensor-core path stores the
# activations between convolutions (and their gradients) in bf16 and feeds bf16 weights to the
# tensor cores.  With EMULATE_BF16 the oracle rounds at exactly those tensor boundaries (round to
# nearest even, straight-through), so that the bf16 kernels can be checked tightly: the Laplace
# likelihood's gradient is sign(x - xhat), a discontinuous function of the forward pass, so without
# identical rounding points ~3e-4 of the pixels flip sign and gradients differ by a few percent.
# --------------------------------------------------------------------------------------
EMULATE_BF16 = False


class bf16_emulation:
    def __init__(self, on=True):
        self.on = on

    def __enter__(self):
        global EMULATE_BF16
        self.prev, EMULATE_BF16 = EMULATE_BF16, self.on

    def __exit__(self, *a):
        global EMULATE_BF16
        EMULATE_BF16 = self.prev


def _r16(t):
    return t.to(torch.bfloat16).to(t.dtype)


class _RoundBoth(torch.autograd.Function):        # activation stored in bf16, and so is its gradient
    @staticmethod
    def forward(ctx, t):
        return _r16(t)

    @staticmethod
    def backward(ctx, g):
        return _r16(g)


class _RoundBwd(torch.autograd.Function):         # fp32 forward value, gradient handed back in bf16
    @staticmethod
    def forward(ctx, t):
        return t.clone()

    @staticmethod
    def backward(ctx, g):
        return _r16(g)


class _RoundFwd(torch.autograd.Function):         # bf16 operand copy of an fp32 master weight
    @staticmethod
    def forward(ctx, t):
        return _r16(t)

    @staticmethod
    def backward(ctx, g):
        return g


def _act16(t):
    return _RoundBoth.apply(t) if EMULATE_BF16 else t


def _grad16(t):
    return _RoundBwd.apply(t) if EMULATE_BF16 else t


def _w16(t):
    return _RoundFwd.apply(t) if EMULATE_BF16 else t


# --------------------------------------------------------------------------------------
# Networks (networks.py)
# --------------------------------------------------------------------------------------
def _conv(h_nhwc, w, b, stride, pad):
    """Keras Conv2D('valid') after an explicit tf.pad of ``pad`` (networks.py:21-29)."""
    y = F.conv2d(h_nhwc.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), b, stride=stride, padding=pad)
    return y.permute(0, 2, 3, 1)


def _convT(h_nhwc, w, b, stride, pad):
    """Keras Conv2DTranspose with kernel [kh,kw,Cout,Cin] (networks.py:45-49).
    'same' k4 s2 == torch padding 1; 'valid' k4 s1 == padding 0."""
    y = F.conv_transpose2d(h_nhwc.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), b, stride=stride, padding=pad)
    return y.permute(0, 2, 3, 1)


def clipped_softplus(t):
    return torch.clamp(F.softplus(t, beta=1.0, threshold=1e30), 1e-3, 1e3)


def encoder(p, x, return_acts: bool = False):
    """networks.py:20-37.  x [B,64,64,3] -> (locs [B,45], scale [B,45])."""
    if EMULATE_BF16:
        x = _r16(x)
    h1 = _act16(F.relu(_conv(x, _w16(p["enc.conv1.w"]), p["enc.conv1.b"], 2, 1)))
    h2 = _act16(F.relu(_conv(h1, _w16(p["enc.conv2.w"]), p["enc.conv2.b"], 2, 1)))
    h3 = _act16(F.relu(_conv(h2, _w16(p["enc.conv3.w"]), p["enc.conv3.b"], 2, 1)))
    h4 = _act16(F.relu(_conv(h3, _w16(p["enc.conv4.w"]), p["enc.conv4.b"], 2, 1)))
    h5 = _act16(F.relu(_conv(h4, _w16(p["enc.conv5.w"]), p["enc.conv5.b"], 1, 0)))
    hf = h5.reshape(h5.shape[0], -1)
    locs = F.relu(_grad16(hf @ _w16(p["enc.locs.w"]) + p["enc.locs.b"]))
    scale = clipped_softplus(_grad16(hf @ _w16(p["enc.std.w"]) + p["enc.std.b"]))
    if return_acts:
        return locs, scale, (h1, h2, h3, h4, h5)
    return locs, scale


def decoder(p, z, return_acts: bool = False):
    """networks.py:51-59 with hidden_dim = z_dim (gated_ccvae.py:34)."""
    g0 = _act16(F.relu(_w16(z) @ _w16(p["dec.fc1.w"]) + p["dec.fc1.b"]))
    g0r = g0.reshape(g0.shape[0], 1, 1, g0.shape[1])
    g1 = _act16(F.relu(_convT(g0r, _w16(p["dec.conv1t.w"]), p["dec.conv1t.b"], 1, 0)))
    g2 = _act16(F.relu(_convT(g1, _w16(p["dec.conv2t.w"]), p["dec.conv2t.b"], 2, 1)))
    g3 = _act16(F.relu(_convT(g2, _w16(p["dec.conv3t.w"]), p["dec.conv3t.b"], 2, 1)))
    g4 = _act16(F.relu(_convT(g3, _w16(p["dec.conv4t.w"]), p["dec.conv4t.b"], 2, 1)))
    xh = torch.sigmoid(_grad16(_convT(g4, _w16(p["dec.conv5t.w"]), p["dec.conv5t.b"], 2, 1)))
    if return_acts:
        return xh, (g0, g1, g2, g3, g4)
    return xh


def classifier(p, encodes_z, gates):
    """networks.py:72-74, 83-86: reduce_sum((z_tiled * gates) * kernel, axis=1) + bias."""
    gated_z = encodes_z * gates
    return torch.sum(gated_z * p["cls.w"], dim=1) + p["cls.b"]


def cond_prior(p, y, c):
    """networks.py:104-106, 118-127.  y tiled [B,Y,Zc]; c [Zc,Y] (transposed inside)."""
    ct = c.t()
    y = y.to(ct.dtype)  # Keras autocast of the float64 input back to the layer dtype
    locs = torch.sum((y * ct) * p["prior.loc_true"], dim=1) + torch.sum(((1 - y) * ct) * p["prior.loc_false"], dim=1)
    scale = torch.sum((y * ct) * p["prior.scale_true"], dim=1) + torch.sum(((1 - y) * ct) * p["prior.scale_false"], dim=1)
    return locs, clipped_softplus(scale)


# --------------------------------------------------------------------------------------
# Samplers (gated_ccvae.py:62-64, 90-93, 102-111)
# --------------------------------------------------------------------------------------
def sample_gumbel(U, eps=1e-20):
    return -torch.log(-torch.log(U + eps) + eps)


def sample_normal(mu, std, epsilon):
    return mu + std * epsilon


def sample_gating_parameter(mu, temperature, U1, U2, EPSILON=1e-20):
    mu = torch.clamp(mu, 0.0, 1.0)
    eps1 = sample_gumbel(U1)
    eps2 = sample_gumbel(U2)
    num = torch.exp((eps2 - eps1) / temperature)
    t1 = torch.pow(mu, 1.0 / temperature)
    t2 = torch.pow(1.0 - mu, 1.0 / temperature) * num
    return t1 / (t1 + t2 + EPSILON)


def initialise_mu(train_config, dtype=torch.float32):
    """gated_ccvae.py:42-60 -> (mu, trainable)."""
    gt, gs = train_config["gate_type"], train_config.get("gate_subtype")
    if gt == "learnable":
        return torch.as_tensor(np.asarray(train_config["mu_init"]), dtype=dtype).clone(), True
    if gt == "fixed" and gs == "inferred":
        return torch.as_tensor(np.asarray(train_config["mu_init"]), dtype=dtype).clone(), False
    if gt == "fixed" and gs == "one-one":
        return torch.eye(Z_CLASSIFY, Y_DIM, dtype=dtype), False
    raise ValueError("Invalid gate type/subtype: {}/{}".format(gt, gs))


# --------------------------------------------------------------------------------------
# TFP distributions restated (call sites: SURVEY.md A9-A11)
# --------------------------------------------------------------------------------------
def _multiply_no_nan(x, y):
    return torch.where(y == 0, torch.zeros_like(x), x * y)


def bernoulli_logits_log_prob(logits, event):
    event = event.to(logits.dtype)
    lp0 = -F.softplus(logits, threshold=1e30)
    lp1 = -F.softplus(-logits, threshold=1e30)
    return _multiply_no_nan(lp0, 1 - event) + _multiply_no_nan(lp1, event)


def bernoulli_probs_log_prob(probs, event):
    event = event.to(probs.dtype)
    return _multiply_no_nan(torch.log1p(-probs), 1 - event) + _multiply_no_nan(torch.log(probs), event)


def bernoulli_sample(logits, U):
    """TFP Bernoulli._sample_n: uniform < probs, cast to int32."""
    return (U < torch.sigmoid(logits)).to(torch.int32)


def get_gaussian_kl_div(locs_q, scale_q, locs_p=None, scale_p=None):
    """utils.py:108-119 with TFP _kl_normal_normal."""
    if locs_p is None:
        locs_p = torch.zeros_like(locs_q)
    if scale_p is None:
        scale_p = torch.ones_like(scale_q)
    diff_log_scale = torch.log(scale_q) - torch.log(scale_p)
    kl = 0.5 * (locs_q / scale_p - locs_p / scale_p) ** 2 + 0.5 * torch.expm1(2.0 * diff_log_scale) - diff_log_scale
    return kl.sum(-1)


def img_log_likelihood(recon, xs):
    """utils.py:101-105: sum_{h,w,c} Laplace(recon, 1).log_prob(xs)."""
    lp = -torch.abs(xs - recon) - math.log(2.0)
    return lp.sum(dim=(1, 2, 3))


# --------------------------------------------------------------------------------------
# Losses (gated_ccvae.py:167-300)
# --------------------------------------------------------------------------------------
def classifier_loss(p, x, y, c, eps_k):
    """gated_ccvae.py:167-182.  eps_k [K,B,45]: the k-th full-width N(0,1) draw."""
    post_locs, post_scales = encoder(p, x)          # second encoder pass, :168
    k = eps_k.shape[0]
    rows = []
    for i in range(k):
        z = sample_normal(post_locs, post_scales, eps_k[i])
        z_classify = z[:, Z_STYLE:]
        z_tiled = z_classify.unsqueeze(-1).repeat(1, 1, Y_DIM)
        logits = classifier(p, z_tiled, c)
        rows.append(bernoulli_logits_log_prob(logits, y).sum(-1).unsqueeze(0))
    stack = torch.cat(rows, dim=0)
    return torch.logsumexp(stack, dim=0) - math.log(float(k))


def _prior_and_kl(p, y, c, post_locs, post_scales):
    b = post_locs.shape[0]
    y_tiled = y.unsqueeze(-1).repeat(1, 1, Z_CLASSIFY).to(torch.float64)
    prior_locs, prior_scales = cond_prior(p, y_tiled, c)
    pl = torch.cat([torch.zeros(b, Z_STYLE, dtype=post_locs.dtype), prior_locs], dim=-1)
    ps = torch.cat([torch.ones(b, Z_STYLE, dtype=post_locs.dtype), prior_scales], dim=-1)
    return get_gaussian_kl_div(post_locs, post_scales, pl, ps), prior_locs, prior_scales


def _l1(mu, train_config):
    return train_config["gating_reg"] * torch.mean(torch.abs(mu))


def sup_loss(p, mu, x, y, noise, train_config, temperature, zc_detached_override=None):
    """gated_ccvae.py:234-300.  noise: eps [B,45], U1,U2 [18,18], eps_k [K,B,45]."""
    post_locs, post_scales = encoder(p, x)
    z = sample_normal(post_locs, post_scales, noise["eps"])
    z_classify = z[:, Z_STYLE:]
    c = sample_gating_parameter(mu, temperature, noise["U1"], noise["U2"])
    z_tiled = z_classify.unsqueeze(-1).repeat(1, 1, Y_DIM)
    logits = classifier(p, z_tiled, c)
    log_qy_zc = bernoulli_logits_log_prob(logits, y).sum(-1)
    log_qy_x = classifier_loss(p, x, y, c, noise["eps_k"])
    p_Y = torch.full((x.shape[0], Y_DIM), 0.5, dtype=post_locs.dtype)   # :141, :259
    log_py = bernoulli_probs_log_prob(p_Y, y).sum(-1)
    kl, prior_locs, prior_scales = _prior_and_kl(p, y, c, post_locs, post_scales)
    recon_x = decoder(p, z)
    log_pxz = img_log_likelihood(recon_x, x)
    zc_det = z_classify.detach() if zc_detached_override is None else zc_detached_override
    logits_ = classifier(p, zc_det.unsqueeze(-1).repeat(1, 1, Y_DIM), c)
    log_qy_zc_ = bernoulli_logits_log_prob(logits_, y).sum(-1)
    w = torch.exp(log_qy_zc_ - log_qy_x)
    elbo = w * (log_pxz - kl - log_qy_zc) + log_py + log_qy_x
    loss = torch.mean(-elbo)
    if train_config["gate_type"] == "learnable":
        loss = loss + _l1(mu, train_config)
    return dict(loss=loss, c=c, post_locs=post_locs, post_scales=post_scales, z=z, logits=logits,
                log_qy_zc=log_qy_zc, log_qy_x=log_qy_x, log_py=log_py, kl=kl, prior_locs=prior_locs,
                prior_scales=prior_scales, recon=recon_x, log_pxz=log_pxz, w=w, elbo=elbo)


def unsup_loss(p, mu, x, noise, train_config, temperature):
    """gated_ccvae.py:184-232.  noise: eps [B,45], U1,U2 [18,18], U_y [B,18]."""
    post_locs, post_scales = encoder(p, x)
    z = sample_normal(post_locs, post_scales, noise["eps"])
    z_classify = z[:, Z_STYLE:]
    c = sample_gating_parameter(mu, temperature, noise["U1"], noise["U2"])
    z_tiled = z_classify.unsqueeze(-1).repeat(1, 1, Y_DIM)
    logits = classifier(p, z_tiled, c)
    y = bernoulli_sample(logits, noise["U_y"])
    log_qy_zc = bernoulli_logits_log_prob(logits, y).sum(-1)
    p_Y = torch.full((x.shape[0], Y_DIM), 0.5, dtype=post_locs.dtype)
    log_py = bernoulli_probs_log_prob(p_Y, y).sum(-1)
    kl, prior_locs, prior_scales = _prior_and_kl(p, y, c, post_locs, post_scales)
    recon_x = decoder(p, z)
    log_pxz = img_log_likelihood(recon_x, x)
    elbo = log_pxz + log_py - kl - log_qy_zc
    loss = torch.mean(-elbo)
    if train_config["gate_type"] == "learnable":
        loss = loss + _l1(mu, train_config)
    return dict(loss=loss, c=c, post_locs=post_locs, post_scales=post_scales, z=z, logits=logits, y=y,
                log_qy_zc=log_qy_zc, log_py=log_py, kl=kl, prior_locs=prior_locs, prior_scales=prior_scales,
                recon=recon_x, log_pxz=log_pxz, elbo=elbo)


def classifier_accuracy(p, mu, x, y, noise, temperature):
    """gated_ccvae.py:421-446."""
    post_locs, post_scales = encoder(p, x)
    z = sample_normal(post_locs, post_scales, noise["eps"])
    c = sample_gating_parameter(mu, temperature, noise["U1"], noise["U2"])
    z_tiled = z[:, Z_STYLE:].unsqueeze(-1).repeat(1, 1, Y_DIM)
    y_hat = torch.round(torch.sigmoid(classifier(p, z_tiled, c)))
    return (y_hat == y.to(y_hat.dtype)).to(torch.float32).mean()


# --------------------------------------------------------------------------------------
# train_step (gated_ccvae.py:302-311): tape.gradient + Keras-2.8 Adam
# --------------------------------------------------------------------------------------
def loss_and_grads(p, mu, x, y, noise, train_config, temperature, supervised: bool):
    """Returns (terms, grads) with grads keyed like ``p`` plus 'mu' (None when mu is frozen)."""
    leaf = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}
    learnable = train_config["gate_type"] == "learnable"
    mu_leaf = mu.detach().clone().requires_grad_(learnable)
    if supervised:
        out = sup_loss(leaf, mu_leaf, x, y, noise, train_config, temperature)
    else:
        out = unsup_loss(leaf, mu_leaf, x, noise, train_config, temperature)
    out["loss"].backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaf.items()}
    grads["mu"] = mu_leaf.grad if learnable else None
    return {k: (v.detach() if torch.is_tensor(v) else v) for k, v in out.items()}, grads


class KerasAdam:
    """tf.keras.optimizers.Adam(lr) of Keras 2.8 (gated_ccvae.py:144): beta1=.9, beta2=.999,
    epsilon=1e-7, update  theta -= lr*sqrt(1-b2^t)/(1-b1^t) * m / (sqrt(v) + eps)."""

    def __init__(self, lr, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        self.lr, self.b1, self.b2, self.eps, self.t = lr, beta_1, beta_2, epsilon, 0
        self.m, self.v = {}, {}

    def apply(self, params: Dict[str, torch.Tensor], grads: Dict[str, Optional[torch.Tensor]]):
        self.t += 1
        lr_t = self.lr * math.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t)
        for k, g in grads.items():
            if g is None:
                continue
            if k not in self.m:
                self.m[k] = torch.zeros_like(params[k])
                self.v[k] = torch.zeros_like(params[k])
            self.m[k] += (g - self.m[k]) * (1 - self.b1)
            self.v[k] += (g * g - self.v[k]) * (1 - self.b2)
            params[k] -= lr_t * self.m[k] / (torch.sqrt(self.v[k]) + self.eps)


# --------------------------------------------------------------------------------------
# Synthetic inputs
# --------------------------------------------------------------------------------------
def make_inputs(batch, k=100, seed=1234, dtype=torch.float32, gate_seed=1234):
    g = torch.Generator().manual_seed(seed)
    gg = torch.Generator().manual_seed(gate_seed)
    x = torch.rand(batch, *IM_SHAPE, generator=g, dtype=torch.float64).to(dtype)
    y = (torch.rand(batch, Y_DIM, generator=g) < 0.5).to(torch.int64)
    noise = dict(
        eps=torch.randn(batch, Z_DIM, generator=g, dtype=torch.float64).to(dtype),
        eps_k=torch.randn(k, batch, Z_DIM, generator=g, dtype=torch.float64).to(dtype),
        U_y=torch.rand(batch, Y_DIM, generator=g, dtype=torch.float64).to(dtype),
        U1=torch.rand(Z_CLASSIFY, Y_DIM, generator=gg, dtype=torch.float64).to(dtype),
        U2=torch.rand(Z_CLASSIFY, Y_DIM, generator=gg, dtype=torch.float64).to(dtype),
    )
    return x, y, noise
