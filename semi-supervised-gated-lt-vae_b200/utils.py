"""utils.py functions that are on the ELBO path (imported at gated_ccvae.py:10), CUDA-backed."""
from __future__ import annotations

import torch

from . import _lib
from ._lib import ptr
from .engine import _stream
from .networks import _default_device, as_device_f32


def img_log_likelihood(recon, xs):
    """utils.py:101-105: sum_{h,w,c} Laplace(recon, 1).log_prob(xs) -> [B]."""
    lib = _lib.load()
    dev = recon.device if torch.is_tensor(recon) and recon.is_cuda else _default_device()
    recon, xs = as_device_f32(recon, dev), as_device_f32(xs, dev)
    if recon.shape != xs.shape:
        raise ValueError("img_log_likelihood: shapes differ: {} vs {}".format(tuple(recon.shape), tuple(xs.shape)))
    B = xs.shape[0]
    out = torch.empty(B, dtype=torch.float32, device=dev)
    _lib.check(lib.gccvae_recon_f32(ptr(xs), ptr(recon), B, xs[0].numel(), None, ptr(out), None, _stream()), "recon")
    return out


def get_gaussian_kl_div(locs_q, scale_q, locs_p=None, scale_p=None):
    """utils.py:108-119: sum_d KL(N(locs_q, scale_q) || N(locs_p, scale_p)) -> [B]."""
    lib = _lib.load()
    dev = locs_q.device if torch.is_tensor(locs_q) and locs_q.is_cuda else _default_device()
    lq, sq = as_device_f32(locs_q, dev), as_device_f32(scale_q, dev)
    lp = None if locs_p is None else as_device_f32(locs_p, dev)
    sp = None if scale_p is None else as_device_f32(scale_p, dev)
    B, D = lq.shape
    out = torch.empty(B, dtype=torch.float32, device=dev)
    _lib.check(lib.gccvae_gaussian_kl_f32(ptr(lq), ptr(sq), ptr(lp), ptr(sp), B, D, ptr(out), _stream()), "kl")
    return out
from .utils_data import create_gating_matrix  # noqa: F401,E402  (utils.py:132-149 lives next to the data module here)
