"""Host-side mirror of the reference's network classes (networks.py:7-127) over the CUDA kernels.
Same constructor arguments, call signatures, shapes (NHWC) and return conventions; parameters are
views into a shared flat ParamStore in Keras layouts.  Module-level calls are forward-only; training
goes through Learner (gated_ccvae.py), which runs the fused forward+backward step."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .engine import DEC_LAYERS, Engine, _stream
from .params import ParamStore, keras_default_init
from ._lib import ptr


def _default_device():
    if not torch.cuda.is_available():
        raise _lib.GccvaeError("no CUDA device: the Gated-CCVAE kernels are sm_100a-only and have no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def as_device_f32(a, device):
    """numpy / torch (any device, any float dtype) -> contiguous fp32 device tensor."""
    t = torch.as_tensor(a) if not torch.is_tensor(a) else a
    t = t.to(device=device, dtype=torch.float32, non_blocking=True).contiguous()
    if t.data_ptr() % 16 != 0:   # views with a storage offset (e.g. eps_k[..., 27:]) may be misaligned
        t = t.clone()
    return t


class _Net:
    def __init__(self, store=None, engine=None, device=None):
        if store is None:
            store = ParamStore(device or _default_device())
            keras_default_init(store)
        self.store = store
        self.engine = engine or Engine(store)
        self.lib = self.engine.lib
        self.device = store.device

    def _weights(self, prefix):
        return [self.store.view(k) for k in self.store.names() if k.startswith(prefix)]


class Encoder(_Net):
    """networks.py:7-37.  encoder(x[B,64,64,3]) -> (locs[B,z], scale[B,z])."""

    def __init__(self, z_dim, hidden_dim=256, **kw):
        if z_dim != 45 or hidden_dim != 256:
            raise ValueError("the sm_100a kernels are specialised for z_dim=45, hidden_dim=256 (configs.py:10)")
        super().__init__(**kw)
        self.z_dim = z_dim

    @property
    def trainable_variables(self):
        return self._weights("enc.")

    def __call__(self, x):
        x = as_device_f32(x, self.device)
        B = x.shape[0]
        b = self.engine.bufs(B)
        loc_pre, scale_pre = self.engine.encoder_fwd(x, b)
        loc_pre, scale_pre = loc_pre.contiguous(), scale_pre.contiguous()
        loc, scale = torch.empty_like(loc_pre), torch.empty_like(scale_pre)
        _lib.check(self.lib.gccvae_head_act_f32(ptr(loc_pre), ptr(scale_pre), loc.numel(), ptr(loc), ptr(scale),
                                                _stream()), "head_act")
        return loc, scale

    call = __call__


class Decoder(_Net):
    """networks.py:40-59 with hidden_dim = z_dim (gated_ccvae.py:34).  decoder(z[B,45]) -> x[B,64,64,3]."""

    def __init__(self, hidden_dim=256, *args, **kw):
        if hidden_dim != 45:
            raise ValueError("the sm_100a kernels are specialised for Decoder(hidden_dim=z_dim=45) (gated_ccvae.py:34)")
        super().__init__(**kw)

    @property
    def trainable_variables(self):
        return self._weights("dec.")

    def __call__(self, z):
        z = as_device_f32(z, self.device)
        b = self.engine.bufs(z.shape[0])
        return self.engine.decoder_fwd(z, b).clone()

    call = __call__


class Classifier(_Net):
    """networks.py:62-86.  logits[b,j] = sum_i z_t[b,i,j] * gates[i,j] * W[i,j] + bias[j].
    Accepts the reference's TILED input [B,Zc,Y] (anything broadcastable to it) or [B,Zc]."""

    def __init__(self, y_dim, **kw):
        if y_dim != 18:
            raise ValueError("kernels are specialised for y_dim = 18 (utils_data.py:23-25)")
        super().__init__(**kw)
        self.y_dim = y_dim

    @property
    def trainable_variables(self):
        return self._weights("cls.")

    def __call__(self, encodes_z, gates):
        z = as_device_f32(encodes_z, self.device)
        gates = as_device_f32(gates, self.device)
        if z.dim() == 2:
            z = z.unsqueeze(-1)
        z = z.expand(z.shape[0], 18, 18)
        B = z.shape[0]
        out = torch.empty(B, 18, dtype=torch.float32, device=self.device)
        sb, si, sj = z.stride()
        _lib.check(self.lib.gccvae_classifier_tiled_f32(ptr(z), sb, si, sj, B, ptr(gates),
                                                        ptr(self.store.view("cls.w")), ptr(self.store.view("cls.b")),
                                                        ptr(out), _stream()), "classifier")
        return out

    call = __call__


class Conditional_Prior(_Net):
    """networks.py:89-127.  cond_prior(y_tiled[B,Y,Zc], c[Zc,Y]) -> (locs[B,Zc], scale[B,Zc])."""

    def __init__(self, z_dim, **kw):
        if z_dim != 18:
            raise ValueError("kernels are specialised for z_classify = 18 (gated_ccvae.py:517)")
        super().__init__(**kw)
        self.z_dim = z_dim

    @property
    def trainable_variables(self):
        return self._weights("prior.")

    def __call__(self, y, c):
        y = as_device_f32(y, self.device)
        c = as_device_f32(c, self.device)
        if y.dim() == 2:
            y = y.unsqueeze(-1)
        y = y.expand(y.shape[0], 18, 18)
        B = y.shape[0]
        loc = torch.empty(B, 18, dtype=torch.float32, device=self.device)
        scale = torch.empty_like(loc)
        sb, sj, si = y.stride()
        v = self.store.view
        _lib.check(self.lib.gccvae_cond_prior_tiled_f32(ptr(y), sb, sj, si, B, ptr(c), ptr(v("prior.loc_true")),
                                                        ptr(v("prior.loc_false")), ptr(v("prior.scale_true")),
                                                        ptr(v("prior.scale_false")), ptr(loc), ptr(scale), _stream()),
                   "cond_prior")
        return loc, scale

    call = __call__
