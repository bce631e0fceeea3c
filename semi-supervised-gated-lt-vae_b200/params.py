"""Flat parameter store.  All trainable tensors of the model live in ONE fp32 device buffer (and
their gradients / Adam moments in matching buffers) so that data-parallel gradient exchange is a
single all-reduce and the optimiser is a single kernel.  Order = the order
tf.keras lists `model.trainable_variables` for encoder, decoder, classifier, cond_prior, mu
(gated_ccvae.py:30-40,309); layouts are Keras': Conv2D [kh,kw,Cin,Cout], Conv2DTranspose
[kh,kw,Cout,Cin], Dense [in,out]."""
from __future__ import annotations

import math
from collections import OrderedDict

import torch

Z_DIM, Z_CLASSIFY, Y_DIM = 45, 18, 18


def param_specs(z_dim=Z_DIM, z_classify=Z_CLASSIFY, y_dim=Y_DIM, hidden_dim=256):
    return [
        ("enc.conv1.w", (4, 4, 3, 32)), ("enc.conv1.b", (32,)),
        ("enc.conv2.w", (4, 4, 32, 32)), ("enc.conv2.b", (32,)),
        ("enc.conv3.w", (4, 4, 32, 64)), ("enc.conv3.b", (64,)),
        ("enc.conv4.w", (4, 4, 64, 128)), ("enc.conv4.b", (128,)),
        ("enc.conv5.w", (4, 4, 128, hidden_dim)), ("enc.conv5.b", (hidden_dim,)),
        ("enc.locs.w", (hidden_dim, z_dim)), ("enc.locs.b", (z_dim,)),
        ("enc.std.w", (hidden_dim, z_dim)), ("enc.std.b", (z_dim,)),
        ("dec.fc1.w", (z_dim, z_dim)), ("dec.fc1.b", (z_dim,)),
        ("dec.conv1t.w", (4, 4, 128, z_dim)), ("dec.conv1t.b", (128,)),
        ("dec.conv2t.w", (4, 4, 64, 128)), ("dec.conv2t.b", (64,)),
        ("dec.conv3t.w", (4, 4, 32, 64)), ("dec.conv3t.b", (32,)),
        ("dec.conv4t.w", (4, 4, 32, 32)), ("dec.conv4t.b", (32,)),
        ("dec.conv5t.w", (4, 4, 3, 32)), ("dec.conv5t.b", (3,)),
        ("cls.w", (z_classify, y_dim)), ("cls.b", (y_dim,)),
        ("prior.loc_true", (y_dim, z_classify)), ("prior.loc_false", (y_dim, z_classify)),
        ("prior.scale_true", (y_dim, z_classify)), ("prior.scale_false", (y_dim, z_classify)),
        ("mu", (z_classify, y_dim)),
    ]


class ParamStore:
    ALIGN = 4  # floats (16 bytes) so every tensor can be read with 128-bit loads

    def __init__(self, device, specs=None):
        self.specs = specs or param_specs()
        self.offsets = OrderedDict()
        off = 0
        for name, shape in self.specs:
            n = int(math.prod(shape))
            self.offsets[name] = (off, n, tuple(shape))
            off += (n + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.total = off
        self.device = torch.device(device)
        self.flat = torch.zeros(self.total, dtype=torch.float32, device=self.device)
        # gradient buffer + 4 trailing floats: slot [total] carries this rank's share of the loss so that the
        # data-parallel step needs ONE all-reduce for gradients and loss together
        self.grad_full = torch.zeros(self.total + 4, dtype=torch.float32, device=self.device)
        self.grad = self.grad_full[: self.total]
        self.loss_slot = self.grad_full[self.total: self.total + 1]
        # mu sits last so that "all but mu" is one contiguous prefix (frozen-mu modes)
        self.n_without_mu = self.offsets["mu"][0]

    def adopt_grad_buffer(self, buf):
        """use `buf` (>= total + 4 floats, e.g. a symmetric-memory allocation the other ranks of the node can address)
        as the gradient buffer.  Must happen before any kernel pointer into the old buffer is captured."""
        if buf.numel() < self.total + 4 or buf.dtype != torch.float32 or buf.device != self.device:
            raise ValueError("gradient buffer must be {} fp32 elements on {}".format(self.total + 4, self.device))
        buf.zero_()
        self.grad_full = buf
        self.grad = self.grad_full[: self.total]
        self.loss_slot = self.grad_full[self.total: self.total + 1]

    def view(self, name, buf=None):
        off, n, shape = self.offsets[name]
        return (self.flat if buf is None else buf)[off:off + n].view(shape)

    def g(self, name):
        return self.view(name, self.grad)

    def names(self):
        return list(self.offsets.keys())

    def numel(self, with_mu=True):
        return sum(n for k, (_, n, _) in self.offsets.items() if with_mu or k != "mu")

    def load_dict(self, d):
        """Copy a {name: array-like} dict (oracle / Keras layouts) into the store."""
        with torch.no_grad():
            for k, v in d.items():
                if k in self.offsets:
                    self.view(k).copy_(torch.as_tensor(v, dtype=torch.float32).reshape(self.offsets[k][2]))

    def to_dict(self, buf=None):
        return {k: self.view(k, buf).detach().clone() for k in self.offsets}


def keras_default_init(store: ParamStore, seed: int = 0):
    """glorot_uniform / zeros for Conv and Dense, random_normal(0.05) for the classifier
    (networks.py:69-70), zeros / ones for the prior kernels (networks.py:113-116)."""
    gen = torch.Generator().manual_seed(seed)
    out = {}
    for name, (_, _, shape) in store.offsets.items():
        if name == "mu":
            continue
        if name.startswith("cls."):
            t = torch.randn(shape, generator=gen, dtype=torch.float64) * 0.05
        elif name in ("prior.loc_true", "prior.loc_false"):
            t = torch.zeros(shape, dtype=torch.float64)
        elif name in ("prior.scale_true", "prior.scale_false"):
            t = torch.ones(shape, dtype=torch.float64)
        elif name.endswith(".b"):
            t = torch.zeros(shape, dtype=torch.float64)
        else:
            if len(shape) == 2:
                fan_in, fan_out = shape
            else:
                rf = shape[0] * shape[1]
                fan_in, fan_out = shape[-2] * rf, shape[-1] * rf
            limit = math.sqrt(6.0 / (fan_in + fan_out))
            t = (torch.rand(shape, generator=gen, dtype=torch.float64) * 2 - 1) * limit
        out[name] = t.float()
    store.load_dict(out)
