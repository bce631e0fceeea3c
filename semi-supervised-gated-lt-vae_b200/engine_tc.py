"""bf16 tensor-core executor of the ELBO step (tcgen05 / TMEM / TMA kernels of csrc/conv_tc.cu).

Activations between the convolutions are bf16 NHWC; accumulation, biases, the latent stage, the
losses and all gradients of parameters are fp32.  Layer -> kernel map:

  enc.conv1   x fp32 -> K=64 im2col matrix X64 (bf16) -> dense tcgen05 GEMM; wgrad from X64
  enc.conv2-4 tap-GEMM (16 taps, stride-2 TMA boxes); dgrad = 4-phase tap-GEMM; wgrad = MN-major GEMM
  enc.conv5   dense tcgen05 GEMM [B,2048]x[2048,256] (fp32 output for the heads)
  heads, dec.fc1, dec.conv1t   (45-wide, <1% of the FLOPs) fp32 CUDA-core kernels of conv_f32.cu
  dec.conv2t-4t  4-phase tap-GEMM; dgrad = 16-tap tap-GEMM; wgrad = MN-major GEMM
  dec.conv5t  4-phase tap-GEMM, N=16 (3 real channels), float4 image output; the reconstruction
              log-likelihood kernel emits its gradient directly as the im2col matrix G64, from which
              dgrad and wgrad are dense GEMMs
Packed bf16 weight operands are refreshed from the fp32 master parameters at the start of every step.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import ACT_ACCUMULATE, ACT_NONE, ACT_RELU, ACT_SIGMOID, Geom, ptr
from .engine import DEC_LAYERS, ENC_LAYERS, HEAD_LAYERS, Engine, _stream, make_geom, out_shape

BF16 = torch.bfloat16
_ENC = {lay[0]: lay for lay in ENC_LAYERS}
_DEC = {lay[0]: lay for lay in DEC_LAYERS}
TC_ENC = ["enc.conv2", "enc.conv3", "enc.conv4", "enc.conv5"]
TC_DEC = ["dec.conv2t", "dec.conv3t", "dec.conv4t"]


def _dense64_geom(rows, cs):
    """[rows, 64] x [64, cs] as a 1x1 'convolution' (used for the im2col'd end layers)."""
    return Geom(rows, 1, 1, 64, 1, 1, cs, 1, 1, 1, 0)


class EngineTC(Engine):
    precision = "bf16"

    def __init__(self, store):
        super().__init__(store)
        dev = self.device
        lib = self.lib
        self.wp = {}
        for name in TC_ENC + TC_DEC + ["dec.conv5t"]:
            lay = _ENC.get(name) or _DEC[name]
            g = make_geom(lay, 1)
            self.wp[name + ".ls"] = torch.zeros(lib.gccvae_packed_weight_elems(C.byref(g), 0), dtype=BF16, device=dev)
            self.wp[name + ".sl"] = torch.zeros(lib.gccvae_packed_weight_elems(C.byref(g), 1), dtype=BF16, device=dev)
        self.wp["enc.conv1.c4"] = torch.zeros(32 * 64, dtype=BF16, device=dev)
        self.wp["dec.conv5t.c4"] = torch.zeros(32 * 64, dtype=BF16, device=dev)
        self._jobs = None

    # ---- packed weights ---------------------------------------------------------------------------------
    def pack_weights(self):
        """refresh every packed bf16 operand from the fp32 master weights: one kernel launch."""
        if self._jobs is None:
            v = self.store.view
            jobs = []
            for name in TC_ENC + TC_DEC:
                lay = _ENC.get(name) or _DEC[name]
                _, _, (HL, WL, CL), (HS, WS, CS), k, s_, p_, _ = lay
                jobs.append((0, k * k, CL, CS, v(name + ".w"), self.wp[name + ".ls"]))
                kind = 2 if (HS == 1 and WS == 1) else 1
                jobs.append((kind, k * k, CL, CS, v(name + ".w"), self.wp[name + ".sl"]))
            jobs.append((1, 16, 3, 32, v("dec.conv5t.w"), self.wp["dec.conv5t.sl"]))
            jobs.append((3, 16, 3, 32, v("enc.conv1.w"), self.wp["enc.conv1.c4"]))
            jobs.append((3, 16, 3, 32, v("dec.conv5t.w"), self.wp["dec.conv5t.c4"]))
            arr = (_lib.PackJob * len(jobs))()
            for i, (kind, taps, CL, CS, W, out) in enumerate(jobs):
                arr[i] = _lib.PackJob(kind, taps, CL, CS, ptr(W), ptr(out))
            self._jobs = arr
        _lib.check(self.lib.gccvae_pack_jobs_bf16(self._jobs, len(self._jobs), _stream()), "pack_jobs")

    # ---- buffers -----------------------------------------------------------------------------------------
    def _alloc(self, B):
        dev = self.device
        e = lambda *s, dt=torch.float32: torch.empty(*s, dtype=dt, device=dev)
        b = {}
        b["X64"] = e(B * 1024, 64, dt=BF16)
        b["G64"] = e(B * 1024, 64, dt=BF16)
        for name in ["enc.conv1", "enc.conv2", "enc.conv3", "enc.conv4"]:
            oh, ow, oc = out_shape(_ENC[name])
            b[name + ".out"] = e(B, oh, ow, oc, dt=BF16)
            b[name + ".dout"] = e(B, oh, ow, oc, dt=BF16)
        b["enc.conv5.out"] = e(B, 1, 1, 256)            # fp32: consumed by the fp32 heads
        b["enc.conv5.dout"] = e(B, 1, 1, 256)
        b["enc.conv5.dout16"] = e(B, 1, 1, 256, dt=BF16)
        for name in ["enc.locs", "enc.std"]:
            b[name + ".out"] = e(B, 1, 1, 45)
            b[name + ".dout"] = e(B, 1, 1, 45)
        b["dec.fc1.out"] = e(B, 1, 1, 45)
        b["dec.fc1.dout"] = e(B, 1, 1, 45)
        b["dec.conv1t.out"] = e(B, 4, 4, 128)           # fp32 (CUDA-core layer) + bf16 copy for conv2t
        b["dec.conv1t.out16"] = e(B, 4, 4, 128, dt=BF16)
        b["dec.conv1t.dout"] = e(B, 4, 4, 128)
        b["dec.conv1t.dout16"] = e(B, 4, 4, 128, dt=BF16)
        for name in TC_DEC:
            oh, ow, oc = out_shape(_DEC[name])
            b[name + ".out"] = e(B, oh, ow, oc, dt=BF16)
            b[name + ".dout"] = e(B, oh, ow, oc, dt=BF16)
        b["xhat4"] = e(B, 64, 64, 4)
        b["dz"] = e(B, 45)
        ws_bytes = 0
        for lay in HEAD_LAYERS + DEC_LAYERS[:2]:
            g = make_geom(lay, B)
            ws_bytes = max(ws_bytes, self.lib.gccvae_wg_f32_workspace_bytes(C.byref(g)))
            oh, ow, oc = out_shape(lay)
            ws_bytes = max(ws_bytes, self.lib.gccvae_colsum_f32_workspace_bytes(B * oh * ow, oc))
        b["ws"] = torch.empty(ws_bytes // 4 + 16, dtype=torch.float32, device=dev)
        b["ws_bytes"] = ws_bytes
        return b

    # ---- helpers --------------------------------------------------------------------------------------------
    def _cast16(self, src, dst):
        _lib.check(self.lib.gccvae_cast_f32_to_bf16(ptr(src), src.numel(), ptr(dst), _stream()), "cast->bf16")

    def _cast32(self, src, dst):
        _lib.check(self.lib.gccvae_cast_bf16_to_f32(ptr(src), src.numel(), ptr(dst), _stream()), "cast->f32")

    def _bias_grad16(self, dout, name):
        rows = dout.numel() // dout.shape[-1]
        _lib.check(self.lib.gccvae_colsum_bf16(ptr(dout), rows, dout.shape[-1], ptr(self.store.g(name + ".b")),
                                               _stream()), name + " bgrad")

    def zero_grads(self):
        self.store.grad.zero_()   # the tensor-core wgrad / bias-grad kernels accumulate (split-K red.global)

    # ---- forward -----------------------------------------------------------------------------------------------
    def encoder_fwd(self, x, b):
        B = x.shape[0]
        lib, st, v = self.lib, _stream(), self.store.view
        self.pack_weights()
        _lib.check(lib.gccvae_im2col_x_bf16(ptr(x), B, ptr(b["X64"]), st), "im2col_x")
        g = _dense64_geom(B * 1024, 32)
        _lib.check(lib.gccvae_ls_bf16(C.byref(g), ptr(b["X64"]), ptr(self.wp["enc.conv1.c4"]), ptr(v("enc.conv1.b")),
                                      ACT_RELU, None, ptr(b["enc.conv1.out"]), 0, st), "conv1 fwd")
        h = b["enc.conv1.out"]
        for name in TC_ENC:
            lay = _ENC[name]
            g = make_geom(lay, B)
            f32 = 1 if name == "enc.conv5" else 0
            _lib.check(lib.gccvae_ls_bf16(C.byref(g), ptr(h), ptr(self.wp[name + ".ls"]), ptr(v(name + ".b")), ACT_RELU,
                                          None, ptr(b[name + ".out"]), f32, st), name + " fwd")
            h = b[name + ".out"]
        for lay in HEAD_LAYERS:
            self.layer_fwd(lay, B, h, b[lay[0] + ".out"])
        return b["enc.locs.out"].view(B, 45), b["enc.std.out"].view(B, 45)

    def decoder_fwd(self, z, b):
        B = z.shape[0]
        lib, st, v = self.lib, _stream(), self.store.view
        self.layer_fwd(DEC_LAYERS[0], B, z, b["dec.fc1.out"])
        self.layer_fwd(DEC_LAYERS[1], B, b["dec.fc1.out"], b["dec.conv1t.out"])
        self._cast16(b["dec.conv1t.out"], b["dec.conv1t.out16"])
        h = b["dec.conv1t.out16"]
        for name in TC_DEC:
            g = make_geom(_DEC[name], B)
            _lib.check(lib.gccvae_sl_bf16(C.byref(g), ptr(h), ptr(self.wp[name + ".sl"]), ptr(v(name + ".b")), ACT_RELU,
                                          None, ptr(b[name + ".out"]), 0, st), name + " fwd")
            h = b[name + ".out"]
        g = make_geom(_DEC["dec.conv5t"], B)
        _lib.check(lib.gccvae_sl_bf16(C.byref(g), ptr(h), ptr(self.wp["dec.conv5t.sl"]), ptr(v("dec.conv5t.b")),
                                      ACT_SIGMOID, None, ptr(b["xhat4"]), 2, st), "conv5t fwd")
        return b["xhat4"]

    def recon(self, x, b, coef, log_pxz, backward):
        B = x.shape[0]
        _lib.check(self.lib.gccvae_recon_im2col_bf16(
            ptr(x), ptr(b["xhat4"]), B, ptr(coef) if backward else None, ptr(log_pxz),
            ptr(b["G64"]) if backward else None, ptr(self.store.g("dec.conv5t.b")) if backward else None, _stream()),
            "recon_im2col")
        return b["xhat4"][..., :3]

    # ---- backward -----------------------------------------------------------------------------------------------
    def decoder_bwd(self, z, b, want_dz=True):
        B = z.shape[0]
        lib, st, g_ = self.lib, _stream(), self.store.g
        ws, wsb = b["ws"], b["ws_bytes"]
        # conv5t from the im2col'd logit gradient
        g4 = b["dec.conv4t.out"]
        _lib.check(lib.gccvae_wg_c4_bf16(B * 1024, ptr(b["G64"]), ptr(g4), 32, ptr(g_("dec.conv5t.w")), st), "conv5t wgrad")
        g = _dense64_geom(B * 1024, 32)
        _lib.check(lib.gccvae_ls_bf16(C.byref(g), ptr(b["G64"]), ptr(self.wp["dec.conv5t.c4"]), None, ACT_NONE, ptr(g4),
                                      ptr(b["dec.conv4t.dout"]), 0, st), "conv5t dgrad")
        prev_of = {"dec.conv4t": "dec.conv3t", "dec.conv3t": "dec.conv2t", "dec.conv2t": "dec.conv1t"}
        for name in reversed(TC_DEC):
            geom = make_geom(_DEC[name], B)
            dout = b[name + ".dout"]
            pn = prev_of[name]
            xin = b[pn + ".out16"] if pn == "dec.conv1t" else b[pn + ".out"]
            dxin = b[pn + ".dout16"] if pn == "dec.conv1t" else b[pn + ".dout"]
            _lib.check(lib.gccvae_wg_bf16(C.byref(geom), ptr(dout), ptr(xin), ptr(g_(name + ".w")), st), name + " wgrad")
            self._bias_grad16(dout, name)
            _lib.check(lib.gccvae_ls_bf16(C.byref(geom), ptr(dout), ptr(self.wp[name + ".ls"]), None, ACT_NONE, ptr(xin),
                                          ptr(dxin), 0, st), name + " dgrad")
        self._cast32(b["dec.conv1t.dout16"], b["dec.conv1t.dout"])
        # conv1t and fc1 on the fp32 CUDA-core path
        self.layer_bwd(DEC_LAYERS[1], B, b["dec.fc1.out"], b["dec.conv1t.dout"], b["dec.fc1.dout"], b["dec.fc1.out"],
                       ws, wsb)
        self.layer_bwd(DEC_LAYERS[0], B, z, b["dec.fc1.dout"], b["dz"] if want_dz else None, None, ws, wsb)
        return b["dz"]

    def encoder_bwd(self, x, b):
        B = x.shape[0]
        lib, st, g_ = self.lib, _stream(), self.store.g
        ws, wsb = b["ws"], b["ws_bytes"]
        h5, dh5 = b["enc.conv5.out"], b["enc.conv5.dout"]
        self.layer_bwd(HEAD_LAYERS[0], B, h5, b["enc.locs.dout"], dh5, None, ws, wsb)
        self.layer_bwd(HEAD_LAYERS[1], B, h5, b["enc.std.dout"], dh5, h5, ws, wsb, accumulate=True)
        self._cast16(dh5, b["enc.conv5.dout16"])
        prev_of = {"enc.conv5": "enc.conv4", "enc.conv4": "enc.conv3", "enc.conv3": "enc.conv2",
                   "enc.conv2": "enc.conv1"}
        for name in reversed(TC_ENC):
            geom = make_geom(_ENC[name], B)
            dout = b["enc.conv5.dout16"] if name == "enc.conv5" else b[name + ".dout"]
            pn = prev_of[name]
            xin, dxin = b[pn + ".out"], b[pn + ".dout"]
            _lib.check(lib.gccvae_wg_bf16(C.byref(geom), ptr(xin), ptr(dout), ptr(g_(name + ".w")), st), name + " wgrad")
            self._bias_grad16(dout, name)
            _lib.check(lib.gccvae_sl_bf16(C.byref(geom), ptr(dout), ptr(self.wp[name + ".sl"]), None, ACT_NONE, ptr(xin),
                                          ptr(dxin), 0, st), name + " dgrad")
        dh1 = b["enc.conv1.dout"]
        _lib.check(lib.gccvae_wg_c4_bf16(B * 1024, ptr(b["X64"]), ptr(dh1), 32, ptr(g_("enc.conv1.w")), st), "conv1 wgrad")
        self._bias_grad16(dh1, "enc.conv1")
