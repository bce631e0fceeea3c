"""GPU: the reference's trained checkpoint (`models/params_0.2_learnable/*_best.*`, Keras .h5 + learned mu) loaded
through `Learner.load_model` gives the oracle's losses and gradients (trained weights have a different scale
structure from fresh initialisers: large decoder biases, a learned gating matrix with entries outside [0,1]); and the
training loop (`Learner.train`, gated_ccvae.py:313-419) runs the reference's schedule, checkpoints and decays T."""
import os

import numpy as np
import pytest
import torch

import gccvae_oracle as O
from helpers import assert_close, cfg_for, rel_err

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
CKPT = os.path.join(HERE, "golden", "models", "params_0.2_learnable")


def _trained_learner(precision, **kw):
    import gccvae_b200 as G
    cfg = cfg_for("learnable", "0.2")
    lrn = G.Learner((64, 64, 3), 45, 18, 18, 1000, 0.2, cfg, precision=precision, **kw)
    lrn.load_model(CKPT, "best")
    return lrn, cfg


def _l2(got, want):
    got, want = torch.as_tensor(got).double().cpu(), torch.as_tensor(want).double().cpu()
    return float((got - want).norm() / want.norm().clamp_min(1e-30))


def test_load_model_fills_every_parameter_from_the_h5_files():
    from gccvae_b200.h5lite import keras_weights
    lrn, _ = _trained_learner("fp32")
    enc = keras_weights(os.path.join(CKPT, "encoder_model_best.h5"))
    assert torch.equal(lrn.store.view("enc.conv1.w").cpu(), torch.from_numpy(enc[0][1]))
    assert torch.equal(lrn.store.view("enc.std.b").cpu(), torch.from_numpy(enc[-1][1]))
    pri = keras_weights(os.path.join(CKPT, "cond_prior_best.h5"))
    assert torch.equal(lrn.store.view("prior.scale_false").cpu(), torch.from_numpy(pri[3][1]))
    mu = np.load(os.path.join(CKPT, "learned_gating_matrix_best.npy"))
    assert torch.equal(lrn.store.view("mu").cpu(), torch.from_numpy(mu))


@pytest.mark.parametrize("supervised", [True, False])
def test_trained_checkpoint_fp32_matches_oracle(supervised):
    lrn, cfg = _trained_learner("fp32")
    p = {k: v.cpu() for k, v in lrn.store.to_dict().items()}
    mu = p.pop("mu")
    x, y, noise = O.make_inputs(16, k=20)
    T = cfg["gating_init_temp"]
    p64 = {k: v.double() for k, v in p.items()}
    n64 = {k: v.double() for k, v in noise.items()}
    o, g = O.loss_and_grads(p64, mu.double(), x.double(), y, n64, cfg, T, supervised)
    loss, c = lrn.loss_and_grads(x, y, supervised, noise=noise, k=20)
    torch.cuda.synchronize()
    assert_close(c, o["c"], 1e-5, "c")
    assert_close(loss, o["loss"], 2e-5, "loss")
    for k in ["kl", "log_pxz"] + (["log_qy_zc", "log_qy_x", "w"] if supervised else []):
        assert_close(lrn.last[k], o[k], 4e-5, k)
    for name in ["enc.conv1.w", "enc.conv4.w", "dec.conv2t.w", "dec.conv5t.b", "cls.w", "prior.scale_true", "mu"]:
        assert_close(lrn.store.g(name), g[name], 1e-4, "grad " + name)


@pytest.mark.parametrize("supervised", [True, False])
def test_trained_checkpoint_bf16_matches_oracle(supervised):
    lrn, cfg = _trained_learner("bf16")
    p = {k: v.cpu() for k, v in lrn.store.to_dict().items()}
    mu = p.pop("mu")
    x, y, noise = O.make_inputs(32, k=100)
    T = cfg["gating_init_temp"]
    p64 = {k: v.double() for k, v in p.items()}
    n64 = {k: v.double() for k, v in noise.items()}
    o, _ = O.loss_and_grads(p64, mu.double(), x.double(), y, n64, cfg, T, supervised)
    with O.bf16_emulation():
        _, gem = O.loss_and_grads(p64, mu.double(), x.double(), y, n64, cfg, T, supervised)
    loss, c = lrn.loss_and_grads(x, y, supervised, noise=noise, k=100)
    torch.cuda.synchronize()
    assert_close(loss, o["loss"], 1e-2, "loss")
    for k in ["post_locs", "post_scales", "kl", "log_pxz"]:
        assert_close(lrn.last[k], o[k], 1e-2, k)
    for name in ["enc.conv2.w", "enc.conv5.w", "dec.conv3t.w", "dec.conv1t.w"]:
        e = _l2(lrn.store.g(name), gem[name])
        assert e < 5e-2, "grad {}: rel L2 {:.3e}".format(name, e)


@pytest.mark.parametrize("precision,graphs", [("fp32", False), ("bf16", True)])
def test_train_loop_schedule_checkpoints_and_temperature(tmp_path, precision, graphs):
    import gccvae_b200 as G
    from gccvae_b200.utils_data import SyntheticReader
    cfg = dict(cfg_for("learnable", "0.5"), perc_supervision=0.2, n_epochs=2, batch_size=16)
    lrn = G.Learner((64, 64, 3), 45, 18, 18, 200, 0.2, cfg, precision=precision, graphs=graphs)
    loaders = dict(sup=SyntheticReader(32, 16, True, seed=1), unsup=SyntheticReader(96, 16, False, seed=2),
                   valid=SyntheticReader(32, 16, True, seed=3))
    seen = []
    before = lrn.store.flat.clone()
    hist = lrn.train(loaders, str(tmp_path), on_batch=lambda e, i, s, loss, c: seen.append((e, i, s)))
    sched = G.Learner.epoch_schedule(0.2, 32, 96, 16)          # 2 sup + 6 unsup batches, period 4
    assert sched == [True, False, False, False, True, False, False, False]
    assert [s for e, i, s in seen if e == 0] == sched and len(seen) == 2 * len(sched)
    assert int(lrn.optimiser.iterations) == 2 * len(sched)
    assert lrn.gating_sampler_temp == pytest.approx(1.0 * 0.99 * 0.99)
    assert len(hist) == 2 and all(np.isfinite(h["sup_loss"]) and np.isfinite(h["unsup_loss"]) for h in hist)
    assert 0.0 <= hist[-1]["val_acc"] <= 1.0
    assert not torch.equal(before, lrn.store.flat) and bool(torch.isfinite(lrn.store.flat).all())
    for mid in ("best", "last"):
        for stem in ("encoder_model", "decoder_model", "classifier", "cond_prior"):
            assert os.path.exists(os.path.join(str(tmp_path), "{}_{}.h5".format(stem, mid)))
        assert os.path.exists(os.path.join(str(tmp_path), "learned_gating_matrix_{}.npy".format(mid)))
        assert os.path.exists(os.path.join(str(tmp_path), "learned_gating_matrix_{}.csv".format(mid)))
    # the saved "last" model restores the exact parameters
    other = G.Learner((64, 64, 3), 45, 18, 18, 200, 0.2, cfg, precision=precision)
    other.load_model(str(tmp_path), "last")
    assert torch.equal(other.store.flat, lrn.store.flat)
