"""GPU: API behaviour around the hot path - differentiable sup_loss / unsup_loss (the reference takes tape.gradient of
their result, gated_ccvae.py:302-309), fresh noise for every forward-only call (the reference's tf.random draws), the
gate temperature as a device scalar that reaches captured graphs (gated_ccvae.py:404-406)."""
import pytest
import torch

import gccvae_oracle as O
from helpers import cfg_for, make_learner

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("mode", ["learnable", "inferred"])
def test_losses_are_differentiable_like_under_the_reference_tape(precision, mode):
    cfg = cfg_for(mode, "0.5")
    p = O.init_params(0, trained_like=True)
    x, y, noise = O.make_inputs(16, k=20)
    ref = make_learner(cfg, p, precision=precision)
    want = {}
    for sup in (True, False):
        loss, _ = ref.loss_and_grads(x, y, sup, noise=noise, k=20)
        torch.cuda.synchronize()
        want[sup] = (float(loss), ref.store.grad.clone())
    lrn = make_learner(cfg, p, precision=precision).requires_grad_()
    for sup in (True, False):
        loss, c = (lrn.sup_loss(x, y, noise=noise, k=20) if sup else lrn.unsup_loss(x, noise=noise))
        assert loss.requires_grad and not c.requires_grad
        assert abs(float(loss.detach()) - want[sup][0]) <= 1e-5 * abs(want[sup][0])
        lrn.store.flat.grad = None
        (2.0 * loss).backward()
        g = lrn.store.flat.grad
        w = 2.0 * want[sup][1]
        if not lrn.model.mu_trainable:
            assert float(lrn.gradients()["mu"].abs().max()) == 0.0
            w = w.clone()
            lrn.store.view("mu", w).zero_()
        # (bf16: weight gradients are accumulated with atomics, the two runs differ in summation order only)
        assert float((g - w).norm() / w.norm()) < (1e-6 if precision == "fp32" else 1e-4)
        assert set(lrn.gradients()) == set(lrn.store.names())
    # without requires_grad the calls stay forward-only and graph-less
    lrn.requires_grad_(False)
    loss, _ = lrn.unsup_loss(x, noise=noise)
    assert not loss.requires_grad and lrn.gradients() is None
    with torch.no_grad():
        assert not lrn.requires_grad_().unsup_loss(x, noise=noise)[0].requires_grad


def test_forward_only_calls_draw_fresh_noise_and_train_steps_stay_reproducible():
    cfg = cfg_for("inferred", "0.2")           # stochastic gates at T = 0.3
    p = O.init_params(0, trained_like=True)
    x, y, _ = O.make_inputs(16, k=10)
    a = make_learner(cfg, p, precision="bf16", seed=3)
    l1, c1 = a.unsup_loss(x)
    l2, c2 = a.unsup_loss(x)
    acc = [float(a.classifier_accuracy(x, y)) for _ in range(3)]
    torch.cuda.synchronize()
    assert not torch.equal(c1, c2) and float(l1) != float(l2)          # the reference draws a new c per call too
    assert all(0.0 <= v <= 1.0 for v in acc)
    # a second learner with the same seed that made NO forward-only calls walks the same training trajectory
    b = make_learner(cfg, p, precision="bf16", seed=3)
    for lrn in (a, b):
        lrn.train_step(x, y, True)
    torch.cuda.synchronize()
    assert torch.equal(a.last["c"], b.last["c"])


def test_gate_temperature_reaches_captured_graphs_without_recapture():
    import gccvae_b200 as G
    cfg = dict(cfg_for("learnable", "0.5"), batch_size=16, lr=0.0)      # lr 0: parameters (and mu) stay put
    g = torch.Generator().manual_seed(1)
    x = torch.randint(0, 256, (16, 64, 64, 3), generator=g, dtype=torch.uint8).cuda()
    y = (torch.rand(16, 18, generator=g) < 0.5).long().cuda()
    lrn = G.Learner((64, 64, 3), 45, 18, 18, 1000, 1.0, cfg, precision="bf16", graphs=True, seed=9)
    lrn.store.load_dict(O.init_params(0, trained_like=True))
    for _ in range(2):
        lrn.train_step(x, y, True)           # both graph variants captured at T = 1.0
    n_graphs = sum(v is not None for vs in lrn._graphs.values() for v in vs)
    eager = G.Learner((64, 64, 3), 45, 18, 18, 1000, 1.0, cfg, precision="bf16", graphs=False, seed=9)
    eager.store.load_dict(O.init_params(0, trained_like=True))
    for _ in range(2):
        eager.train_step(x, y, True)
    for lr_ in (lrn, eager):
        lr_.gating_sampler_temp = 0.3
    _, cg = lrn.train_step(x, y, True)       # replay: same graph, new temperature
    _, ce = eager.train_step(x, y, True)
    torch.cuda.synchronize()
    assert sum(v is not None for vs in lrn._graphs.values() for v in vs) == n_graphs
    assert torch.equal(cg, ce)
    # colder gates are more decided: further from 1/2 on average than the T = 1.0 sample of the same step counter
    lrn.gating_sampler_temp = 1.0
    _, c_warm = lrn.train_step(x, y, True)
    torch.cuda.synchronize()
    assert float((cg - 0.5).abs().mean()) > float((c_warm - 0.5).abs().mean())
    with pytest.raises(ValueError):
        lrn.gating_sampler_temp = 0.0


def test_inputs_staged_ahead_of_the_replay_give_the_same_steps():
    """graph replay stages a batch (copy into the variant's buffers, x2 transform) on the copy stream; with host batches,
    and with resident batches flagged `inputs_ready`, that runs under the step in flight.  Same trajectory either way."""
    import gccvae_b200 as G
    cfg = dict(cfg_for("learnable", "0.5"), batch_size=32)
    g = torch.Generator().manual_seed(2)
    xs = [torch.randint(0, 256, (32, 64, 64, 3), generator=g, dtype=torch.uint8) for _ in range(3)]
    ys = [(torch.rand(32, 18, generator=g) < 0.5).long() for _ in range(3)]
    outs = []
    for mode in ("host", "resident", "resident-ready"):
        lrn = G.Learner((64, 64, 3), 45, 18, 18, 1000, 0.2, cfg, precision="bf16", graphs=True, seed=4)
        lrn.store.load_dict(O.init_params(0, trained_like=True))
        losses = []
        for i in range(6):
            x, y = xs[i % 3], ys[i % 3]
            if mode == "host":
                x, y = x.pin_memory(), y.pin_memory()
            else:
                x, y = x.cuda(), y.cuda()
                torch.cuda.synchronize()
            sup = i % 2 == 0
            losses.append(lrn.train_step(x, y if sup else None, sup, inputs_ready=(mode == "resident-ready"))[0])
        torch.cuda.synchronize()
        outs.append([float(v) for v in losses])
    for other in outs[1:]:
        for a, b in zip(outs[0], other):
            assert abs(a - b) <= 2e-3 * abs(a), outs      # (weight gradients are accumulated with atomics)
