"""GPU: bookkeeping around the one-launch Adam (gccvae_adam_fused_f32): the replayed step carries no gradient memset and
relies on the previous update having cleared the buffer; a forward/backward call in between (dirty gradients) must not
leak into the next train_step; the two-part update (bulk under the last dgrad + first layer at the end) and the plain
two-launch update give bit-identical parameters."""
import pytest
import torch

import gccvae_oracle as O
from helpers import cfg_for

pytestmark = pytest.mark.gpu


def _learner(graphs, seed=5):
    import gccvae_b200 as G
    cfg = dict(cfg_for("learnable", "0.5"), batch_size=32)
    lrn = G.Learner((64, 64, 3), 45, 18, 18, 1000, 0.2, cfg, precision="bf16", graphs=graphs, seed=seed)
    lrn.store.load_dict(O.init_params(0, trained_like=True))
    return lrn


def _data(n=32):
    g = torch.Generator().manual_seed(3)
    x = torch.randint(0, 256, (n, 64, 64, 3), generator=g, dtype=torch.uint8).cuda()
    y = (torch.rand(n, 18, generator=g) < 0.5).long().cuda()
    return x, y


@pytest.mark.parametrize("graphs", [False, True])
def test_dirty_gradients_do_not_leak_into_the_next_step(graphs):
    x, y = _data()
    a, b = _learner(graphs), _learner(graphs)
    for lrn in (a, b):                       # bring both to the same state (and capture both graph variants)
        for _ in range(2):
            lrn.train_step(x, y, True)
            lrn.train_step(x, None, False)
    torch.cuda.synchronize()
    assert ((a.store.flat - b.store.flat).abs() > 2e-6).float().mean().item() < 0.2
    b.loss_and_grads(x, y, True)             # leaves gradients in the buffer, no update, no step-counter change
    assert not b._grads_clean
    a.train_step(x, y, True)
    b.train_step(x, y, True)
    a.train_step(x, None, False)
    b.train_step(x, None, False)
    torch.cuda.synchronize()
    assert a.optimiser.iterations == b.optimiser.iterations == 6
    # weight gradients are accumulated with red.global (order varies run to run) and Adam's first steps move every
    # parameter by ~lr = 1e-4 whatever the gradient's size: entries whose gradient is at the level of the summation
    # noise may differ by O(lr), all others agree to ~1e-9.  Leaked gradients would shift (almost) EVERY entry by >1e-6.
    differs = ((a.store.flat - b.store.flat).abs() > 2e-6).float().mean().item()
    assert differs < 0.2, differs      # (atomics-order noise reaches ~0.1 after a few Adam steps; a leak moves ~every entry)
    assert float(b.store.grad.abs().max()) == 0.0 and b._grads_clean


def test_two_part_update_equals_the_plain_update():
    import gccvae_b200 as G
    x, y = _data()
    outs = []
    for tail in (True, False):
        cfg = dict(cfg_for("learnable", "0.5"), batch_size=32)
        lrn = G.Learner((64, 64, 3), 45, 18, 18, 1000, 0.2, cfg, precision="bf16", graphs=False, seed=5, adam_tail=tail)
        lrn.store.load_dict(O.init_params(0, trained_like=True))
        for _ in range(2):
            lrn.train_step(x, y, True)
            lrn.train_step(x, None, False)
        torch.cuda.synchronize()
        outs.append((lrn.store.flat.clone(), lrn.optimiser.m.clone(), lrn.optimiser.iterations))
    assert outs[0][2] == outs[1][2] == 4
    differs = ((outs[1][0] - outs[0][0]).abs() > 2e-6).float().mean().item()
    assert differs < 0.2, differs


@pytest.mark.parametrize("graphs", [False, True])
def test_train_step_results_are_not_overwritten_by_the_next_steps(graphs):
    """gated_ccvae.py:311 returns fresh tensors; here the publishing Adam launch copies (loss, c) into a ring, so what
    train_step handed back must still hold its own step's values after the following steps ran."""
    x, y = _data()
    lrn = _learner(graphs)
    kept, values = [], []
    for i in range(6):
        loss, c = lrn.train_step(x, y if i % 2 == 0 else None, i % 2 == 0)
        torch.cuda.synchronize()
        kept.append((loss, c))
        values.append((float(loss), c.clone()))
    assert len({round(v, 4) for v, _ in values}) > 2          # the steps really differ
    for (loss, c), (v, cv) in zip(kept, values):
        assert float(loss) == v and torch.equal(c, cv)
