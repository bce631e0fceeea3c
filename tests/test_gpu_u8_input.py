"""uint8 images through the Learner (bf16 engine): the raw 0..255 pixels the reference's loader holds before its
/255 (utils_data.py:56-59) give bit-identical results to passing the normalised fp32 image, eager and graphed."""
import numpy as np
import pytest
import torch

import gccvae_oracle as O
from helpers import make_learner

pytestmark = pytest.mark.gpu


def _cfg():
    import os
    root = os.path.dirname(os.path.abspath(__file__))
    mu0 = np.load(os.path.join(root, "golden", "data", "gating_matrix_0.5.npy"))
    return dict(gate_type="learnable", gate_subtype=None, mu_init=mu0, gating_reg=0.2, lr=1e-4, gating_init_temp=1.0,
                batch_size=16)


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
@pytest.mark.parametrize("supervised", [True, False])
def test_uint8_input_equals_normalised_fp32(precision, supervised):
    cfg = _cfg()
    p = O.init_params(0, trained_like=True)
    B = 16
    g = torch.Generator().manual_seed(11)
    xu = torch.randint(0, 256, (B, 64, 64, 3), generator=g, dtype=torch.uint8)
    xf = torch.from_numpy(xu.numpy().astype(np.float32) / 255.0)
    _, y, noise = O.make_inputs(B, k=20)
    out = []
    for x in (xu, xf):
        lrn = make_learner(cfg, p, precision=precision)
        loss, _ = lrn.loss_and_grads(x, y, supervised, noise=noise, k=20)
        torch.cuda.synchronize()
        out.append((float(loss), lrn.store.grad.clone(), lrn.last["log_pxz"].clone()))
    # identical inputs -> identical arithmetic; only the order of the fp32 atomic accumulations (split-K weight
    # gradients, per-image likelihood sums) may differ between two runs
    assert abs(out[0][0] - out[1][0]) <= 1e-6 * abs(out[1][0]), "loss differs between uint8 and fp32 input"
    assert float((out[0][2] - out[1][2]).abs().max()) <= 1e-5 * float(out[1][2].abs().max())
    gdiff = float((out[0][1] - out[1][1]).abs().max()) / float(out[1][1].abs().max())
    assert gdiff <= 1e-5, "gradients differ between uint8 and fp32 input: {:.2e}".format(gdiff)


def test_uint8_graphed_train_step_runs_and_matches_fp32_graph():
    cfg = _cfg()
    p = O.init_params(0, trained_like=True)
    B = 16
    g = torch.Generator().manual_seed(12)
    xu = torch.randint(0, 256, (B, 64, 64, 3), generator=g, dtype=torch.uint8)
    xf = torch.from_numpy(xu.numpy().astype(np.float32) / 255.0)
    y = (torch.rand(B, 18, generator=g) < 0.5).long()
    res = []
    for x in (xu.pin_memory(), xf.pin_memory()):
        lrn = make_learner(cfg, p, precision="bf16", graphs=True, seed=77)
        for _ in range(3):
            loss, _ = lrn.train_step(x, y, True)
            loss, _ = lrn.train_step(x, None, False)
        torch.cuda.synchronize()
        res.append((float(loss), lrn.store.flat.clone()))
    assert np.isfinite(res[0][0])
    assert abs(res[0][0] - res[1][0]) <= 1e-4 * abs(res[1][0])
    # Adam moves a parameter by at most lr = 1e-4 per step; the two runs may differ by atomic-accumulation order only
    assert float((res[0][1] - res[1][1]).abs().max()) < 1e-4, "parameters after 6 graphed steps differ"
