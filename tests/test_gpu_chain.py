"""GPU: the fused dense-chain kernels (csrc/chain.cu: conv5 output -> heads -> latent -> fc1 -> conv1t and the reverse,
one launch each way) against the layer-by-layer kernels of the same engine on the same inputs and the same explicit
noise.  Both formulations round to bf16 at the same tensor boundaries, so everything the step produces agrees to
fp32 summation order (forward) / to the bf16 ulp of a few borderline roundings (gradients)."""
import pytest
import torch

import gccvae_oracle as O
from helpers import cfg_for

pytestmark = pytest.mark.gpu


def _pair(mode, B, K, supervised, seed=0):
    import gccvae_b200 as G
    cfg = cfg_for(mode, "0.2")
    p = O.init_params(seed, trained_like=True)
    x, y, noise = O.make_inputs(B, k=K)
    outs = []
    for fused in (True, False):
        lrn = G.Learner((64, 64, 3), 45, 18, 18, 1000, 1.0, cfg, precision="bf16", engine_options=dict(fused_chain=fused))
        assert lrn.engine.chain == fused
        lrn.store.load_dict(p)
        loss, c = lrn.loss_and_grads(x, y, supervised, noise=noise, k=K)
        torch.cuda.synchronize()
        b = lrn.engine.bufs(B)
        last = {k: (v.detach().float().cpu().clone() if torch.is_tensor(v) else v) for k, v in lrn.last.items()}
        bufs = {k: b[k].detach().float().cpu().clone() for k in ("pre96", "z16", "dec.fc1.out", "dec.conv1t.out",
                                                                   "dec.fc1.dout", "dpre16", "enc.conv5.dout")}
        grads = {k: lrn.store.g(k).detach().cpu().clone() for k in lrn.store.names()}
        outs.append((float(loss), c.cpu().clone(), last, bufs, grads))
    return outs


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("mode,supervised,B,K", [
    ("inferred", True, 1024, 100),     # BASELINE.json configs[1]: 8 images per CTA, 2 warps per image
    ("inferred", False, 1024, 100),
    ("learnable", True, 200, 100),     # configs[0]'s batch: 2 images per CTA, 8 warps per image
    ("learnable", False, 200, 100),
    ("one-one", True, 37, 7),          # ragged last group, K smaller than the warps sharing it
    ("learnable", True, 3, 1),
    ("learnable", False, 5, 100),
])
def test_fused_chain_equals_the_layer_by_layer_kernels(mode, supervised, B, K):
    (lf, cf, lastf, bf, gf), (lu, cu, lastu, bu, gu) = _pair(mode, B, K, supervised)
    assert torch.equal(cf, cu)
    # forward: the heads' pre-activations differ only by the summation order of the k-split
    assert _rel(bf["pre96"], bu["pre96"]) < 2e-6
    for k in ("post_locs", "post_scales", "z", "kl", "log_qy_zc", "log_qy_x", "w", "log_py"):
        assert _rel(lastf[k], lastu[k]) < 2e-5, k
    if not supervised:
        assert torch.equal(lastf["y"], lastu["y"])
    # bf16 tensors between the layers: a borderline rounding may move single entries by one bf16 ulp
    for k, tol in (("z16", 1e-2), ("dec.fc1.out", 1e-2), ("dec.conv1t.out", 2e-2)):
        d = (bf[k] - bu[k]).abs()
        assert float(d.max()) <= tol * float(bu[k].abs().max()), k
        assert float((d > 0).float().mean()) < 0.02, k
    assert _rel(lastf["log_pxz"], lastu["log_pxz"]) < 2e-3
    assert abs(lf - lu) <= 2e-4 * abs(lu)
    # backward: tensors handed to the next kernels, then every parameter gradient
    for k in ("dec.fc1.dout", "dpre16", "enc.conv5.dout"):
        num = float((bf[k] - bu[k]).norm())
        assert num <= 2e-2 * float(bu[k].norm()), (k, num / float(bu[k].norm()))
    worst = ("", 0.0)
    for name in gu:
        den = float(gu[name].double().norm())
        if den == 0.0:
            assert float(gf[name].abs().max()) == 0.0, name
            continue
        e = float((gf[name].double() - gu[name].double()).norm()) / den
        worst = max(worst, (name, e), key=lambda t: t[1])
    print("fused chain vs layer-by-layer: worst gradient", worst)
    assert worst[1] < 2e-2, worst


def test_fused_chain_train_step_graph_replay_matches_eager():
    """the captured step (CUDA graph) and the eager step of the fused engine walk the same trajectory."""
    import gccvae_b200 as G
    cfg = dict(cfg_for("learnable", "0.5"), batch_size=64)
    g = torch.Generator().manual_seed(3)
    x = torch.randint(0, 256, (64, 64, 64, 3), generator=g, dtype=torch.uint8).cuda()
    y = (torch.rand(64, 18, generator=g) < 0.5).long().cuda()
    outs = []
    for graphs in (False, True):
        lrn = G.Learner((64, 64, 3), 45, 18, 18, 1000, 0.2, cfg, precision="bf16", graphs=graphs, seed=5)
        lrn.store.load_dict(O.init_params(0, trained_like=True))
        losses = []
        for i in range(4):
            loss, _ = lrn.train_step(x, y if i % 2 == 0 else None, i % 2 == 0)
            losses.append(loss)
        torch.cuda.synchronize()
        outs.append(([float(v) for v in losses], lrn.store.flat.clone()))
    for a, b in zip(outs[0][0], outs[1][0]):
        assert abs(a - b) <= 1e-3 * abs(a), (outs[0][0], outs[1][0])
    # Adam's first steps move every parameter by ~lr whatever the gradient's size, so entries whose gradient is at the
    # level of the atomics' summation noise may differ by O(lr).  Measured: the two runs are usually bit-identical; when
    # the order of the bias-gradient atomics of one step differs (one ulp in dec.conv1t.b / dec.fc1.b), 5 % of the
    # entries differ after that step and 10 % after the next one.  A wrong step moves (almost) all of them.
    differs = ((outs[0][1] - outs[1][1]).abs() > 2e-6).float().mean().item()
    assert differs < 0.3, differs
