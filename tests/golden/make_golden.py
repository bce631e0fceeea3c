"""Generate tests/golden/oracle_golden.npz: outputs of the CPU oracle (fp64) on small seeded cases.

The reference itself cannot run here (TensorFlow / TFP / Keras absent), so these are golden vectors
OF THE ORACLE: they pin the oracle against accidental change and travel to the GPU box, where the CUDA path is
compared with them without re-deriving anything.  Re-generate with:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gccvae_oracle as O  # noqa: E402
from helpers import cfg_for  # noqa: E402

CASES = [  # name, mode, frac, supervised, B, K, T
    ("one_one_sup", "one-one", "0.5", True, 4, 8, 0.3),
    ("one_one_unsup", "one-one", "0.5", False, 4, 8, 0.3),
    ("inferred02_sup", "inferred", "0.2", True, 4, 8, 0.3),
    ("learnable05_sup", "learnable", "0.5", True, 4, 8, 1.0),
    ("learnable10_unsup", "learnable", "1.0", False, 4, 8, 1.0),
]
TERMS = ["loss", "c", "post_locs", "post_scales", "z", "logits", "log_qy_zc", "log_py", "kl", "log_pxz"]
PROBE = 64  # gradient entries sampled per tensor


def run_case(mode, frac, supervised, B, K, T):
    cfg = cfg_for(mode, frac)
    p = O.init_params(0, dtype=torch.float64, trained_like=True)
    mu, _ = O.initialise_mu(cfg, dtype=torch.float64)
    x, y, noise = O.make_inputs(B, k=K, dtype=torch.float64)
    return O.loss_and_grads(p, mu, x, y, noise, cfg, T, supervised)


def main():
    out = {}
    for name, mode, frac, sup, B, K, T in CASES:
        o, g = run_case(mode, frac, sup, B, K, T)
        for t in TERMS + (["log_qy_x", "w"] if sup else ["y"]):
            out["{}/{}".format(name, t)] = o[t].numpy()
        rng = np.random.RandomState(0)
        for k, v in g.items():
            if v is None:
                continue
            flat = v.reshape(-1).numpy()
            idx = rng.choice(flat.size, size=min(PROBE, flat.size), replace=False)
            out["{}/grad_idx/{}".format(name, k)] = idx.astype(np.int64)
            out["{}/grad_val/{}".format(name, k)] = flat[idx]
            out["{}/grad_norm/{}".format(name, k)] = np.array(np.linalg.norm(flat))
    np.savez_compressed(os.path.join(HERE, "oracle_golden.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
