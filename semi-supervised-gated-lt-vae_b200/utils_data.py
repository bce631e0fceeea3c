"""Gating-matrix files: the part of the reference's data module that is on the ELBO path.
Mirrors CelebAReader.set_gating_prob's load branch (utils_data.py:147-152) and the learned-mu
load of Learner.load_model (gated_ccvae.py:161-165).  Row index = z_c dim i, column = label j;
no transpose on load; float64 on disk, cast to fp32 by CCVAE.initialise_mu."""
from __future__ import annotations

import os

import numpy as np

CELEBA_EASY_LABELS = ['Arched_Eyebrows', 'Bags_Under_Eyes', 'Bangs', 'Black_Hair', 'Blond_Hair', 'Brown_Hair',
                      'Bushy_Eyebrows', 'Chubby', 'Eyeglasses', 'Heavy_Makeup', 'Male', 'No_Beard', 'Pale_Skin',
                      'Receding_Hairline', 'Smiling', 'Wavy_Hair', 'Wearing_Necktie', 'Young']


def gating_matrix_path(root, sup_frac):
    return os.path.join(root, "gating_matrix_{}.npy".format(sup_frac))


def load_gating_matrix(root, sup_frac):
    """-> float64 ndarray [18,18] exactly as stored (utils_data.py:149-152)."""
    path = gating_matrix_path(root, sup_frac)
    if not os.path.exists(path):
        raise FileNotFoundError("No gating matrix found at {} (generation from label co-occurrence "
                                "is outside the ELBO path)".format(path))
    mu = np.load(path)
    if mu.ndim != 2 or mu.shape[0] != mu.shape[1]:
        raise ValueError("gating matrix {} has shape {}".format(path, mu.shape))
    return mu


def load_learned_gating_matrix(param_dir, model_id):
    """gated_ccvae.py:161-163: learned mu saved by training ('best' | 'last')."""
    return np.load(os.path.join(param_dir, "learned_gating_matrix_{}.npy".format(model_id)))


class GatingMatrixReader:
    """The slice of CelebAReader the Learner consumes: `.init_gating_prob` (gated_ccvae.py:506)."""

    def __init__(self, root, sup_frac, batch_size=None):
        self.root, self.sup_frac, self.batch_size = root, sup_frac, batch_size
        self.init_gating_prob = load_gating_matrix(root, sup_frac)


class SyntheticReader:
    """Stand-in for CelebAReader (utils_data.py:31-176) with the interface the Learner consumes: `.n_s` samples,
    `.step()` yielding `(xs, ys)` batches forever (xs uint8 or fp32 [B,64,64,3], ys int64 [B,18] or None).  CelebA
    itself is not available; images are seeded synthetic 8-bit pixels, labels Bernoulli(0.5)."""

    def __init__(self, n_samples, batch_size, supervised=True, seed=0, dtype="uint8", y_dim=18):
        import torch
        self.n_s, self.batch_size, self.supervised = int(n_samples), int(batch_size), bool(supervised)
        gen = torch.Generator().manual_seed(seed)
        self.x = torch.randint(0, 256, (self.n_s, 64, 64, 3), generator=gen, dtype=torch.uint8)
        if dtype != "uint8":
            self.x = self.x.to(torch.float32) / 255.0
        self.y = (torch.rand(self.n_s, y_dim, generator=gen) < 0.5).to(torch.int64)

    def step(self):
        i = 0
        while True:
            idx = [(i + j) % self.n_s for j in range(self.batch_size)]
            i = (i + self.batch_size) % self.n_s
            yield self.x[idx], (self.y[idx] if self.supervised else None)
