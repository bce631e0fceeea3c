"""B200-native Gated-CCVAE ELBO training step (drop-in for the hot path of
jabhinav/Semi-Supervised-Gated-LT-VAE: gated_ccvae.py over networks.py).

The directory name carries hyphens, so import it through the root shim:  `import gccvae_b200`.
"""
from ._lib import GccvaeError, LIB_PATH, load as load_library  # noqa: F401
from .gated_ccvae import CCVAE, KerasAdam, Learner  # noqa: F401
from .networks import Classifier, Conditional_Prior, Decoder, Encoder  # noqa: F401
from .utils import get_gaussian_kl_div, img_log_likelihood  # noqa: F401
from .utils_data import (CELEBA_EASY_LABELS, CELEBA_LABELS, CelebAReader, DataLoader, GatingMatrixReader,  # noqa: F401
                         SyntheticReader, create_gating_matrix, load_gating_matrix, load_learned_gating_matrix)

__all__ = ["CCVAE", "Learner", "KerasAdam", "Encoder", "Decoder", "Classifier", "Conditional_Prior",
           "img_log_likelihood", "get_gaussian_kl_div", "load_gating_matrix", "load_learned_gating_matrix",
           "GatingMatrixReader", "CelebAReader", "DataLoader", "SyntheticReader", "create_gating_matrix",
           "CELEBA_EASY_LABELS", "CELEBA_LABELS", "GccvaeError", "load_library", "LIB_PATH"]
