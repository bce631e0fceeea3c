"""The block ("x2" / "s2d") identities the tensor-core kernels rest on, in numpy against the oracle's convolutions
(oracle/block_forms.py): stride-2 convolution as a 2x2-tap gather over blocks, and transposed convolution producing
its output directly in block form - the formulation planned for the S -> L layers."""
import numpy as np
import pytest
import torch

import block_forms as BF
import gccvae_oracle as O


@pytest.mark.parametrize("n,h,cl,cs,seed", [(2, 8, 3, 5, 0), (1, 32, 32, 32, 1), (3, 4, 16, 8, 2)])
def test_stride2_conv_is_a_2x2_tap_gather_over_blocks(n, h, cl, cs, seed):
    rng = np.random.default_rng(seed)
    L = rng.standard_normal((n, h, h, cl))
    W = rng.standard_normal((4, 4, cl, cs))                       # Keras Conv2D kernel [kh, kw, Cin, Cout]
    want = O._conv(torch.from_numpy(L), torch.from_numpy(W), None, 2, 1).numpy()
    blocks = BF.to_blocks(L)
    assert blocks.shape == (n, h // 2 + 1, h // 2 + 1, 4, cl)
    assert np.array_equal(BF.from_blocks(blocks, h, h), L)
    got = BF.conv_k4s2p1_from_blocks(blocks, W)
    assert got.shape == want.shape and np.abs(got - want).max() < 1e-10 * max(1.0, np.abs(want).max())


@pytest.mark.parametrize("n,hs,cl,cs,seed", [(2, 4, 3, 5, 0), (1, 16, 32, 32, 1), (2, 8, 32, 64, 2)])
def test_transposed_conv_produces_its_output_in_block_form(n, hs, cl, cs, seed):
    rng = np.random.default_rng(seed)
    S = rng.standard_normal((n, hs, hs, cs))
    W = rng.standard_normal((4, 4, cl, cs))                       # Keras Conv2DTranspose kernel [kh, kw, Cout, Cin]
    want = O._convT(torch.from_numpy(S), torch.from_numpy(W), None, 2, 1).numpy()       # [n, 2hs, 2hs, cl]
    blocks = BF.convT_k4s2_to_blocks(S, W)
    assert blocks.shape == (n, hs + 1, hs + 1, 4, cl)
    assert np.abs(blocks - BF.to_blocks(want)).max() < 1e-10 * max(1.0, np.abs(want).max())
    assert np.abs(BF.from_blocks(blocks, 2 * hs, 2 * hs) - want).max() < 1e-10 * max(1.0, np.abs(want).max())


def test_block_gemm_operand_of_the_S_to_L_form():
    """[pixels, (a, b, cs)] x [(a, b, cs), (dy, dx, cl)]: one GEMM with K = 4 C_S and N = 4 C_L per block row."""
    rng = np.random.default_rng(3)
    n, hs, cl, cs = 1, 4, 8, 16
    S = rng.standard_normal((n, hs, hs, cs))
    W = rng.standard_normal((4, 4, cl, cs))
    Sp = np.zeros((n, hs + 2, hs + 2, cs))
    Sp[:, 1:-1, 1:-1] = S
    A = np.zeros((n, hs + 1, hs + 1, 2, 2, cs))
    for a in (0, 1):
        for b in (0, 1):
            A[:, :, :, a, b] = Sp[:, 1 - a:1 - a + hs + 1, 1 - b:1 - b + hs + 1]
    out = (A.reshape(-1, 4 * cs) @ BF.gemm_operands_S_to_L(W)).reshape(n, hs + 1, hs + 1, 4, cl)
    want = BF.to_blocks(O._convT(torch.from_numpy(S), torch.from_numpy(W), None, 2, 1).numpy())
    inside = BF.to_blocks(np.ones((n, 2 * hs, 2 * hs, 1)))[..., 0] > 0          # slots that lie inside the image
    assert np.abs(out - want)[inside].max() < 1e-10
    assert np.abs(out[~inside]).max() > 0      # the raw GEMM does write the border slots: the epilogue has to mask them
