"""Index algebra of the block ("x2" / "s2d") formulations of the k4 / s2 / p1 layers, restated in numpy against the
oracle's own convolutions.  TEST INFRASTRUCTURE (like the rest of oracle/): only tests/ import it.

The kernels of csrc/conv_tc.cu store the large tensor L of a stride-2 layer in block form
    blocks[n, i, j, (dy, dx), c] = L[n, 2i - 1 + dy, 2j - 1 + dx, c]      i, j in [0, H/2], zero outside the image
and use two identities, both checked in tests/test_block_forms.py:

  L -> S (Conv2D k4 s2 p1, networks.py:11-15):
      S[n, i, j, cs] = sum_{a,b in {0,1}} sum_{(dy,dx), cl} blocks[n, i + a, j + b, (dy,dx), cl] * W[2a + dy, 2b + dx, cl, cs]
      a 2x2-tap stride-1 gather over blocks with K = 4 taps x 4 C_L                      (tap4_ls, c3conv kernels)

  S -> L (Conv2DTranspose k4 s2 'same', networks.py:45-49; kernel layout [kh, kw, C_L, C_S]):
      blocks[n, i, j, (dy,dx), cl] = sum_{a,b in {0,1}} sum_cs S[n, i - a, j - b, cs] * W[2a + dy, 2b + dx, cl, cs]
      for the slots that lie inside the image: a 2x2-tap stride-1 gather over S (zero outside) that produces the output
      directly in block form with N = 4 C_L - the formulation DESIGN.md lists as the next step for the S -> L layers
      (conv4t forward, conv2 dgrad), which today run on the halo kernel with 64-byte operand rows.
"""
import numpy as np


def to_blocks(L):
    """[N, H, W, C] -> [N, H/2 + 1, W/2 + 1, 4, C] (slot = 2*dy + dx), zero where the pixel is outside the image."""
    n, h, w, c = L.shape
    out = np.zeros((n, h // 2 + 1, w // 2 + 1, 4, c), dtype=L.dtype)
    for dy in (0, 1):
        for dx in (0, 1):
            ys = 2 * np.arange(h // 2 + 1) - 1 + dy
            xs = 2 * np.arange(w // 2 + 1) - 1 + dx
            vy, vx = (ys >= 0) & (ys < h), (xs >= 0) & (xs < w)
            out[:, np.ix_(vy, vx)[0], np.ix_(vy, vx)[1], 2 * dy + dx] = L[:, ys[vy]][:, :, xs[vx]]
    return out


def from_blocks(blocks, h, w):
    """inverse of to_blocks for the slots inside the image."""
    n, hb, wb, _, c = blocks.shape
    L = np.zeros((n, h, w, c), dtype=blocks.dtype)
    for dy in (0, 1):
        for dx in (0, 1):
            ys = 2 * np.arange(hb) - 1 + dy
            xs = 2 * np.arange(wb) - 1 + dx
            vy, vx = (ys >= 0) & (ys < h), (xs >= 0) & (xs < w)
            L[:, ys[vy][:, None], xs[vx][None, :]] = blocks[:, np.ix_(vy, vx)[0], np.ix_(vy, vx)[1], 2 * dy + dx]
    return L


def conv_k4s2p1_from_blocks(blocks, W):
    """L -> S.  blocks [N, HS+1, WS+1, 4, CL], W [4, 4, CL, CS] (Keras Conv2D layout) -> S [N, HS, WS, CS]."""
    n, hb, wb, _, cl = blocks.shape
    hs, ws = hb - 1, wb - 1
    S = np.zeros((n, hs, ws, W.shape[3]), dtype=np.result_type(blocks, W))
    for a in (0, 1):
        for b in (0, 1):
            tap = blocks[:, a:a + hs, b:b + ws]                           # [N, HS, WS, 4, CL]
            for dy in (0, 1):
                for dx in (0, 1):
                    S += tap[:, :, :, 2 * dy + dx] @ W[2 * a + dy, 2 * b + dx]
    return S


def convT_k4s2_to_blocks(S, W):
    """S -> L in block form.  S [N, HS, WS, CS], W [4, 4, CL, CS] (Keras Conv2DTranspose layout) ->
    blocks [N, HS+1, WS+1, 4, CL]; slots outside the image are zero (the kernel's epilogue masks them)."""
    n, hs, ws, cs = S.shape
    cl = W.shape[2]
    Sp = np.zeros((n, hs + 2, ws + 2, cs), dtype=S.dtype)                 # one zero pixel on every side
    Sp[:, 1:-1, 1:-1] = S
    out = np.zeros((n, hs + 1, ws + 1, 4, cl), dtype=np.result_type(S, W))
    for a in (0, 1):
        for b in (0, 1):
            tap = Sp[:, 1 - a:1 - a + hs + 1, 1 - b:1 - b + ws + 1]       # S[i - a, j - b] for i in [0, HS], j in [0, WS]
            for dy in (0, 1):
                for dx in (0, 1):
                    out[:, :, :, 2 * dy + dx] += tap @ W[2 * a + dy, 2 * b + dx].T
    # slots whose pixel lies outside the image: (i = 0, dy = 0), (i = HS, dy = 1), same for columns
    out[:, 0, :, [0, 1]] = 0
    out[:, hs, :, [2, 3]] = 0
    out[:, :, 0, [0, 2]] = 0
    out[:, :, ws, [1, 3]] = 0
    return out


def gemm_operands_S_to_L(W):
    """the B operand of the S -> L block GEMM: [(a, b, cs), (dy, dx, cl)] = W[2a + dy, 2b + dx, cl, cs]."""
    cl, cs = W.shape[2], W.shape[3]
    B = np.zeros((2, 2, cs, 2, 2, cl), dtype=W.dtype)
    for a in (0, 1):
        for b in (0, 1):
            for dy in (0, 1):
                for dx in (0, 1):
                    B[a, b, :, dy, dx, :] = W[2 * a + dy, 2 * b + dx].T
    return B.reshape(4 * cs, 4 * cl)
