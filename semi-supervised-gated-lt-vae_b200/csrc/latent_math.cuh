// Device helpers shared by the latent kernels (latent.cu) and the fused dense-chain kernels (chain.cu): layout of the
// per-step gate workspace, the shared-memory image of the gated matrices, the noise draws and the closed forms of
// the reference's TFP distributions (gated_ccvae.py:62-64,90-93,102-111,167-182; networks.py:17-18,33-34,72-74,
// 83-86,104-106,118-127; utils.py:108-119).
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace gccvae {

constexpr int Z = GCCVAE_Z, ZS = GCCVAE_ZS, ZC = GCCVAE_ZC, Y = GCCVAE_Y;
constexpr int NP = ZC * Y;  // 324 (i,j) pairs
constexpr int WARPS = 8;
constexpr int PAIR_SLOTS = (NP + 31) / 32;  // 11
constexpr int MT_LD = 20;                   // padded row of the transposed gated classifier matrix

// gate_ws layout (floats)
constexpr int GW_C = 0, GW_M = NP, GW_B = 2 * NP, GW_PLT = 2 * NP + 32, GW_PLF = 3 * NP + 32, GW_PST = 4 * NP + 32,
              GW_PSF = 5 * NP + 32, GW_DCDM = 6 * NP + 32, GW_TOTAL = 7 * NP + 32;
// partial layout (floats)
constexpr int PT_DB = 5 * NP,  // dM | dPlt | dPlf | dPst | dPsf (5 x 324) | db[18] | loss | pad
              PT_LOSS = 5 * NP + 18, PT_TOTAL = GCCVAE_LATENT_PARTIAL_FLOATS;

static_assert(GCCVAE_GATE_WS_FLOATS >= GW_TOTAL, "gate workspace too small");

// ---------------------------------------------------------------------------------------------
// shared-memory image of the gated matrices
// ---------------------------------------------------------------------------------------------
struct GateSmem {
  float MT[Y * MT_LD];  // MT[j][i] = c[i,j] * Wcls[i,j], rows padded to 20 floats (LDS.128)
  float M[NP];          // M[i][j]
  float b[32];
  float Plt[NP], Plf[NP], Pst[NP], Psf[NP];  // [j][i]
};

__device__ __forceinline__ void load_gate_smem(GateSmem& g, const float* __restrict__ ws) {
  for (int t = threadIdx.x; t < NP; t += blockDim.x) {
    const float m = ws[GW_M + t];
    g.M[t] = m;
    g.MT[(t % Y) * MT_LD + (t / Y)] = m;
    g.Plt[t] = ws[GW_PLT + t];
    g.Plf[t] = ws[GW_PLF + t];
    g.Pst[t] = ws[GW_PST + t];
    g.Psf[t] = ws[GW_PSF + t];
  }
  for (int t = threadIdx.x; t < Y * (MT_LD - ZC); t += blockDim.x) g.MT[(t / 2) * MT_LD + ZC + (t % 2)] = 0.0f;
  if (threadIdx.x < 32) g.b[threadIdx.x] = ws[GW_B + threadIdx.x];
}

// one standard normal for (image b, dim d) — fixed tensor or Philox
__device__ __forceinline__ float draw_eps(const float* __restrict__ eps, uint64_t seed, uint64_t offset, int b, int d) {
  if (eps != nullptr) return eps[(size_t)b * Z + d];
  float n[4];
  philox_normal4(seed, offset, PH_EPS, (uint64_t)b * 12 + (d >> 2), n);
  return n[d & 3];
}

// the 18 classify-dim normals of importance sample k of image b
__device__ __forceinline__ void draw_eps_k(const float* __restrict__ eps_k, uint64_t seed, uint64_t offset, int b,
                                           int k, int B, int K, float (&e)[ZC]) {
  if (eps_k != nullptr) {
    const float* src = eps_k + ((size_t)k * B + b) * ZC;  // 72-byte rows, 8-byte aligned
    const float2* s2 = reinterpret_cast<const float2*>(src);
#pragma unroll
    for (int t = 0; t < ZC / 2; ++t) {
      float2 v = __ldg(s2 + t);
      e[2 * t] = v.x;
      e[2 * t + 1] = v.y;
    }
  } else {
    const uint64_t base = ((uint64_t)b * K + k) * 5;
#pragma unroll
    for (int q = 0; q < 5; ++q) {
      float n[4];
      philox_normal4(seed, offset, PH_EPS_K, base + q, n);
#pragma unroll
      for (int t = 0; t < 4; ++t)
        if (q * 4 + t < ZC) e[q * 4 + t] = n[t];
    }
  }
}

// logits of one sample: l[j] = b[j] + sum_i zk[i] * M[i][j], via the transposed padded copy
__device__ __forceinline__ void sample_logits(const GateSmem& g, const float (&zk)[ZC], float (&l)[Y]) {
#pragma unroll
  for (int j = 0; j < Y; ++j) {
    const float4* row = reinterpret_cast<const float4*>(&g.MT[j * MT_LD]);
    float acc = g.b[j];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 m = row[q];
      acc = fmaf(zk[4 * q + 0], m.x, acc);
      acc = fmaf(zk[4 * q + 1], m.y, acc);
      acc = fmaf(zk[4 * q + 2], m.z, acc);
      acc = fmaf(zk[4 * q + 3], m.w, acc);
    }
    const float4 m = row[4];
    acc = fmaf(zk[16], m.x, acc);
    acc = fmaf(zk[17], m.y, acc);
    l[j] = acc;
  }
}

// Bernoulli(logits=l).log_prob(y) = -softplus((1-2y) l)
__device__ __forceinline__ float bern_lp(float l, bool y1) { return -softplus_f(y1 ? -l : l); }

// posterior heads' activations (networks.py:17-18,33-34)
__device__ __forceinline__ void head_act(float lp, float sp, float& loc, float& sc) {
  loc = fmaxf(lp, 0.0f);
  sc = clip_f(softplus_f(sp), 1e-3f, 1e3f);
}

// conditional prior of classify dim i given labels (networks.py:118-127)
__device__ __forceinline__ void prior_i(const GateSmem& g, uint32_t ymask, int i, float& mp, float& spr) {
  float a = 0.0f, s = 0.0f;
#pragma unroll
  for (int j = 0; j < Y; ++j) {
    const bool y1 = (ymask >> j) & 1u;
    a += y1 ? g.Plt[j * ZC + i] : g.Plf[j * ZC + i];
    s += y1 ? g.Pst[j * ZC + i] : g.Psf[j * ZC + i];
  }
  mp = a;
  spr = s;
}

// TFP _kl_normal_normal
__device__ __forceinline__ float kl_dim(float lq, float sq, float lp, float sp) {
  const float dls = logf(sq) - logf(sp);
  const float d = lq / sp - lp / sp;
  return 0.5f * d * d + 0.5f * expm1f(2.0f * dls) - dls;
}

}  // namespace gccvae
