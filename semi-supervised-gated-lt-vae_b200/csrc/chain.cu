// Fused "dense chain" of the ELBO step: everything between the encoder's last convolution and the decoder's first
// transposed convolution is ROW-LOCAL (one image = one row of every operand), so one CTA can carry a group of images
// through all of it without talking to any other CTA:
//
//   forward  (chain_fwd_kernel):   h5 [B,256] -> heads (networks.py:17-18,31-34) -> reparameterised z, gated classifier,
//                                  conditional prior, KL, K-sample log q(y|x) (gated_ccvae.py:167-182,187-218,237-268)
//                                  -> fc1 + ReLU (networks.py:43,52) -> conv1t + ReLU (networks.py:45,54) -> g1 [B,2048]
//   backward (chain_bwd_kernel):   dg1 [B,2048] -> conv1t dgrad -> ReLU mask -> fc1 dgrad -> latent backward (all three
//                                  gradient routes of SURVEY.md 8a) -> heads dgrad -> ReLU mask -> dh5 [B,256],
//                                  plus the bias gradients of conv1t, fc1, the heads and conv5 and the per-CTA partial sums
//                                  of the gate / classifier / prior gradients (reduced by gccvae_gate_bwd).
//
// Before: ten launches on 8..32 CTAs each, every one paying a full launch + pipeline latency on the step's critical
// path (profiles/r01d_graph_timeline.txt: 55 + 107 us of a 690 us supervised step).  Here: one launch each way on
// ceil(B / rows) CTAs of 16 warps.
//
// GEMMs: per CTA the M dimension is the (tiny) image group, so the roles are swapped - the WEIGHT matrix is the
// 16-row A operand of mma.sync.m16n8k16 (bf16 in, fp32 accumulate) and the 8 image slots are the N = 8 columns:
// D[feature, image] = W[feature, k] * X[image, k]^T.  Weights come straight from global memory (L2-resident packed bf16
// operands, 16-byte loads: the k order inside an MMA is permuted so that every thread reads 8 consecutive k), the
// activation tiles live in shared memory.  tcgen05 needs M >= 64 rows of real work per CTA to pay for its pipeline;
// these layers have 8.  The bf16 rounding points are the same as on the tensor-core path (activations stored as bf16
// between layers, fp32 accumulation, fp32 bias / activation / latent math).
#include <cuda_bf16.h>
#include <string.h>

#include "common.cuh"
#include "latent_math.cuh"
#include "tc_common.cuh"

namespace gccvae {
using tc::pack_bf16x2;
using tc::pdl_launch_dependents;
using tc::pdl_wait;

constexpr int CH_THREADS = 512, CH_WARPS = 16, CH_SLOTS = 8;
constexpr int LD_H5 = 256 + 8, LD_G1 = 2048 + 8, LD_64 = 64 + 8, LD_DP = 96 + 8;   // bf16 tile rows, padded by 16 bytes
constexpr int ST_LD = 20;   // staging rows of the K-sample outer product

__device__ __forceinline__ void named_bar(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

__device__ __forceinline__ void mma16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// d[16 features x 8 images] += W[f0 .. f0+16)[k0 .. k0 + 32 NKB) * X[8][k0 ..)^T.
// W: global, row-major [*, ldw] bf16 (packed operand, zero padded); X: shared [8][ldx] bf16.
// Thread (g = lane / 4, tig = lane % 4) loads 16 bytes = k0 + 32 kb + 8 tig .. + 7 of rows f0 + g and f0 + g + 8 and of
// image g: element pairs (0,1) / (2,3) feed the first MMA of the k-block, (4,5) / (6,7) the second.
// Result: d[0] = (f0 + g, image 2 tig), d[1] = (f0 + g, 2 tig + 1), d[2] / d[3] = the same for f0 + g + 8.
template <int NKB>
__device__ __forceinline__ void tile_gemm(float (&d)[4], const __nv_bfloat16* __restrict__ W, int ldw, int f0, int k0,
                                          const __nv_bfloat16* sX, int ldx, int lane) {
  const int g = lane >> 2, tig = lane & 3;
  const uint4* wa = reinterpret_cast<const uint4*>(W + (size_t)(f0 + g) * ldw + k0 + 8 * tig);
  const uint4* wb = reinterpret_cast<const uint4*>(W + (size_t)(f0 + g + 8) * ldw + k0 + 8 * tig);
  const uint4* xs = reinterpret_cast<const uint4*>(sX + g * ldx + k0 + 8 * tig);
  uint4 A0[NKB], A1[NKB];
#pragma unroll
  for (int kb = 0; kb < NKB; ++kb) {
    A0[kb] = __ldg(wa + 4 * kb);
    A1[kb] = __ldg(wb + 4 * kb);
  }
#pragma unroll
  for (int kb = 0; kb < NKB; ++kb) {
    const uint4 x = xs[4 * kb];
    mma16816(d, A0[kb].x, A1[kb].x, A0[kb].y, A1[kb].y, x.x, x.y);
    mma16816(d, A0[kb].z, A1[kb].z, A0[kb].w, A1[kb].w, x.z, x.w);
  }
}

// K-sample loops (100 x 18 softplus / sigmoid per image, the bulk of the kernels' instructions): MUFU-based forms.
// |error| of softplus_k <= ~1e-7 absolute on terms of O(1) that are summed 18 at a time - far inside the bf16 engine's
// budget (the fp32 engine's stand-alone latent kernels keep the library forms).
__device__ __forceinline__ float softplus_k(float x) { return fmaxf(x, 0.0f) + __logf(1.0f + __expf(-fabsf(x))); }
__device__ __forceinline__ float bern_lp_k(float l, bool y1) { return -softplus_k(y1 ? -l : l); }
__device__ __forceinline__ float sigmoid_k(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

// ---------------------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------------------
struct FwdSmem {
  GateSmem g;
  alignas(16) __nv_bfloat16 h5[CH_SLOTS * LD_H5];
  float preP[2][96][CH_SLOTS];   // the two k-halves of the heads GEMM
  float pre[CH_SLOTS][96];       // heads' pre-activations (+ bias): loc at 0..44, scale at 48..92
  float lat[CH_SLOTS][64];       // per image: locc[18] | scc[18] | zc[18] | ymask | lq
  float lse[CH_SLOTS][CH_WARPS][2];
  alignas(16) __nv_bfloat16 z16[CH_SLOTS * LD_64];
  alignas(16) __nv_bfloat16 g0[CH_SLOTS * LD_64];
  alignas(16) __nv_bfloat16 out[CH_SLOTS * LD_G1];
};

template <bool SUP>
__global__ void __launch_bounds__(CH_THREADS, 1) chain_fwd_kernel(const gccvae_chain_fwd_args a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FwdSmem& s = *reinterpret_cast<FwdSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rows = a.rows_per_cta, wpi = CH_WARPS / rows;
  const int row0 = blockIdx.x * rows;
  const int nrows = min(rows, a.batch - row0);
  const int B = a.batch;
  pdl_launch_dependents();
  pdl_wait();
  uint64_t offset = a.offset;
  if (a.step_dev) offset += (uint64_t)(*a.step_dev);
  load_gate_smem(s.g, a.gate_ws);
  // ---- h5 tile (rows beyond the group are zero) --------------------------------------------------------------------
  for (int idx = tid; idx < CH_SLOTS * 32; idx += CH_THREADS) {
    const int r = idx >> 5, c = idx & 31;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r < nrows) v = __ldg(reinterpret_cast<const uint4*>(a.h5) + (size_t)(row0 + r) * 32 + c);
    *reinterpret_cast<uint4*>(&s.h5[r * LD_H5 + 8 * c]) = v;
  }
  for (int idx = tid; idx < CH_SLOTS * LD_64 / 2; idx += CH_THREADS) {
    reinterpret_cast<uint32_t*>(s.z16)[idx] = 0u;
    reinterpret_cast<uint32_t*>(s.g0)[idx] = 0u;
  }
  __syncthreads();
  // ---- heads: [96 x 256] x h5^T, 6 feature tiles x 2 k-halves on 12 warps --------------------------------------------
  if (warp < 12) {
    const int tile = warp % 6, kh = warp / 6;
    float d[4] = {0.f, 0.f, 0.f, 0.f};
    tile_gemm<4>(d, reinterpret_cast<const __nv_bfloat16*>(a.w_heads), 256, 16 * tile, 128 * kh, s.h5, LD_H5, lane);
    const int g = lane >> 2, tig = lane & 3;
    s.preP[kh][16 * tile + g][2 * tig] = d[0];
    s.preP[kh][16 * tile + g][2 * tig + 1] = d[1];
    s.preP[kh][16 * tile + g + 8][2 * tig] = d[2];
    s.preP[kh][16 * tile + g + 8][2 * tig + 1] = d[3];
  }
  __syncthreads();
  for (int idx = tid; idx < CH_SLOTS * 96; idx += CH_THREADS) {
    const int r = idx / 96, f = idx % 96;
    const float v = s.preP[0][f][r] + s.preP[1][f][r] + a.b_heads[f];
    s.pre[r][f] = v;
    if (r < nrows) a.pre[(size_t)(row0 + r) * 96 + f] = v;
  }
  __syncthreads();
  // ---- latent stage: image slot = warp / wpi; its wpi warps share the K importance samples --------------------------
  {
    const int slot = warp / wpi, part = warp % wpi;
    if (slot < nrows) {
      const int b = row0 + slot;
      const float invBg = 1.0f / (float)a.batch_global;
      float* lat = s.lat[slot];
      float* locc = lat;
      float* scc = lat + ZC;
      float* zcs = lat + 2 * ZC;
      const float* pre = s.pre[slot];
      float lq = 0.0f, kl = 0.0f;
      uint32_t ymask = 0u;
      if (part == 0) {
        float kl_part = 0.0f;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int d = lane + 32 * h;
          if (d < Z) {
            float loc, sc;
            head_act(pre[d], pre[48 + d], loc, sc);
            const float e = draw_eps(a.eps, a.seed, offset, b, d);
            const float z = fmaf(sc, e, loc);
            const __nv_bfloat16 zb = __float2bfloat16(z);
            s.z16[slot * LD_64 + d] = zb;
            reinterpret_cast<__nv_bfloat16*>(a.z16)[(size_t)b * 64 + d] = zb;
            a.loc[(size_t)b * Z + d] = loc;
            a.scale[(size_t)b * Z + d] = sc;
            a.z[(size_t)b * Z + d] = z;
            if (d < ZS) {
              kl_part += kl_dim(loc, sc, 0.0f, 1.0f);
            } else {
              locc[d - ZS] = loc;
              scc[d - ZS] = sc;
              zcs[d - ZS] = z;
            }
          } else if (d < 64) {
            reinterpret_cast<__nv_bfloat16*>(a.z16)[(size_t)b * 64 + d] = __float2bfloat16(0.0f);
          }
        }
        __syncwarp();
        float logit = 0.0f;
        bool y1 = false;
        if (lane < Y) {
          float acc = s.g.b[lane];
#pragma unroll
          for (int i = 0; i < ZC; ++i) acc = fmaf(zcs[i], s.g.M[i * Y + lane], acc);
          logit = acc;
          if (SUP) {
            y1 = a.y[(size_t)b * Y + lane] != 0;
          } else {
            float u;
            if (a.U_y != nullptr) {
              u = a.U_y[(size_t)b * Y + lane];
            } else {
              uint32_t r[4];
              philox4x32(a.seed, offset, PH_UY, (uint64_t)b * 5 + (lane >> 2), r);
              u = u32_to_unit(r[lane & 3]);
            }
            y1 = u < sigmoid_f(logit);
          }
          a.logits[(size_t)b * Y + lane] = logit;
          if (a.y_out != nullptr) a.y_out[(size_t)b * Y + lane] = y1 ? 1 : 0;
        }
        ymask = __ballot_sync(0xffffffffu, y1);
        lq = warp_sum(lane < Y ? bern_lp(logit, y1) : 0.0f);
        if (lane < ZC) {
          float mp, spr;
          prior_i(s.g, ymask, lane, mp, spr);
          const float sp = clip_f(softplus_f(spr), 1e-3f, 1e3f);
          kl_part += kl_dim(locc[lane], scc[lane], mp, sp);
        }
        kl = warp_sum(kl_part);
        if (lane == 0) lat[3 * ZC] = __uint_as_float(ymask);
      }
      float lqx = 0.0f, w = 1.0f;
      if (SUP) {
        named_bar(1 + slot, wpi * 32);     // locc / scc / ymask of this image are in shared memory
        ymask = __float_as_uint(lat[3 * ZC]);
        float lc[ZC], sc[ZC];
#pragma unroll
        for (int i = 0; i < ZC; ++i) {
          lc[i] = locc[i];
          sc[i] = scc[i];
        }
        const int Kp = (a.K + wpi - 1) / wpi;
        const int kend = min(a.K, (part + 1) * Kp);
        float m_run = -INFINITY, s_run = 0.0f;
        for (int k = part * Kp + lane; k < kend; k += 32) {
          float e[ZC], zk[ZC], l[Y];
          draw_eps_k(a.eps_k, a.seed, offset, b, k, B, a.K, e);
#pragma unroll
          for (int i = 0; i < ZC; ++i) zk[i] = fmaf(sc[i], e[i], lc[i]);
          sample_logits(s.g, zk, l);
          float acc = 0.0f;
#pragma unroll
          for (int j = 0; j < Y; ++j) acc += bern_lp_k(l[j], (ymask >> j) & 1u);
          const float m_new = fmaxf(m_run, acc);
          s_run = s_run * expf(m_run - m_new) + expf(acc - m_new);
          m_run = m_new;
        }
        const float mx = warp_max(m_run);
        const float pt = (m_run == -INFINITY) ? 0.0f : s_run * expf(m_run - mx);
        const float tot = warp_sum(pt);
        if (lane == 0) {
          s.lse[slot][part][0] = mx;
          s.lse[slot][part][1] = tot;
        }
        named_bar(1 + slot, wpi * 32);
        if (part == 0) {
          float M = -INFINITY;
          for (int q = 0; q < wpi; ++q) M = fmaxf(M, s.lse[slot][q][0]);
          float T = 0.0f;
          for (int q = 0; q < wpi; ++q) {
            const float mq = s.lse[slot][q][0];
            if (mq != -INFINITY) T += s.lse[slot][q][1] * expf(mq - M);
          }
          lqx = M + logf(T) - logf((float)a.K);
          w = expf(lq - lqx);
        }
      }
      if (part == 0 && lane == 0) {
        a.terms[0 * (size_t)B + b] = kl;
        a.terms[1 * (size_t)B + b] = lq;
        a.terms[2 * (size_t)B + b] = lqx;
        a.terms[3 * (size_t)B + b] = w;
        a.terms[4 * (size_t)B + b] = (float)Y * logf(0.5f);
        a.terms[5 * (size_t)B + b] = -w * invBg;
      }
    }
  }
  __syncthreads();
  // ---- fc1: relu(W[45 x 45] z + b), 3 feature tiles ---------------------------------------------------------------
  if (warp < 3) {
    float d[4] = {0.f, 0.f, 0.f, 0.f};
    tile_gemm<2>(d, reinterpret_cast<const __nv_bfloat16*>(a.w_fc1), 64, 16 * warp, 0, s.z16, LD_64, lane);
    const int g = lane >> 2, tig = lane & 3;
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int f = 16 * warp + g + 8 * hh;
      const float bf = f < Z ? a.b_fc1[f] : 0.0f;
      const float v0 = f < Z ? fmaxf(d[2 * hh] + bf, 0.0f) : 0.0f;
      const float v1 = f < Z ? fmaxf(d[2 * hh + 1] + bf, 0.0f) : 0.0f;
      s.g0[(2 * tig) * LD_64 + f] = __float2bfloat16(v0);
      s.g0[(2 * tig + 1) * LD_64 + f] = __float2bfloat16(v1);
    }
  }
  __syncthreads();
  if (tid < CH_SLOTS * 8) {   // fc1 output rows -> global (128 bytes per image)
    const int r = tid >> 3, c = tid & 7;
    if (r < nrows)
      reinterpret_cast<uint4*>(a.g0)[(size_t)(row0 + r) * 8 + c] = *reinterpret_cast<const uint4*>(&s.g0[r * LD_64 + 8 * c]);
  }
  // ---- conv1t: relu(W[2048 x 45] g0 + b[f % 128]), 128 feature tiles, 8 per warp ------------------------------------
  {
    const int g = lane >> 2, tig = lane & 3;
#pragma unroll 2
    for (int t = 0; t < 8; ++t) {
      const int f0 = (warp * 8 + t) * 16;
      float d[4] = {0.f, 0.f, 0.f, 0.f};
      tile_gemm<2>(d, reinterpret_cast<const __nv_bfloat16*>(a.w_conv1t), 64, f0, 0, s.g0, LD_64, lane);
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int f = f0 + g + 8 * hh;
        const float bf = a.b_conv1t[f & 127];
        s.out[(2 * tig) * LD_G1 + f] = __float2bfloat16(fmaxf(d[2 * hh] + bf, 0.0f));
        s.out[(2 * tig + 1) * LD_G1 + f] = __float2bfloat16(fmaxf(d[2 * hh + 1] + bf, 0.0f));
      }
    }
  }
  __syncthreads();
  for (int idx = tid; idx < CH_SLOTS * 256; idx += CH_THREADS) {
    const int r = idx >> 8, c = idx & 255;
    if (r < nrows)
      reinterpret_cast<uint4*>(a.g1)[(size_t)(row0 + r) * 256 + c] = *reinterpret_cast<const uint4*>(&s.out[r * LD_G1 + 8 * c]);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------------------------------
struct BwdImg {
  float locc[ZC], scc[ZC], zc[ZC];
  float ev[48];
  float G[ST_LD], Gz[ST_LD], D[ST_LD];
  float dmp[ZC], dspr[ZC], dlockl[ZC], dsckl[ZC];
  float dzc[ZC], dlx[ZC], dsx[ZC];
  float gl[48], gs[48];       // gradients of the heads' pre-activations (fp32, for the bias gradients)
  float E[NP];
  float sc_g_lqx, sc_lse, loss;
  uint32_t ymask;
};
struct alignas(16) BwdStage {  // one 32-sample chunk of one warp
  float E[32][ST_LD], D[32][ST_LD];
};
struct BwdSmem {
  GateSmem g;
  BwdImg img[CH_SLOTS];
  float pre[CH_SLOTS][96];
  float dz[CH_SLOTS][48];
  alignas(16) __nv_bfloat16 g0[CH_SLOTS * LD_64];     // fc1 output (ReLU mask of dg0)
  alignas(16) __nv_bfloat16 dg0[CH_SLOTS * LD_64];
  alignas(16) __nv_bfloat16 h5[CH_SLOTS * LD_H5];     // conv5 output (ReLU mask of dh5)
  alignas(16) __nv_bfloat16 dp[CH_SLOTS * LD_DP];     // gradient of the heads' pre-activations, bf16 (B operand of the heads dgrad)
  alignas(16) __nv_bfloat16 dh[CH_SLOTS * LD_H5];
  union alignas(16) {
    struct {
      alignas(16) __nv_bfloat16 dg1[CH_SLOTS * LD_G1];
      float part[CH_WARPS][48][CH_SLOTS];   // k-slices of the conv1t dgrad
    } c;
    BwdStage st[CH_WARPS];
  } u;
};

template <bool SUP>
__global__ void __launch_bounds__(CH_THREADS, 1) chain_bwd_kernel(const gccvae_chain_bwd_args a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BwdSmem& s = *reinterpret_cast<BwdSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, tig = lane & 3;
  const int rows = a.rows_per_cta, wpi = CH_WARPS / rows;
  const int row0 = blockIdx.x * rows;
  const int nrows = min(rows, a.batch - row0);
  const int B = a.batch;
  pdl_launch_dependents();
  pdl_wait();
  uint64_t offset = a.offset;
  if (a.step_dev) offset += (uint64_t)(*a.step_dev);
  load_gate_smem(s.g, a.gate_ws);
  // ---- tiles ------------------------------------------------------------------------------------------------------------
  for (int idx = tid; idx < CH_SLOTS * 256; idx += CH_THREADS) {
    const int r = idx >> 8, c = idx & 255;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r < nrows) v = __ldg(reinterpret_cast<const uint4*>(a.dg1) + (size_t)(row0 + r) * 256 + c);
    *reinterpret_cast<uint4*>(&s.u.c.dg1[r * LD_G1 + 8 * c]) = v;
  }
  for (int idx = tid; idx < CH_SLOTS * 32; idx += CH_THREADS) {
    const int r = idx >> 5, c = idx & 31;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r < nrows) v = __ldg(reinterpret_cast<const uint4*>(a.h5) + (size_t)(row0 + r) * 32 + c);
    *reinterpret_cast<uint4*>(&s.h5[r * LD_H5 + 8 * c]) = v;
  }
  if (tid < CH_SLOTS * 8) {
    const int r = tid >> 3, c = tid & 7;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r < nrows) v = __ldg(reinterpret_cast<const uint4*>(a.g0) + (size_t)(row0 + r) * 8 + c);
    *reinterpret_cast<uint4*>(&s.g0[r * LD_64 + 8 * c]) = v;
  }
  for (int idx = tid; idx < CH_SLOTS * 96; idx += CH_THREADS) {
    const int r = idx / 96, f = idx % 96;
    s.pre[r][f] = r < nrows ? a.pre[(size_t)(row0 + r) * 96 + f] : 0.0f;
  }
  for (int idx = tid; idx < CH_SLOTS * LD_64 / 2; idx += CH_THREADS) reinterpret_cast<uint32_t*>(s.dg0)[idx] = 0u;
  for (int idx = tid; idx < CH_SLOTS * LD_DP / 2; idx += CH_THREADS) reinterpret_cast<uint32_t*>(s.dp)[idx] = 0u;
  __syncthreads();
  // ---- conv1t dgrad: [48 x 2048] x dg1^T, every warp takes a k-slice of 128 ----------------------------------------------
#pragma unroll 1
  for (int tile = 0; tile < 3; ++tile) {
    float d[4] = {0.f, 0.f, 0.f, 0.f};
    tile_gemm<4>(d, reinterpret_cast<const __nv_bfloat16*>(a.w_conv1t_t), 2048, 16 * tile, 128 * warp, s.u.c.dg1, LD_G1,
                 lane);
    s.u.c.part[warp][16 * tile + g][2 * tig] = d[0];
    s.u.c.part[warp][16 * tile + g][2 * tig + 1] = d[1];
    s.u.c.part[warp][16 * tile + g + 8][2 * tig] = d[2];
    s.u.c.part[warp][16 * tile + g + 8][2 * tig + 1] = d[3];
  }
  // bias gradient of conv1t: column sums of the dg1 tile over the group's images and the 16 positions
  if (a.db_conv1t != nullptr && tid >= 384) {
    const int c = tid - 384;
    float acc = 0.0f;
    for (int r = 0; r < nrows; ++r)
#pragma unroll
      for (int pos = 0; pos < 16; ++pos) acc += __bfloat162float(s.u.c.dg1[r * LD_G1 + pos * 128 + c]);
    atomicAdd(a.db_conv1t + c, acc);
  }
  __syncthreads();
  if (tid < 48 * CH_SLOTS) {
    const int f = tid >> 3, r = tid & 7;
    float v = 0.0f;
#pragma unroll
    for (int wq = 0; wq < CH_WARPS; ++wq) v += s.u.c.part[wq][f][r];
    const bool on = f < Z && r < nrows && __bfloat162float(s.g0[r * LD_64 + f]) > 0.0f;
    const __nv_bfloat16 vb = __float2bfloat16(on ? v : 0.0f);
    s.dg0[r * LD_64 + f] = vb;
    // bias gradient of fc1 = column sums of the (bf16) tensor the weight gradient reads
    float cs = __bfloat162float(vb);
    cs += __shfl_xor_sync(0xffffffffu, cs, 1);
    cs += __shfl_xor_sync(0xffffffffu, cs, 2);
    cs += __shfl_xor_sync(0xffffffffu, cs, 4);
    if (r == 0 && f < Z && a.db_fc1 != nullptr) atomicAdd(a.db_fc1 + f, cs);
  }
  __syncthreads();
  if (tid < CH_SLOTS * 8) {
    const int r = tid >> 3, c = tid & 7;
    if (r < nrows)
      reinterpret_cast<uint4*>(a.dg0)[(size_t)(row0 + r) * 8 + c] = *reinterpret_cast<const uint4*>(&s.dg0[r * LD_64 + 8 * c]);
  }
  // ---- fc1 dgrad: dz = W[45 x 45] dg0 ---------------------------------------------------------------------------------
  if (warp < 3) {
    float d[4] = {0.f, 0.f, 0.f, 0.f};
    tile_gemm<2>(d, reinterpret_cast<const __nv_bfloat16*>(a.w_fc1_t), 64, 16 * warp, 0, s.dg0, LD_64, lane);
    s.dz[2 * tig][16 * warp + g] = d[0];
    s.dz[2 * tig + 1][16 * warp + g] = d[1];
    s.dz[2 * tig][16 * warp + g + 8] = d[2];
    s.dz[2 * tig + 1][16 * warp + g + 8] = d[3];
  }
  __syncthreads();     // from here on the staging buffers may overwrite the dg1 tile
  // ---- latent backward -------------------------------------------------------------------------------------------------
  {
    const int slot = warp / wpi, part = warp % wpi;
    if (slot < nrows) {
      const int b = row0 + slot;
      BwdImg& im = s.img[slot];
      const float* pre = s.pre[slot];
      const float invBg = 1.0f / (float)a.batch_global;
      float locv[2] = {0.f, 0.f}, scv[2] = {0.f, 0.f}, lpre[2] = {0.f, 0.f}, spre[2] = {0.f, 0.f};
      float cA = 0.0f;
      if (part == 0) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int d = lane + 32 * h;
          if (d < Z) {
            lpre[h] = pre[d];
            spre[h] = pre[48 + d];
            head_act(lpre[h], spre[h], locv[h], scv[h]);
            const float e = draw_eps(a.eps, a.seed, offset, b, d);
            im.ev[d] = e;
            if (d >= ZS) {
              im.locc[d - ZS] = locv[h];
              im.scc[d - ZS] = scv[h];
              im.zc[d - ZS] = fmaf(scv[h], e, locv[h]);
            }
          }
        }
        const bool y1 = (lane < Y) ? (a.y[(size_t)b * Y + lane] != 0) : false;
        const uint32_t ymask = __ballot_sync(0xffffffffu, y1);
        const float kl = a.terms[b];
        const float lq = a.terms[1 * (size_t)B + b];
        const float lqx = a.terms[2 * (size_t)B + b];
        const float w = a.terms[3 * (size_t)B + b];
        const float lpy = a.terms[4 * (size_t)B + b];
        const float lpx = a.log_pxz[b];
        const float A = lpx - kl - lq;
        cA = w * invBg;
        const float g_lqp = SUP ? -A * w * invBg : 0.0f;
        const float g_lqx = SUP ? (A * w - 1.0f) * invBg : 0.0f;
        if (lane == 0) {
          im.loss = SUP ? -(w * A + lpy + lqx) * invBg : -(lpx + lpy - kl - lq) * invBg;
          im.ymask = ymask;
          im.sc_g_lqx = g_lqx;
          im.sc_lse = lqx + (SUP ? logf((float)a.K) : 0.0f);
        }
        for (int t = lane; t < NP; t += 32) im.E[t] = 0.0f;
        if (lane < ST_LD) im.D[lane] = 0.0f;
        __syncwarp();
        if (lane < Y) {
          float acc = s.g.b[lane];
#pragma unroll
          for (int i = 0; i < ZC; ++i) acc = fmaf(im.zc[i], s.g.M[i * Y + lane], acc);
          const float r = (y1 ? 1.0f : 0.0f) - sigmoid_f(acc);
          im.G[lane] = (cA + g_lqp) * r;
          im.Gz[lane] = cA * r;
        }
        if (lane < ZC) {
          float mp, spr;
          prior_i(s.g, ymask, lane, mp, spr);
          const float spv = softplus_f(spr);
          const float sp = clip_f(spv, 1e-3f, 1e3f);
          const float inv = 1.0f / sp;
          const float lqv = im.locc[lane], sq = im.scc[lane];
          const float dq = lqv - mp;
          const float dl = cA * dq * inv * inv;
          im.dlockl[lane] = dl;
          im.dsckl[lane] = cA * (sq * inv * inv - 1.0f / sq);
          im.dmp[lane] = -dl;
          const float dsp = cA * (-(dq * dq) * inv * inv * inv - sq * sq * inv * inv * inv + inv);
          im.dspr[lane] = (spv >= 1e-3f && spv <= 1e3f) ? dsp * sigmoid_f(spr) : 0.0f;
        }
      }
      if (SUP) {
        named_bar(1 + slot, wpi * 32);
        // K-sample responsibilities: D[j] = sum_k dl^k_j, E[i][j] = sum_k eps^k_i dl^k_j over this warp's samples
        BwdStage& st = s.u.st[warp];
        const uint32_t ymask = im.ymask;
        const float g_lqx = im.sc_g_lqx, lse = im.sc_lse;
        const int Kp = (a.K + wpi - 1) / wpi;
        const int kbeg = part * Kp, kend = min(a.K, (part + 1) * Kp);
        const int ib = lane / 3, jb = lane % 3;          // this lane's 2 x 6 block of (i, j) pairs (lanes 0..26)
        float Eacc[2][6];
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
          for (int c = 0; c < 6; ++c) Eacc[q][c] = 0.0f;
        float Dacc = 0.0f;
        for (int k0 = kbeg; k0 < kend; k0 += 32) {
          const int k = k0 + lane;
          if (k < kend) {
            float e[ZC], l[Y];
            {
              float zk[ZC];
              draw_eps_k(a.eps_k, a.seed, offset, b, k, B, a.K, e);
#pragma unroll
              for (int i = 0; i < ZC; ++i) {
                zk[i] = fmaf(im.scc[i], e[i], im.locc[i]);
                st.E[lane][i] = e[i];
              }
              sample_logits(s.g, zk, l);
            }
            float acc = 0.0f;
#pragma unroll
            for (int j = 0; j < Y; ++j) acc += bern_lp_k(l[j], (ymask >> j) & 1u);
            const float rho = expf(acc - lse) * g_lqx;
#pragma unroll
            for (int j = 0; j < Y; ++j)
              st.D[lane][j] = rho * ((((ymask >> j) & 1u) ? 1.0f : 0.0f) - sigmoid_k(l[j]));
          } else {
#pragma unroll
            for (int j = 0; j < Y; ++j) st.D[lane][j] = 0.0f;
#pragma unroll
            for (int i = 0; i < ZC; ++i) st.E[lane][i] = 0.0f;
          }
          __syncwarp();
          if (ib < 9) {
#pragma unroll 8
            for (int t = 0; t < 32; ++t) {
              const float2 e2 = *reinterpret_cast<const float2*>(&st.E[t][2 * ib]);
              const float2 d0 = *reinterpret_cast<const float2*>(&st.D[t][6 * jb]);
              const float2 d1 = *reinterpret_cast<const float2*>(&st.D[t][6 * jb + 2]);
              const float2 d2 = *reinterpret_cast<const float2*>(&st.D[t][6 * jb + 4]);
              Eacc[0][0] = fmaf(e2.x, d0.x, Eacc[0][0]);
              Eacc[0][1] = fmaf(e2.x, d0.y, Eacc[0][1]);
              Eacc[0][2] = fmaf(e2.x, d1.x, Eacc[0][2]);
              Eacc[0][3] = fmaf(e2.x, d1.y, Eacc[0][3]);
              Eacc[0][4] = fmaf(e2.x, d2.x, Eacc[0][4]);
              Eacc[0][5] = fmaf(e2.x, d2.y, Eacc[0][5]);
              Eacc[1][0] = fmaf(e2.y, d0.x, Eacc[1][0]);
              Eacc[1][1] = fmaf(e2.y, d0.y, Eacc[1][1]);
              Eacc[1][2] = fmaf(e2.y, d1.x, Eacc[1][2]);
              Eacc[1][3] = fmaf(e2.y, d1.y, Eacc[1][3]);
              Eacc[1][4] = fmaf(e2.y, d2.x, Eacc[1][4]);
              Eacc[1][5] = fmaf(e2.y, d2.y, Eacc[1][5]);
            }
          }
          if (lane < Y) {
#pragma unroll 8
            for (int t = 0; t < 32; ++t) Dacc += st.D[t][lane];
          }
          __syncwarp();
        }
        if (kbeg < kend) {
          if (ib < 9) {
#pragma unroll
            for (int q = 0; q < 2; ++q)
#pragma unroll
              for (int c = 0; c < 6; ++c) atomicAdd(&im.E[(2 * ib + q) * Y + 6 * jb + c], Eacc[q][c]);
          }
          if (lane < Y) atomicAdd(&im.D[lane], Dacc);
        }
        named_bar(1 + slot, wpi * 32);
      }
      if (part == 0) {
        __syncwarp();
        // back through the gated classifier matrix to z_c / loc_c / scale_c
        if (lane < ZC) {
          float dzv = 0.0f, dlx = 0.0f, dsx = 0.0f;
#pragma unroll
          for (int j = 0; j < Y; ++j) {
            const float m = s.g.M[lane * Y + j];
            dzv = fmaf(m, im.Gz[j], dzv);
            dlx = fmaf(m, im.D[j], dlx);
            dsx = fmaf(m, im.E[lane * Y + j], dsx);
          }
          im.dzc[lane] = dzv;
          im.dlx[lane] = dlx;
          im.dsx[lane] = dsx;
        }
        __syncwarp();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int d = lane + 32 * h;
          if (d < Z) {
            float dzt = s.dz[slot][d];
            float dloc, dsc;
            if (d < ZS) {
              dloc = dzt + cA * locv[h];
              dsc = dzt * im.ev[d] + cA * (scv[h] - 1.0f / scv[h]);
            } else {
              const int i = d - ZS;
              dzt += im.dzc[i];
              dloc = dzt + im.dlockl[i] + im.dlx[i];
              dsc = dzt * im.ev[d] + im.dsckl[i] + im.dsx[i];
            }
            const float spv = softplus_f(spre[h]);
            const float gl = (lpre[h] > 0.0f) ? dloc : 0.0f;
            const float gs = (spv >= 1e-3f && spv <= 1e3f) ? dsc * sigmoid_f(spre[h]) : 0.0f;
            s.dp[slot * LD_DP + d] = __float2bfloat16(gl);
            s.dp[slot * LD_DP + 48 + d] = __float2bfloat16(gs);
            im.gl[d] = gl;
            im.gs[d] = gs;
          }
        }
      }
    }
  }
  __syncthreads();
  // ---- gradient of the heads' pre-activations -> global (operand of the heads' weight gradient) ---------------------------
  if (tid < CH_SLOTS * 12) {
    const int r = tid / 12, c = tid % 12;
    if (r < nrows)
      reinterpret_cast<uint4*>(a.dpre16)[(size_t)(row0 + r) * 12 + c] = *reinterpret_cast<const uint4*>(&s.dp[r * LD_DP + 8 * c]);
  }
  // ---- per-CTA partial sums of the gate / classifier / prior gradients (fixed order over the group's images) ---------------
  {
    float* out = a.partials + (size_t)blockIdx.x * PT_TOTAL;
    for (int t = tid; t < 5 * NP; t += CH_THREADS) {
      const int mtx = t / NP, p = t % NP;
      const int i = p / Y, j = p % Y;
      float v = 0.0f;
      for (int r = 0; r < nrows; ++r) {
        const BwdImg& im = s.img[r];
        const float yj = ((im.ymask >> j) & 1u) ? 1.0f : 0.0f;
        if (mtx == 0) {
          float t0 = im.zc[i] * im.G[j];
          if (SUP) t0 += im.locc[i] * im.D[j] + im.scc[i] * im.E[p];
          v += t0;
        } else if (mtx == 1) {
          v += yj * im.dmp[i];
        } else if (mtx == 2) {
          v += (1.0f - yj) * im.dmp[i];
        } else if (mtx == 3) {
          v += yj * im.dspr[i];
        } else {
          v += (1.0f - yj) * im.dspr[i];
        }
      }
      const int dst = (mtx == 0) ? p : j * ZC + i;    // prior matrices are stored [j][i] like the reference kernels
      out[mtx * NP + dst] = v;
    }
    if (tid < PT_TOTAL - PT_DB) {
      float v = 0.0f;
      if (tid < Y) {
        for (int r = 0; r < nrows; ++r) v += s.img[r].G[tid] + (SUP ? s.img[r].D[tid] : 0.0f);
      } else if (tid == Y) {
        for (int r = 0; r < nrows; ++r) v += s.img[r].loss;
      }
      out[PT_DB + tid] = v;
    }
    if (tid >= 64 && tid < 64 + 2 * 48 && a.db_loc != nullptr) {   // bias gradients of the two heads
      const int q = tid - 64, d = q % 48;
      if (d < Z) {
        float v = 0.0f;
        for (int r = 0; r < nrows; ++r) v += q < 48 ? s.img[r].gl[d] : s.img[r].gs[d];
        atomicAdd((q < 48 ? a.db_loc : a.db_scale) + d, v);
      }
    }
  }
  // ---- heads dgrad: dh5 = W[256 x 96] dpre, one feature tile per warp, ReLU mask of conv5's output -----------------------
  {
    float d[4] = {0.f, 0.f, 0.f, 0.f};
    tile_gemm<3>(d, reinterpret_cast<const __nv_bfloat16*>(a.w_heads_t), 96, 16 * warp, 0, s.dp, LD_DP, lane);
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int f = 16 * warp + g + 8 * hh;
      const bool on0 = __bfloat162float(s.h5[(2 * tig) * LD_H5 + f]) > 0.0f;
      const bool on1 = __bfloat162float(s.h5[(2 * tig + 1) * LD_H5 + f]) > 0.0f;
      const __nv_bfloat16 v0 = __float2bfloat16(on0 ? d[2 * hh] : 0.0f), v1 = __float2bfloat16(on1 ? d[2 * hh + 1] : 0.0f);
      s.dh[(2 * tig) * LD_H5 + f] = v0;
      s.dh[(2 * tig + 1) * LD_H5 + f] = v1;
      float cs = __bfloat162float(v0) + __bfloat162float(v1);     // bias gradient of conv5
      cs += __shfl_xor_sync(0xffffffffu, cs, 1);
      cs += __shfl_xor_sync(0xffffffffu, cs, 2);
      if (tig == 0 && a.db_conv5 != nullptr) atomicAdd(a.db_conv5 + f, cs);
    }
  }
  __syncthreads();
  for (int idx = tid; idx < CH_SLOTS * 32; idx += CH_THREADS) {
    const int r = idx >> 5, c = idx & 31;
    if (r < nrows)
      reinterpret_cast<uint4*>(a.dh5)[(size_t)(row0 + r) * 32 + c] = *reinterpret_cast<const uint4*>(&s.dh[r * LD_H5 + 8 * c]);
  }
}

static int chain_rows(int batch) {
  int r = CH_SLOTS;
  while (r > 1 && (batch + r / 2 - 1) / (r / 2) <= 148) r /= 2;
  return r;
}

template <class Args>
static cudaError_t launch_chain(void (*kernel)(const Args), int grid, size_t smem, cudaStream_t st, const Args& a) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid, 1, 1);
  cfg.blockDim = dim3(CH_THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, a);
}

}  // namespace gccvae

using namespace gccvae;

extern "C" int gccvae_chain_rows(int batch) { return batch > 0 ? chain_rows(batch) : 0; }

extern "C" int gccvae_chain_partials(int batch) {
  if (batch <= 0) return 0;
  const int r = chain_rows(batch);
  return (batch + r - 1) / r;
}

extern "C" int gccvae_chain_fwd(const gccvae_chain_fwd_args* a_in, void* stream) {
  GCC_REQUIRE(a_in, "chain_fwd: null args");
  gccvae_chain_fwd_args a = *a_in;
  GCC_REQUIRE(a.batch > 0 && a.batch_global >= a.batch, "chain_fwd: bad batch %d/%d", a.batch, a.batch_global);
  GCC_REQUIRE(a.h5 && a.w_heads && a.b_heads && a.w_fc1 && a.b_fc1 && a.w_conv1t && a.b_conv1t && a.gate_ws,
              "chain_fwd: null input");
  GCC_REQUIRE(a.pre && a.loc && a.scale && a.z && a.terms && a.logits && a.z16 && a.g0 && a.g1, "chain_fwd: null output");
  if (a.supervised) GCC_REQUIRE(a.y && a.K >= 1, "chain_fwd: supervised needs y and K >= 1");
  GCC_REQUIRE((uintptr_t)a.eps_k % 8 == 0, "chain_fwd: eps_k must be 8-byte aligned");
  GCC_REQUIRE(((uintptr_t)a.h5 | (uintptr_t)a.w_heads | (uintptr_t)a.w_fc1 | (uintptr_t)a.w_conv1t | (uintptr_t)a.z16 |
               (uintptr_t)a.g0 | (uintptr_t)a.g1) % 16 == 0,
              "chain_fwd: bf16 operands must be 16-byte aligned");
  a.rows_per_cta = chain_rows(a.batch);
  const int grid = (a.batch + a.rows_per_cta - 1) / a.rows_per_cta;
  const size_t smem = sizeof(FwdSmem);
  static bool attr_done = false;
  if (!attr_done) {
    GCC_CUDA(cudaFuncSetAttribute(chain_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GCC_CUDA(cudaFuncSetAttribute(chain_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  if (a.supervised)
    GCC_CUDA(launch_chain(chain_fwd_kernel<true>, grid, smem, (cudaStream_t)stream, a));
  else
    GCC_CUDA(launch_chain(chain_fwd_kernel<false>, grid, smem, (cudaStream_t)stream, a));
  GCC_CHECK_LAUNCH("chain_fwd");
  return GCCVAE_OK;
}

extern "C" int gccvae_chain_bwd(const gccvae_chain_bwd_args* a_in, void* stream) {
  GCC_REQUIRE(a_in, "chain_bwd: null args");
  gccvae_chain_bwd_args a = *a_in;
  GCC_REQUIRE(a.batch > 0 && a.batch_global >= a.batch, "chain_bwd: bad batch");
  GCC_REQUIRE(a.dg1 && a.g0 && a.h5 && a.pre && a.y && a.terms && a.log_pxz && a.gate_ws && a.w_conv1t_t && a.w_fc1_t &&
                  a.w_heads_t,
              "chain_bwd: null input");
  GCC_REQUIRE(a.dg0 && a.dpre16 && a.dh5 && a.partials, "chain_bwd: null output");
  GCC_REQUIRE((a.db_loc == nullptr) == (a.db_scale == nullptr), "chain_bwd: db_loc and db_scale go together");
  GCC_REQUIRE((uintptr_t)a.eps_k % 8 == 0, "chain_bwd: eps_k must be 8-byte aligned");
  GCC_REQUIRE(((uintptr_t)a.dg1 | (uintptr_t)a.g0 | (uintptr_t)a.h5 | (uintptr_t)a.w_conv1t_t | (uintptr_t)a.w_fc1_t |
               (uintptr_t)a.w_heads_t | (uintptr_t)a.dg0 | (uintptr_t)a.dpre16 | (uintptr_t)a.dh5) % 16 == 0,
              "chain_bwd: bf16 operands must be 16-byte aligned");
  a.rows_per_cta = chain_rows(a.batch);
  const int grid = (a.batch + a.rows_per_cta - 1) / a.rows_per_cta;
  GCC_REQUIRE(a.n_partials == grid, "chain_bwd: n_partials must be %d", grid);
  const size_t smem = sizeof(BwdSmem);
  static bool attr_done = false;
  if (!attr_done) {
    GCC_CUDA(cudaFuncSetAttribute(chain_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GCC_CUDA(cudaFuncSetAttribute(chain_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  if (a.supervised)
    GCC_CUDA(launch_chain(chain_bwd_kernel<true>, grid, smem, (cudaStream_t)stream, a));
  else
    GCC_CUDA(launch_chain(chain_bwd_kernel<false>, grid, smem, (cudaStream_t)stream, a));
  GCC_CHECK_LAUNCH("chain_bwd");
  return GCCVAE_OK;
}
