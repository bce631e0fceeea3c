// fp32 (CUDA-core) implicit-GEMM kernels for the layer relation of gccvae.h:
//   LS  : S = act(gather(L) W + b)          Conv2D / Dense forward, Conv2DTranspose dgrad
//   SL  : L = act(scatter(S) W^T + b)       Conv2DTranspose forward, Conv2D / Dense dgrad
//   WG  : dW = gather(L)^T S                wgrad of every layer kind (deterministic split-K)
// Reference layers: networks.py:11-18,21-36 (encoder), :43-49,52-58 (decoder).
// This is the exact-arithmetic path (fp32 in, fp32 FMA accumulate) that meets the 1e-5 parity bar
// and serves as the on-GPU reference for the bf16 tcgen05 path in conv_tc.cu.
//
// One tiled SGEMM core (64x64x16 or 128x32x16 CTA tile, 4x4 register micro-tile, register-staged
// prefetch of the next k-slab) parameterised by a "problem" policy that maps (m,k)/(k,n) to memory.
#include "common.cuh"

namespace gccvae {

constexpr int BK = 16;

// ---- problem policies --------------------------------------------------------------------------
struct EpiArgs {
  const float* bias;
  const float* mask;
  int act;  // gccvae_act | GCCVAE_ACT_ACCUMULATE
};

__device__ __forceinline__ float apply_epi(float v, const EpiArgs& e, int bias_idx, size_t out_idx, const float* dst) {
  if (e.bias) v += e.bias[bias_idx];
  if (e.act & GCCVAE_ACT_ACCUMULATE) v += dst[out_idx];
  const int act = e.act & 0xff;
  if (act == GCCVAE_ACT_RELU) v = fmaxf(v, 0.0f);
  else if (act == GCCVAE_ACT_SIGMOID) v = sigmoid_f(v);
  if (e.mask && !(e.mask[out_idx] > 0.0f)) v = 0.0f;
  return v;
}

// L -> S
struct ProbLS {
  gccvae_geom g;
  const float* L;
  const float* W;
  float* S;
  EpiArgs epi;
  static constexpr bool A_KFAST = true, B_KFAST = false;
  __device__ int M() const { return g.batch * g.HS * g.WS; }
  __device__ int N() const { return g.CS; }
  __device__ int K() const { return g.KH * g.KW * g.CL; }
  __device__ int splitk_begin() const { return 0; }
  struct Row { int base, ih0, iw0; };
  struct Col { int kh, kw, cl; };
  __device__ Row rowA(int m) const {
    const int ow = m % g.WS, t = m / g.WS, oh = t % g.HS, n = t / g.HS;
    return {n * g.HL * g.WL * g.CL, g.stride * oh - g.pad, g.stride * ow - g.pad};
  }
  __device__ Col colA(int k) const {
    const int cl = k % g.CL, t = k / g.CL;
    return {t / g.KW, t % g.KW, cl};
  }
  __device__ float fetchA(const Row& r, const Col& c) const {
    const int ih = r.ih0 + c.kh, iw = r.iw0 + c.kw;
    if ((unsigned)ih >= (unsigned)g.HL || (unsigned)iw >= (unsigned)g.WL) return 0.0f;
    return __ldg(L + r.base + (ih * g.WL + iw) * g.CL + c.cl);
  }
  __device__ float fetchB(int k, int n) const { return __ldg(W + (size_t)k * g.CS + n); }
  __device__ void store(int m, int n, float v) const {
    const size_t o = (size_t)m * g.CS + n;
    S[o] = apply_epi(v, epi, n, o, S);
  }
};

// S -> L, 4x4 kernel, stride 2, pad 1: one launch per output-parity phase (blockIdx.z)
struct ProbSL2 {
  gccvae_geom g;
  const float* S;
  const float* W;
  float* L;
  EpiArgs epi;
  static constexpr bool A_KFAST = true, B_KFAST = true;
  __device__ int M() const { return g.batch * g.HS * g.WS; }
  __device__ int N() const { return g.CL; }
  __device__ int K() const { return 4 * g.CS; }
  struct Row { int base, a, b; };
  struct Col { int th, tw, cs; };
  __device__ Row rowA(int m) const {
    const int b = m % g.WS, t = m / g.WS, a = t % g.HS, n = t / g.HS;
    return {n * g.HS * g.WS * g.CS, a, b};
  }
  __device__ Col colA(int k) const {
    const int cs = k % g.CS, t = k / g.CS;
    return {t >> 1, t & 1, cs};
  }
  // output row ih = 2a+ph uses kh = (ph+1)%2 + 2*th and input row oh = a + ph - th
  __device__ float fetchA(const Row& r, const Col& c) const {
    const int ph = blockIdx.z >> 1, pw = blockIdx.z & 1;
    const int oh = r.a + ph - c.th, ow = r.b + pw - c.tw;
    if ((unsigned)oh >= (unsigned)g.HS || (unsigned)ow >= (unsigned)g.WS) return 0.0f;
    return __ldg(S + r.base + (oh * g.WS + ow) * g.CS + c.cs);
  }
  __device__ float fetchB(int k, int n) const {
    const int ph = blockIdx.z >> 1, pw = blockIdx.z & 1;
    const int cs = k % g.CS, t = k / g.CS;
    const int kh = ((ph + 1) & 1) + 2 * (t >> 1), kw = ((pw + 1) & 1) + 2 * (t & 1);
    return __ldg(W + ((size_t)(kh * 4 + kw) * g.CL + n) * g.CS + cs);
  }
  __device__ void store(int m, int n, float v) const {
    const int ph = blockIdx.z >> 1, pw = blockIdx.z & 1;
    const int b = m % g.WS, t = m / g.WS, a = t % g.HS, img = t / g.HS;
    const size_t o = (((size_t)img * g.HL + (2 * a + ph)) * g.WL + (2 * b + pw)) * g.CL + n;
    L[o] = apply_epi(v, epi, n, o, L);
  }
};

// S -> L with HS = WS = 1 (Dense, conv5 dgrad, conv1t forward): L[n, (kh,kw,cl)] = S[n,:] . W[(kh,kw,cl), :]
struct ProbSLDense {
  gccvae_geom g;
  const float* S;
  const float* W;
  float* L;
  EpiArgs epi;
  static constexpr bool A_KFAST = true, B_KFAST = true;
  __device__ int M() const { return g.batch; }
  __device__ int N() const { return g.KH * g.KW * g.CL; }
  __device__ int K() const { return g.CS; }
  struct Row { int base; };
  struct Col { int k; };
  __device__ Row rowA(int m) const { return {m * g.CS}; }
  __device__ Col colA(int k) const { return {k}; }
  __device__ float fetchA(const Row& r, const Col& c) const { return __ldg(S + r.base + c.k); }
  __device__ float fetchB(int k, int n) const { return __ldg(W + (size_t)n * g.CS + k); }
  __device__ void store(int m, int n, float v) const {
    const size_t o = (size_t)m * N() + n;
    L[o] = apply_epi(v, epi, n % g.CL, o, L);
  }
};

// weight gradient: dW[(kh,kw,cl), cs] = sum_pix gather(L)[pix,(kh,kw,cl)] * S[pix, cs]; the GEMM's
// reduction axis is the pixel axis, split over blockIdx.z; partials go to the workspace.
struct ProbWG {
  gccvae_geom g;
  const float* L;
  const float* S;
  float* part;
  int pix_per_split;
  static constexpr bool A_KFAST = false, B_KFAST = false;
  __device__ int M() const { return g.KH * g.KW * g.CL; }
  __device__ int N() const { return g.CS; }
  __device__ int K() const {  // this split's end
    const int tot = g.batch * g.HS * g.WS;
    const int e = (blockIdx.z + 1) * pix_per_split;
    return e < tot ? e : tot;
  }
  __device__ int k_begin() const { return blockIdx.z * pix_per_split; }
  struct Row { int kh, kw, cl; };          // "row" of A^T = weight row m
  struct Col { int base, ih0, iw0; };      // "col" = pixel
  __device__ Row rowA(int m) const {
    const int cl = m % g.CL, t = m / g.CL;
    return {t / g.KW, t % g.KW, cl};
  }
  __device__ Col colA(int k) const {
    const int ow = k % g.WS, t = k / g.WS, oh = t % g.HS, n = t / g.HS;
    return {n * g.HL * g.WL * g.CL, g.stride * oh - g.pad, g.stride * ow - g.pad};
  }
  __device__ float fetchA(const Row& r, const Col& c) const {
    const int ih = c.ih0 + r.kh, iw = c.iw0 + r.kw;
    if ((unsigned)ih >= (unsigned)g.HL || (unsigned)iw >= (unsigned)g.WL) return 0.0f;
    return __ldg(L + c.base + (ih * g.WL + iw) * g.CL + r.cl);
  }
  __device__ float fetchB(int k, int n) const { return __ldg(S + (size_t)k * g.CS + n); }
  __device__ void store(int m, int n, float v) const {
    part[((size_t)blockIdx.z * M() + m) * g.CS + n] = v;
  }
};

template <class P>
struct HasKBegin { static constexpr bool value = false; };
template <>
struct HasKBegin<ProbWG> { static constexpr bool value = true; };

// ---- tiled SGEMM core ------------------------------------------------------------------------------
template <int BM, int BN, class P>
__global__ void __launch_bounds__(256) igemm_f32_kernel(const P p) {
  constexpr int TM = 4, TN = 4;
  constexpr int TX = BN / TN;              // threads along n
  constexpr int A_PER = BM * BK / 256;     // A elements per thread per slab
  constexpr int B_PER = BN * BK / 256;
  constexpr int LDA = BM + 4, LDB = BN + 4;
  static_assert((BM / TM) * TX == 256, "256 threads");
  __shared__ __align__(16) float As[BK][LDA];
  __shared__ __align__(16) float Bs[BK][LDB];

  const int tid = threadIdx.x;
  const int tx = tid % TX, ty = tid / TX;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int M = p.M(), N = p.N(), Kend = p.K();
  int kbeg = 0;
  if constexpr (HasKBegin<P>::value) kbeg = p.k_begin();

  // fixed per-thread A rows (KFAST) or A cols (MFAST)
  typename P::Row rows[A_PER];
  int a_m[A_PER], a_k[A_PER];
#pragma unroll
  for (int r = 0; r < A_PER; ++r) {
    const int idx = tid + 256 * r;
    if (P::A_KFAST) { a_k[r] = idx % BK; a_m[r] = idx / BK; }
    else            { a_m[r] = idx % BM; a_k[r] = idx / BM; }
    const int m = m0 + a_m[r];
    rows[r] = p.rowA(m < M ? m : 0);
  }
  int b_n[B_PER], b_k[B_PER];
#pragma unroll
  for (int r = 0; r < B_PER; ++r) {
    const int idx = tid + 256 * r;
    if (P::B_KFAST) { b_k[r] = idx % BK; b_n[r] = idx / BK; }
    else            { b_n[r] = idx % BN; b_k[r] = idx / BN; }
  }

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;

  float ra[A_PER], rb[B_PER];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int r = 0; r < A_PER; ++r) {
      const int k = k0 + a_k[r], m = m0 + a_m[r];
      ra[r] = (k < Kend && m < M) ? p.fetchA(rows[r], p.colA(k)) : 0.0f;
    }
#pragma unroll
    for (int r = 0; r < B_PER; ++r) {
      const int k = k0 + b_k[r], n = n0 + b_n[r];
      rb[r] = (k < Kend && n < N) ? p.fetchB(k, n) : 0.0f;
    }
  };

  fetch(kbeg);
  for (int k0 = kbeg; k0 < Kend; k0 += BK) {
#pragma unroll
    for (int r = 0; r < A_PER; ++r) As[a_k[r]][a_m[r]] = ra[r];
#pragma unroll
    for (int r = 0; r < B_PER; ++r) Bs[b_k[r]][b_n[r]] = rb[r];
    __syncthreads();
    if (k0 + BK < Kend) fetch(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * TM]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * TN]);
      const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * TM + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n < N) p.store(m, n, acc[i][j]);
    }
  }
}

// sum the split-K partials in fixed order
__global__ void wg_reduce_kernel(const float* __restrict__ part, int splits, int n, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float acc = 0.0f;
  for (int s = 0; s < splits; ++s) acc += part[(size_t)s * n + i];
  out[i] = acc;
}

static int check_geom(const gccvae_geom* g, const char* who) {
  GCC_REQUIRE(g, "%s: null geometry", who);
  GCC_REQUIRE(g->batch > 0 && g->HL > 0 && g->WL > 0 && g->CL > 0 && g->HS > 0 && g->WS > 0 && g->CS > 0 &&
                  g->KH > 0 && g->KW > 0 && g->stride > 0 && g->pad >= 0,
              "%s: non-positive dimension", who);
  GCC_REQUIRE((g->HL + 2 * g->pad - g->KH) / g->stride + 1 == g->HS && (g->WL + 2 * g->pad - g->KW) / g->stride + 1 == g->WS,
              "%s: inconsistent geometry L %dx%d k%d s%d p%d -> S %dx%d", who, g->HL, g->WL, g->KH, g->stride, g->pad,
              g->HS, g->WS);
  GCC_REQUIRE((long long)g->batch * g->HL * g->WL * g->CL < (1LL << 31) &&
                  (long long)g->batch * g->HS * g->WS * g->CS < (1LL << 31),
              "%s: tensor exceeds 2^31 elements (shard the batch)", who);
  return GCCVAE_OK;
}

template <class P>
static int launch_igemm(const P& p, int M, int N, int gz, cudaStream_t st, const char* name) {
  if (N <= 32) {
    dim3 grid((M + 127) / 128, (N + 31) / 32, gz);
    igemm_f32_kernel<128, 32, P><<<grid, 256, 0, st>>>(p);
  } else {
    dim3 grid((M + 63) / 64, (N + 63) / 64, gz);
    igemm_f32_kernel<64, 64, P><<<grid, 256, 0, st>>>(p);
  }
  GCC_CHECK_LAUNCH(name);
  return GCCVAE_OK;
}

static int wg_splits(const gccvae_geom* g) {
  const int M = g->KH * g->KW * g->CL, N = g->CS;
  const long long pix = (long long)g->batch * g->HS * g->WS;
  const int tiles = (N <= 32) ? ((M + 127) / 128) * ((N + 31) / 32) : ((M + 63) / 64) * ((N + 63) / 64);
  long long want = (148LL * 4 + tiles - 1) / tiles;
  const long long maxs = (pix + 63) / 64;
  if (want > maxs) want = maxs;
  if (want < 1) want = 1;
  if (want > 512) want = 512;
  return (int)want;
}

}  // namespace gccvae

using namespace gccvae;

extern "C" int gccvae_ls_f32(const gccvae_geom* g, const float* L, const float* W, const float* bias, int act,
                             const float* mask, float* S, void* stream) {
  if (int rc = check_geom(g, "ls_f32")) return rc;
  GCC_REQUIRE(L && W && S, "ls_f32: null pointer");
  ProbLS p{*g, L, W, S, {bias, mask, act}};
  return launch_igemm(p, g->batch * g->HS * g->WS, g->CS, 1, (cudaStream_t)stream, "ls_f32");
}

extern "C" int gccvae_sl_f32(const gccvae_geom* g, const float* S, const float* W, const float* bias, int act,
                             const float* mask, float* L, void* stream) {
  if (int rc = check_geom(g, "sl_f32")) return rc;
  GCC_REQUIRE(S && W && L, "sl_f32: null pointer");
  if (g->HS == 1 && g->WS == 1 && g->pad == 0 && g->stride == 1) {
    ProbSLDense p{*g, S, W, L, {bias, mask, act}};
    return launch_igemm(p, g->batch, g->KH * g->KW * g->CL, 1, (cudaStream_t)stream, "sl_f32(dense)");
  }
  GCC_REQUIRE(g->KH == 4 && g->KW == 4 && g->stride == 2 && g->pad == 1,
              "sl_f32: only k4/s2/p1 or 1x1-spatial S are supported (got k%d s%d p%d, S %dx%d)", g->KH, g->stride,
              g->pad, g->HS, g->WS);
  ProbSL2 p{*g, S, W, L, {bias, mask, act}};
  return launch_igemm(p, g->batch * g->HS * g->WS, g->CL, 4, (cudaStream_t)stream, "sl_f32(s2)");
}

extern "C" size_t gccvae_wg_f32_workspace_bytes(const gccvae_geom* g) {
  if (!g) return 0;
  return (size_t)wg_splits(g) * g->KH * g->KW * g->CL * g->CS * sizeof(float);
}

extern "C" int gccvae_wg_f32(const gccvae_geom* g, const float* L, const float* S, float* dW, void* workspace,
                             size_t workspace_bytes, void* stream) {
  if (int rc = check_geom(g, "wg_f32")) return rc;
  GCC_REQUIRE(L && S && dW && workspace, "wg_f32: null pointer");
  const size_t need = gccvae_wg_f32_workspace_bytes(g);
  if (workspace_bytes < need) {
    set_error("wg_f32: workspace %zu < %zu bytes", workspace_bytes, need);
    return GCCVAE_ENOMEM;
  }
  const int splits = wg_splits(g);
  const int pix = g->batch * g->HS * g->WS;
  const int M = g->KH * g->KW * g->CL, N = g->CS;
  ProbWG p{*g, L, S, (float*)workspace, (pix + splits - 1) / splits};
  if (int rc = launch_igemm(p, M, N, splits, (cudaStream_t)stream, "wg_f32")) return rc;
  const int n = M * N;
  wg_reduce_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const float*)workspace, splits, n, dW);
  GCC_CHECK_LAUNCH("wg_reduce");
  return GCCVAE_OK;
}
