// Microbenchmark 3: TMA 4-D box throughput per SM as a function of the inner row size (channels) and element stride.
// Tensor [N, H, W, C] bf16; each load is a box (C, 32, 4, 1) = 128 rows of C*2 bytes; `ctas_per_sm` CTAs per SM,
// each keeping `stages` loads in flight.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "tc_common.cuh"
using namespace gccvae::tc;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct P { CUtensorMap tm; int box_bytes, iters, stages, H, W, N, es; long long* out; };

__global__ void __launch_bounds__(64) k(const __grid_constant__ P p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  __shared__ uint64_t bar[8];
  if (threadIdx.x == 0) { for (int s = 0; s < p.stages; ++s) mbar_init(&bar[s], 1); fence_mbar_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int stride = (p.box_bytes + 1023) & ~1023;
    const int tiles_h = p.H / (4 * p.es);
    const long long t0 = clock64();
    uint32_t ph = 0;
    int stage = 0;
    for (int i = 0; i < p.iters + p.stages; ++i) {
      if (i >= p.stages) { mbar_wait(&bar[stage], ph); }
      if (i < p.iters) {
        const int tile = (blockIdx.x * p.iters + i);
        const int n = (tile / tiles_h) % p.N, h0 = (tile % tiles_h) * 4 * p.es;
        mbar_expect_tx(&bar[stage], (uint32_t)p.box_bytes);
        tma_load_4d(smem + stage * stride, &p.tm, &bar[stage], 0, 0, h0, n);
      }
      if (++stage == p.stages) { stage = 0; if (i >= p.stages) ph ^= 1; }
    }
    const long long t1 = clock64();
    if (blockIdx.x == 0) p.out[0] = t1 - t0;
  }
}

int main() {
  EncodeTiledFn enc = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q);
  long long* d;
  cudaMalloc(&d, 8);
  const int N = 1024, H = 64, W = 64;
  void* buf;
  cudaMalloc(&buf, (size_t)N * H * W * 64 * 2);
  cudaMemset(buf, 0, (size_t)N * H * W * 64 * 2);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 70 * 1024);
  for (int es : {1, 2})
    for (int C : {16, 32, 64})
      for (int per_sm : {1, 3})
        for (int stages : {2, 4}) {
          P p;
          cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
          cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
          cuuint32_t box[4] = {(cuuint32_t)C, (cuuint32_t)(32 * es), (cuuint32_t)(4 * es), 1};
          cuuint32_t estr[4] = {1, (cuuint32_t)es, (cuuint32_t)es, 1};
          CUtensorMapSwizzle sw = C == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : C == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
          CUresult r = enc(&p.tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, buf, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
          if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
          p.box_bytes = 128 * C * 2; p.iters = 400; p.stages = stages; p.H = H; p.W = W; p.N = N; p.es = es; p.out = d;
          const size_t smem = (size_t)stages * ((p.box_bytes + 1023) & ~1023) + 1024;
          k<<<148 * per_sm, 64, smem>>>(p);
          cudaError_t e = cudaDeviceSynchronize();
          long long c = 0;
          cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
          const double clk_per_box = (double)c / p.iters / per_sm;   // SM-level
          printf("es %d C %2d (%3d B rows) ctas/SM %d stages %d : %7.1f clk/box/SM  %5.2f clk/row  %6.1f B/clk/SM  (%s)\n", es, C,
                 C * 2, per_sm, stages, clk_per_box, clk_per_box / 128, p.box_bytes / clk_per_box, cudaGetErrorString(e));
        }
  return 0;
}
