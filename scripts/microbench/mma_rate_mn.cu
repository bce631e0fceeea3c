// Microbenchmark: cycles per tcgen05.mma (M = 128, K = 16, bf16) with MN-major operands (the weight-gradient kernels'
// form: A = rows of 128 bytes / 64 channels, SWIZZLE_128B, two slabs; B = rows of 2 N bytes) as a function of N and of
// the ROW OFFSET of the A start address (0 = aligned to the 8-row swizzle atom; 1, 17, 18 = the tap offsets of the halo
// form), from a fully unrolled issue loop.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I ../../semi-supervised-gated-lt-vae_b200/csrc
//        -o mma_rate_mn mma_rate_mn.cu
#include <cstdio>
#include <cstdlib>
#include "tc_common.cuh"
using namespace gccvae::tc;

struct P { int N, off_rows, kmajor; long long* out; };
constexpr int A_SLAB = 40 * 1024;

__global__ void __launch_bounds__(128) k(P p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 120 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc(&slot, 256);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    const uint32_t rowB = (uint32_t)p.N * 2u;
    uint64_t ad, bd;
    uint32_t idesc, astep, bstep;
    if (p.kmajor) {   // reference point: K-major operands, 64-byte rows (K = 32 per row), SW64
      idesc = instr_desc_bf16(128, p.N, 0, 0);
      ad = smem_desc(smem_u32(smem) + (uint32_t)p.off_rows * 64u, 16, 8 * 64, SW_64);
      bd = smem_desc(smem_u32(smem) + 2 * A_SLAB, 16, 8 * 64, SW_64);
      astep = 0; bstep = 0;
    } else {
      idesc = instr_desc_bf16(128, p.N, 1, 1);
      ad = smem_desc(smem_u32(smem) + (uint32_t)p.off_rows * 128u, A_SLAB, 8 * 128, SW_128);
      bd = smem_desc(smem_u32(smem) + 2 * A_SLAB, 0, 8 * rowB, rowB >= 128 ? SW_128 : SW_64);
      astep = 128; bstep = rowB;
    }
    const long long t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) umma_bf16(tm + (uint32_t)(i & 3) * 32u, ad + (uint64_t)((i >> 2) % 8 * astep), bd + (uint64_t)((i >> 2) % 8 * bstep), idesc, i >= 4 ? 1u : 0u);
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    p.out[0] = t1 - t0;
  }
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 256);
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 130 * 1024);
  for (int km = 0; km < 2; ++km)
    for (int N : {32, 64})
      for (int off : {0, 1, 2, 8, 17, 18}) {
        long long best = 1LL << 60;
        for (int rep = 0; rep < 5; ++rep) {
          P p{N, off, km, d_out};
          k<<<1, 128, 130 * 1024>>>(p);
          if (cudaDeviceSynchronize() != cudaSuccess) { printf("CUDA error\n"); return 1; }
          long long h;
          cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
          if (h < best) best = h;
        }
        printf("%s  N %3d  A start offset %2d rows: %6.1f cycles per MMA (64 MMAs incl. commit + wait)\n",
               km ? "K-major (64-byte rows, SW64) " : "MN-major (128-byte rows, SW128)", N, off, best / 64.0);
      }
  return 0;
}
