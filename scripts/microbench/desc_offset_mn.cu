// Probe (companion of desc_offset.cu): MN-major operands, as the weight-gradient kernel uses them (rows = reduction
// index k = pixels, 128-byte rows of 64 channels, SWIZZLE_128B; B: 64-byte rows of 32 channels, SWIZZLE_64B).
// May the A descriptor start at an arbitrary ROW (= k offset), so that the row-shifted "taps" of a convolution's
// weight gradient read ONE staged halo tile?   D[m][n] = sum_k A[k + off][m] B[k][n],  M = 128 (two 64-channel slabs), N = 32.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I ../../semi-supervised-gated-lt-vae_b200/csrc
//        -o desc_offset_mn desc_offset_mn.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "tc_common.cuh"
using namespace gccvae::tc;

struct P { int off_rows, ksteps, a_rows; float* out; };
__host__ __device__ inline int aval(int k, int m) { return (k * 7 + m * 3) % 13 - 6; }
__host__ __device__ inline int bval(int k, int n) { return (n * 5 + k) % 7 - 3; }
constexpr int A_SLAB = 48 * 1024;   // bytes between the two 64-channel slabs of A (LBO)

__global__ void __launch_bounds__(128) k(P p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  uint8_t* sA = smem;
  uint8_t* sB = smem + 2 * A_SLAB;
  for (int i = threadIdx.x; i < p.a_rows * 128; i += 128) {
    const int r = i / 128, m = i % 128;
    uint32_t a = (uint32_t)(r * 128 + (m % 64) * 2);
    a ^= ((a >> 7) & 7u) << 4;
    *reinterpret_cast<__nv_bfloat16*>(sA + (m / 64) * A_SLAB + a) = __float2bfloat16((float)aval(r, m));
  }
  for (int i = threadIdx.x; i < 16 * p.ksteps * 32; i += 128) {
    const int r = i / 32, n = i % 32;
    uint32_t a = (uint32_t)(r * 64 + n * 2);
    a ^= ((a >> 7) & 3u) << 4;
    *reinterpret_cast<__nv_bfloat16*>(sB + a) = __float2bfloat16((float)bval(r, n));
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc(&slot, 32);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = instr_desc_bf16(128, 32, 1, 1);
    const uint64_t ad = smem_desc(smem_u32(sA) + (uint32_t)p.off_rows * 128u, A_SLAB, 8 * 128, SW_128);
    const uint64_t bd = smem_desc(smem_u32(sB), 0, 8 * 64, SW_64);
    for (int s = 0; s < p.ksteps; ++s) umma_bf16(tm, ad + (uint64_t)(s * 128), bd + (uint64_t)(s * 64), idesc, s > 0 ? 1u : 0u);
    umma_commit(&bar);
    mbar_wait(&bar, 0);
  }
  __syncthreads();
  tc_fence_after();
  for (int c0 = 0; c0 < 32; c0 += 16) {
    uint32_t acc[16];
    tmem_ld16(tm + ((uint32_t)((threadIdx.x >> 5) * 32) << 16) + (uint32_t)c0, acc);
    tmem_ld_wait();
    for (int n = 0; n < 16; ++n) p.out[threadIdx.x * 32 + c0 + n] = __uint_as_float(acc[n]);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 32);
}

int main() {
  float* d_out;
  cudaMalloc(&d_out, 128 * 32 * sizeof(float));
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024);
  const int offs[] = {0, 1, 2, 3, 7, 8, 9, 17, 18, 33, 34};
  for (int ksteps : {2, 9}) {
    for (int off : offs) {
      P p{off, ksteps, 16 * ksteps + 40, d_out};
      cudaMemset(d_out, 0xff, 128 * 32 * sizeof(float));
      k<<<1, 128, 120 * 1024>>>(p);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("off %d: CUDA error %s\n", off, cudaGetErrorString(e)); return 1; }
      std::vector<float> h(128 * 32);
      cudaMemcpy(h.data(), d_out, h.size() * sizeof(float), cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 32; ++n) {
          float ref = 0.f;
          for (int kk = 0; kk < 16 * ksteps; ++kk) ref += (float)(aval(kk + off, m) * bval(kk, n));
          if (h[m * 32 + n] != ref) ++bad;
        }
      printf("MN-major A (128-byte rows, SW128) x MN-major B (64-byte rows, SW64), K = %3d, A start offset %2d rows: %s (%d wrong)\n",
             16 * ksteps, off, bad == 0 ? "EXACT" : "WRONG", bad);
    }
  }
  return 0;
}
