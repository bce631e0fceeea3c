// Microbenchmark / probe: may the start address of a K-major, swizzled UMMA shared-memory descriptor be moved by an
// ARBITRARY number of operand rows (not a multiple of the 8-row swizzle atom)?  If the swizzle XOR is taken from the
// absolute shared-memory address (as TMA writes it), a tile stored once with a halo can serve several row-shifted
// "taps" of a convolution through descriptor offsets alone (one TMA box instead of one per tap).
// A [rows x K] and B [16 x K] are written with the swizzle pattern of a 1024-byte aligned buffer; for each row offset
// the MMA result D[m][n] = sum_k A[m + off][k] B[n][k] is compared with the host's (small integers: exact in bf16/fp32).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I ../../semi-supervised-gated-lt-vae_b200/csrc
//        -o desc_offset desc_offset.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "tc_common.cuh"
using namespace gccvae::tc;

struct P { int rowb, off_rows, use_base_offset, a_rows; float* out; };

__host__ __device__ inline int aval(int r, int k) { return (r * 7 + k * 3) % 13 - 6; }
__host__ __device__ inline int bval(int n, int k) { return (n * 5 + k) % 7 - 3; }

__global__ void __launch_bounds__(128) k(P p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int kdim = p.rowb / 2;
  const int bbits = p.rowb >= 128 ? 3 : p.rowb >= 64 ? 2 : 1;
  uint8_t* sA = smem;
  uint8_t* sB = smem + 32 * 1024;
  for (int i = threadIdx.x; i < p.a_rows * kdim; i += 128) {
    const int r = i / kdim, kk = i % kdim;
    uint32_t a = (uint32_t)(r * p.rowb + kk * 2);
    a ^= ((a >> 7) & ((1u << bbits) - 1u)) << 4;
    *reinterpret_cast<__nv_bfloat16*>(sA + a) = __float2bfloat16((float)aval(r, kk));
  }
  for (int i = threadIdx.x; i < 16 * kdim; i += 128) {
    const int r = i / kdim, kk = i % kdim;
    uint32_t a = (uint32_t)(r * p.rowb + kk * 2);
    a ^= ((a >> 7) & ((1u << bbits) - 1u)) << 4;
    *reinterpret_cast<__nv_bfloat16*>(sB + a) = __float2bfloat16((float)bval(r, kk));
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc(&slot, 32);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    const uint32_t swz = p.rowb >= 128 ? SW_128 : p.rowb >= 64 ? SW_64 : SW_32;
    const uint32_t idesc = instr_desc_bf16(128, 16, 0, 0);
    const uint32_t a_start = smem_u32(sA) + (uint32_t)(p.off_rows * p.rowb);
    uint64_t ad = smem_desc(a_start, 16, 8u * p.rowb, swz);
    if (p.use_base_offset) ad |= (uint64_t)((a_start >> 7) & 7u) << 49;
    const uint64_t bd = smem_desc(smem_u32(sB), 16, 8u * p.rowb, swz);
    for (int s = 0; s < p.rowb / 32; ++s) umma_bf16(tm, ad + 2 * s, bd + 2 * s, idesc, s > 0 ? 1u : 0u);
    umma_commit(&bar);
    mbar_wait(&bar, 0);
  }
  __syncthreads();
  tc_fence_after();
  uint32_t acc[16];
  tmem_ld16(tm + ((uint32_t)((threadIdx.x >> 5) * 32) << 16), acc);
  tmem_ld_wait();
  for (int n = 0; n < 16; ++n) p.out[threadIdx.x * 16 + n] = __uint_as_float(acc[n]);
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 32);
}

int main() {
  float* d_out;
  cudaMalloc(&d_out, 128 * 16 * sizeof(float));
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
  const int offs[] = {0, 1, 2, 3, 4, 5, 7, 8, 9, 16, 33, 34, 35, 66};
  for (int rowb : {32, 64, 128}) {
    for (int ubo = 0; ubo < 2; ++ubo) {
      for (int off : offs) {
        P p{rowb, off, ubo, 128 + 70, d_out};
        cudaMemset(d_out, 0xff, 128 * 16 * sizeof(float));
        k<<<1, 128, 80 * 1024>>>(p);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("rowb %d off %d: CUDA error %s\n", rowb, off, cudaGetErrorString(e)); return 1; }
        std::vector<float> h(128 * 16);
        cudaMemcpy(h.data(), d_out, h.size() * sizeof(float), cudaMemcpyDeviceToHost);
        int bad = 0, bad_rows = 0;
        for (int m = 0; m < 128; ++m) {
          int rb = 0;
          for (int n = 0; n < 16; ++n) {
            float ref = 0.f;
            for (int kk = 0; kk < rowb / 2; ++kk) ref += (float)(aval(m + off, kk) * bval(n, kk));
            if (h[m * 16 + n] != ref) { ++bad; rb = 1; }
          }
          bad_rows += rb;
        }
        printf("row bytes %3d  base_offset field %s  start offset %2d rows (%5d B): %s (%d wrong values in %d rows)\n", rowb,
               ubo ? "set " : "zero", off, off * rowb, bad == 0 ? "EXACT" : "WRONG", bad, bad_rows);
      }
    }
  }
  return 0;
}
