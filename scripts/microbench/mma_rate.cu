// Microbenchmark: cycles per tcgen05.mma (kind::f16, bf16, cta_group::1, SS mode) as a function of
// N, the operand swizzle (row pitch 32/64/128 B) and M, with operands resident in shared memory.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I ../../semi-supervised-gated-lt-vae_b200/csrc
//        -o mma_rate mma_rate.cu
#include <cstdio>
#include <cstdlib>
#include "tc_common.cuh"
using namespace gccvae::tc;

struct P { int M, N, rowb, iters, kslices, a_major_mn, a_rows_distinct; long long* out; };

__global__ void __launch_bounds__(128) k(P p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc(&slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    const uint32_t swz = p.rowb >= 128 ? SW_128 : p.rowb >= 64 ? SW_64 : SW_32;
    const uint32_t idesc = instr_desc_bf16(p.M, p.N, 0, 0);
    const uint64_t ad0 = smem_desc(smem_u32(smem), 16, 8u * p.rowb, swz);
    const uint64_t bd0 = smem_desc(smem_u32(smem) + 48 * 1024, 16, 8u * p.rowb, swz);
    const uint64_t ad1 = ad0 + (16 * 1024 >> 4);
    const long long t0 = clock64();
    umma_bf16(tm, ad0, bd0, idesc, 0u);
    if (p.kslices == 4) {
      for (int i = 0; i < p.iters; i += 8) {
        umma_bf16(tm, ad0, bd0, idesc, 1u); umma_bf16(tm, ad0 + 2, bd0 + 2, idesc, 1u);
        umma_bf16(tm, ad0 + 4, bd0 + 4, idesc, 1u); umma_bf16(tm, ad0 + 6, bd0 + 6, idesc, 1u);
        umma_bf16(tm, ad1, bd0, idesc, 1u); umma_bf16(tm, ad1 + 2, bd0 + 2, idesc, 1u);
        umma_bf16(tm, ad1 + 4, bd0 + 4, idesc, 1u); umma_bf16(tm, ad1 + 6, bd0 + 6, idesc, 1u);
      }
    } else if (p.kslices == 2) {
      for (int i = 0; i < p.iters; i += 8) {
        umma_bf16(tm, ad0, bd0, idesc, 1u); umma_bf16(tm, ad0 + 2, bd0 + 2, idesc, 1u);
        umma_bf16(tm, ad1, bd0, idesc, 1u); umma_bf16(tm, ad1 + 2, bd0 + 2, idesc, 1u);
        umma_bf16(tm, ad0, bd0, idesc, 1u); umma_bf16(tm, ad0 + 2, bd0 + 2, idesc, 1u);
        umma_bf16(tm, ad1, bd0, idesc, 1u); umma_bf16(tm, ad1 + 2, bd0 + 2, idesc, 1u);
      }
    } else {
      for (int i = 0; i < p.iters; i += 8) {
        umma_bf16(tm, ad0, bd0, idesc, 1u); umma_bf16(tm, ad1, bd0, idesc, 1u);
        umma_bf16(tm, ad0, bd0, idesc, 1u); umma_bf16(tm, ad1, bd0, idesc, 1u);
        umma_bf16(tm, ad0, bd0, idesc, 1u); umma_bf16(tm, ad1, bd0, idesc, 1u);
        umma_bf16(tm, ad0, bd0, idesc, 1u); umma_bf16(tm, ad1, bd0, idesc, 1u);
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) p.out[0] = t1 - t0;
  }
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int Ns[] = {16, 32, 64, 128, 256};
  const int rowbs[] = {32, 64, 128};
  for (int grid : {148})
    for (int M : {128, 64})
      for (int rowb : rowbs)
        for (int N : Ns) {
          P p{M, N, rowb, 2000, rowb / 32, 0, 1, d};
          k<<<grid, 128, 100 * 1024>>>(p);
          cudaError_t e = cudaDeviceSynchronize();
          long long c = 0;
          cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
          printf("grid %3d M %3d rowpitch %3d N %3d : %7.1f clk/MMA (%s)\n", grid, M, rowb, N, (double)c / p.iters,
                 cudaGetErrorString(e));
        }
  return 0;
}
