// Microbenchmark 2: a batch of `nm` MMAs (M=128, N, K=16) followed by tcgen05.commit and an mbarrier wait by the
// SAME thread (latency of one tile's worth of MMAs), optionally while 4 other warps hammer TMEM with tcgen05.ld
// (epilogue traffic) or shared memory stores (TMA-like write traffic).
#include <cstdio>
#include <cstdlib>
#include "tc_common.cuh"
using namespace gccvae::tc;

struct P { int N, rowb, batches, nm, ld_traffic, st_traffic; long long* out; };

__global__ void __launch_bounds__(192) k(P p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  __shared__ volatile int done;
  for (int i = threadIdx.x; i < 128 * 1024 / 4; i += 192) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); done = 0; }
  if (threadIdx.x < 32) tmem_alloc(&slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    const uint32_t swz = p.rowb >= 128 ? SW_128 : p.rowb >= 64 ? SW_64 : SW_32;
    const uint32_t idesc = instr_desc_bf16(128, p.N, 0, 0);
    const uint64_t ad0 = smem_desc(smem_u32(smem), 16, 8u * p.rowb, swz);
    const uint64_t bd0 = smem_desc(smem_u32(smem) + 64 * 1024, 16, 8u * p.rowb, swz);
    const int ks = p.rowb / 32;
    const long long t0 = clock64();
    uint32_t ph = 0;
    for (int b = 0; b < p.batches; ++b) {
      uint32_t aoff = 0, boff = 0;
      for (int v = 0; v < p.nm; v += ks) {
        umma_bf16(tm, ad0 + aoff, bd0 + boff, idesc, v > 0 ? 1u : 0u);
        for (int kk = 1; kk < ks; ++kk) umma_bf16(tm, ad0 + aoff + 2u * kk, bd0 + boff + 2u * kk, idesc, 1u);
        aoff += 2048 >> 4;
        boff += 4096 >> 4;
      }
      umma_commit(&bar);
      mbar_wait(&bar, ph);
      ph ^= 1;
    }
    const long long t1 = clock64();
    if (blockIdx.x == 0) p.out[0] = t1 - t0;
    done = 1;
  } else if (warp >= 2) {
    const int q = warp & 3;
    uint32_t acc = 0;
    while (!done) {
      if (p.ld_traffic) {
        uint32_t r[16];
        tmem_ld16(tm + ((uint32_t)(q * 32) << 16) + 256, r);
        tmem_ld_wait();
        acc += r[3];
      }
      if (p.st_traffic) {
        uint4* dst = reinterpret_cast<uint4*>(smem + 96 * 1024) + (threadIdx.x - 64);
#pragma unroll
        for (int j = 0; j < 8; ++j) dst[j * 128] = make_uint4(acc, j, 2, 3);
      }
      if (!p.ld_traffic && !p.st_traffic) __nanosleep(100);
    }
    if (acc == 0x12345) p.out[1] = acc;
  }
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 140 * 1024);
  for (int traffic = 0; traffic < 4; traffic += 3)
    for (int rowb : {64, 128})
      for (int N : {16, 64, 128})
        for (int nm : {4, 8, 16, 32}) {
          P p{N, rowb, 200, nm, traffic & 1, (traffic >> 1) & 1, d};
          k<<<148, 192, 140 * 1024>>>(p);
          cudaError_t e = cudaDeviceSynchronize();
          long long c = 0;
          cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
          printf("tmem-ld %d smem-st %d rowpitch %3d N %3d nm %2d : %8.1f clk/batch  %6.1f clk/MMA (%s)\n", p.ld_traffic,
                 p.st_traffic, rowb, N, nm, (double)c / p.batches, (double)c / p.batches / nm, cudaGetErrorString(e));
        }
  return 0;
}
