// Data-parallel gradient exchange fused with the optimiser: ONE kernel per call does
//     barrier over the ranks  ->  two-shot all-reduce (sum) of a range of the flat fp32 gradient buffer over NVLink peer
//     memory  ->  barrier  ->  the Keras-Adam update of that range on this rank's replica (gated_ccvae.py:309-310).
// The reference is single-process; this is the exchange step of the data-parallel design.
//
// Every rank's gradient buffer lives in symmetric memory (torch.distributed._symmetric_memory: the same allocation
// mapped into every process of the node), so a rank reads and writes its peers' buffers with plain loads / stores that
// travel over NVLink / NVSwitch:
//   shot 1  rank r sums ITS shard [lo_r, hi_r) of the range over all ranks' buffers (fixed rank order: the result is
//           bit-identical whichever rank computes it) ...
//   shot 2  ... and writes the sum into every rank's buffer at the same place.
// Per rank and call: (W - 1) / W of the range in, the same out - against (W - 1) x the range for a one-shot exchange.
// After the second barrier every buffer holds the complete sum and nobody touches a peer's buffer any more, so the
// Adam update reads the local buffer and clears it behind the read for the next backward pass.
// Being one ordinary kernel it can be captured into the step's CUDA graph and runs on a side stream UNDER the last
// dgrad for everything but the first layer's parameters (whose gradients are the last to complete); NCCL's all-reduce
// sat exposed between the replayed graph and the optimiser (0.93 / 0.90 scaling efficiency at 2 / 8 GPUs in round 1).
//
// Rank barrier: per rank a block of 32-bit words in symmetric memory, all MONOTONIC counters (no reset race):
//   word q (< 16)  arrival flag written by rank q,   word 16 local block-arrival counter (reset by the last block),
//   word 17 local gate,   word 18 number of completed calls,   word 19 exit ticket,   word 20 error flag.
// Block 0 of every rank signals all peers (st.release.sys) and waits for all of them (ld.acquire.sys); the other
// blocks wait at the local gate.  All waits are bounded (~20 s): a lost peer traps instead of hanging the GPU.
#include <string.h>

#include "common.cuh"

namespace gccvae {

constexpr int DP_THREADS = 256;
constexpr int DPW_ARRIVE = 16, DPW_GATE = 17, DPW_CALLS = 18, DPW_TICKET = 19, DPW_ERROR = 20;

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_relaxed_sys_f4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_relaxed_sys_f(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// wait until *p >= want (monotonic counters); sys = the word is written by another GPU
template <bool SYS>
__device__ __forceinline__ void wait_ge(const uint32_t* p, uint32_t want, uint32_t* err) {
  const unsigned long long t0 = globaltimer_ns();
  for (uint32_t spin = 0;; ++spin) {
    const uint32_t v = SYS ? ld_acquire_sys(p) : ld_acquire_gpu(p);
    if ((int32_t)(v - want) >= 0) return;
    if ((spin & 1023u) == 1023u && globaltimer_ns() - t0 > 20000000000ull) {
      *err = 1u;
      printf("gccvae: data-parallel barrier timed out (block %d thread %d: have %u want %u)\n", blockIdx.x, threadIdx.x, v,
             want);
      __trap();
    }
    __nanosleep(64);
  }
}

// all blocks of all ranks; `local_first`: every local block's earlier (peer) stores are complete before any rank passes
__device__ __forceinline__ void world_barrier(const gccvae_dp_args& a, uint32_t* sync, uint32_t value, bool local_first) {
  __syncthreads();
  if (blockIdx.x == 0) {
    if (local_first && threadIdx.x == 0) {
      __threadfence_system();
      wait_ge<false>(sync + DPW_ARRIVE, gridDim.x - 1, sync + DPW_ERROR);
      __threadfence_system();
    }
    __syncthreads();
    if ((int)threadIdx.x < a.world) {
      st_release_sys(reinterpret_cast<uint32_t*>(a.sync[threadIdx.x]) + a.rank, value);
      wait_ge<true>(sync + threadIdx.x, value, sync + DPW_ERROR);
    }
    __syncthreads();
    if (threadIdx.x == 0) st_release_gpu(sync + DPW_GATE, value);
  } else {
    if (threadIdx.x == 0) {
      if (local_first) {
        __threadfence_system();
        atomicAdd(sync + DPW_ARRIVE, 1u);
      }
      wait_ge<false>(sync + DPW_GATE, value, sync + DPW_ERROR);
    }
  }
  __syncthreads();
}

// Keras Adam (tf.keras.optimizers.Adam, epsilon outside the root) on 4 consecutive elements
__device__ __forceinline__ void adam4(const gccvae_dp_args& a, long long e, const float4 g4, float lr_t, float omb1, float omb2) {
  float4 m4 = *reinterpret_cast<float4*>(a.m + e), v4 = *reinterpret_cast<float4*>(a.v + e);
  float4 p4 = *reinterpret_cast<float4*>(a.param + e);
  m4.x += (g4.x - m4.x) * omb1; m4.y += (g4.y - m4.y) * omb1; m4.z += (g4.z - m4.z) * omb1; m4.w += (g4.w - m4.w) * omb1;
  v4.x += (g4.x * g4.x - v4.x) * omb2; v4.y += (g4.y * g4.y - v4.y) * omb2;
  v4.z += (g4.z * g4.z - v4.z) * omb2; v4.w += (g4.w * g4.w - v4.w) * omb2;
  p4.x -= lr_t * m4.x / (sqrtf(v4.x) + a.eps); p4.y -= lr_t * m4.y / (sqrtf(v4.y) + a.eps);
  p4.z -= lr_t * m4.z / (sqrtf(v4.z) + a.eps); p4.w -= lr_t * m4.w / (sqrtf(v4.w) + a.eps);
  *reinterpret_cast<float4*>(a.m + e) = m4;
  *reinterpret_cast<float4*>(a.v + e) = v4;
  *reinterpret_cast<float4*>(a.param + e) = p4;
}
__device__ __forceinline__ float adam_lr_t(const gccvae_dp_args& a, int t) {
  return (float)((double)a.lr * sqrt(1.0 - pow((double)a.beta2, (double)t)) / (1.0 - pow((double)a.beta1, (double)t)));
}
// result ring + exit ticket: the last block closes the call (every block has read `calls` and passed the barriers)
__device__ __forceinline__ void finish_call(const gccvae_dp_args& a, uint32_t* sync, uint32_t calls, int t, float loss) {
  if (a.publish && a.result_ring != nullptr && blockIdx.x == 0) {
    float* dst = a.result_ring + (size_t)((t - 1) % a.ring_slots) * GCCVAE_RESULT_SLOT_FLOATS;
    for (int i = threadIdx.x; i < 1 + GCCVAE_ZC * GCCVAE_Y; i += DP_THREADS)
      dst[i] = i == 0 ? (a.loss_index >= 0 ? loss : a.result_loss[0]) : a.result_c[i - 1];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const uint32_t ticket = atomicAdd(sync + DPW_TICKET, 1u);
    if (ticket == gridDim.x - 1) {
      sync[DPW_TICKET] = 0u;
      sync[DPW_ARRIVE] = 0u;
      if (a.publish) a.step_state[0] = t;
      __threadfence();
      st_release_gpu(sync + DPW_CALLS, calls + 1u);
    }
  }
}

__global__ void __launch_bounds__(DP_THREADS) dp_reduce_adam_kernel(const gccvae_dp_args a) {
  __shared__ uint32_t s_calls;
  __shared__ float s_loss;
  uint32_t* sync = reinterpret_cast<uint32_t*>(a.sync[a.rank]);
  float* grad = a.grad[a.rank];
  if (threadIdx.x == 0) s_calls = ld_acquire_gpu(sync + DPW_CALLS);
  __syncthreads();
  const uint32_t calls = s_calls;
  const int W = a.world;
  // ---- all ranks' gradients of the range are complete ---------------------------------------------------------------
  world_barrier(a, sync, 2u * calls + 1u, false);
  // ---- shot 1 + 2: this rank's shard, summed in rank order, written to every rank --------------------------------------
  const long long n4 = (a.n - a.i0) >> 2;                   // the range is 16-byte aligned on both ends
  const long long per = (n4 + W - 1) / W;
  const long long lo = (long long)a.rank * per, hi = min(n4, lo + per);
  for (long long i = lo + blockIdx.x * (long long)DP_THREADS + threadIdx.x; i < hi; i += (long long)gridDim.x * DP_THREADS) {
    const long long e = a.i0 + 4 * i;
    float4 acc = ld_relaxed_sys_f4(a.grad[0] + e);
    for (int q = 1; q < W; ++q) {
      const float4 v = ld_relaxed_sys_f4(a.grad[q] + e);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    for (int q = 0; q < W; ++q) *reinterpret_cast<float4*>(a.grad[q] + e) = acc;
  }
  // the loss rides in one extra slot: every rank sums it for itself (one-shot) and keeps the sum until after the barrier
  if (a.loss_index >= 0 && blockIdx.x == 0 && threadIdx.x == 0) {
    float v = 0.0f;
    for (int q = 0; q < W; ++q) v += ld_relaxed_sys_f(a.grad[q] + a.loss_index);
    s_loss = v;
  }
  // ---- every shard has landed everywhere; peers are done with this rank's buffer --------------------------------------
  world_barrier(a, sync, 2u * calls + 2u, true);
  if (a.loss_index >= 0 && blockIdx.x == 0 && threadIdx.x == 0) grad[a.loss_index] = s_loss;
  // ---- Adam on the local replica, gradient cleared behind the read ------------------------------------------------------
  const int t = a.step_state[0] + 1;
  const float lr_t = adam_lr_t(a, t);
  const float omb1 = 1.0f - a.beta1, omb2 = 1.0f - a.beta2;
  const long long z4 = (a.n_zero - a.i0) >> 2;              // n_zero is 16-byte aligned like the range
  for (long long i = blockIdx.x * (long long)DP_THREADS + threadIdx.x; i < z4; i += (long long)gridDim.x * DP_THREADS) {
    const long long e = a.i0 + 4 * i;
    if (e < a.n) adam4(a, e, __ldcg(reinterpret_cast<const float4*>(grad + e)), lr_t, omb1, omb2);
    *reinterpret_cast<float4*>(grad + e) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  finish_call(a, sync, calls, t, a.loss_index >= 0 && blockIdx.x == 0 ? s_loss : 0.0f);
}

// The same exchange for a SHORT range (the last gradients of the step: a few thousand elements on the critical path) with
// ONE barrier instead of two: every rank pushes its gradients into slot `rank` of every rank's receive buffer, one
// barrier, then every rank sums its own receive slots in rank order - nobody reads a peer's memory after the barrier, so
// the local gradient buffer can be cleared at once.  The receive slots are rewritten by the next push only; the caller
// guarantees a two-barrier call (the bulk of the next step's exchange) between two pushes.
__global__ void __launch_bounds__(DP_THREADS) dp_push_adam_kernel(const gccvae_dp_args a) {
  __shared__ uint32_t s_calls;
  __shared__ float s_loss;
  uint32_t* sync = reinterpret_cast<uint32_t*>(a.sync[a.rank]);
  float* grad = a.grad[a.rank];
  if (threadIdx.x == 0) s_calls = ld_acquire_gpu(sync + DPW_CALLS);
  __syncthreads();
  const uint32_t calls = s_calls;
  const int W = a.world;
  const long long n4 = (a.n - a.i0) >> 2;
  const size_t mine = (size_t)a.rank * a.recv_stride;
  for (long long i = blockIdx.x * (long long)DP_THREADS + threadIdx.x; i < n4; i += (long long)gridDim.x * DP_THREADS) {
    const float4 g4 = __ldcg(reinterpret_cast<const float4*>(grad + a.i0 + 4 * i));
    for (int q = 0; q < W; ++q) *reinterpret_cast<float4*>(a.recv[q] + mine + 4 * i) = g4;
  }
  if (a.loss_index >= 0 && blockIdx.x == 0 && threadIdx.x == 0) {
    const float lv = __ldcg(grad + a.loss_index);
    for (int q = 0; q < W; ++q) a.recv[q][mine + 4 * n4] = lv;
  }
  world_barrier(a, sync, 2u * calls + 2u, true);
  const float* rcv = a.recv[a.rank];
  if (a.loss_index >= 0 && blockIdx.x == 0 && threadIdx.x == 0) {
    float v = 0.0f;
    for (int q = 0; q < W; ++q) v += ld_relaxed_sys_f(rcv + (size_t)q * a.recv_stride + 4 * n4);
    s_loss = v;
    grad[a.loss_index] = v;
  }
  const int t = a.step_state[0] + 1;
  const float lr_t = adam_lr_t(a, t);
  const float omb1 = 1.0f - a.beta1, omb2 = 1.0f - a.beta2;
  const long long z4 = (a.n_zero - a.i0) >> 2;
  for (long long i = blockIdx.x * (long long)DP_THREADS + threadIdx.x; i < z4; i += (long long)gridDim.x * DP_THREADS) {
    const long long e = a.i0 + 4 * i;
    if (e < a.n) {
      float4 acc = ld_relaxed_sys_f4(rcv + 4 * i);
      for (int q = 1; q < W; ++q) {
        const float4 v = ld_relaxed_sys_f4(rcv + (size_t)q * a.recv_stride + 4 * i);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      adam4(a, e, acc, lr_t, omb1, omb2);
    }
    *reinterpret_cast<float4*>(grad + e) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  finish_call(a, sync, calls, t, a.loss_index >= 0 && blockIdx.x == 0 ? s_loss : 0.0f);
}

}  // namespace gccvae

using namespace gccvae;

extern "C" int gccvae_dp_reduce_adam_f32(const gccvae_dp_args* a_in, void* stream) {
  GCC_REQUIRE(a_in, "dp_reduce_adam: null args");
  const gccvae_dp_args& a = *a_in;
  GCC_REQUIRE(a.world >= 2 && a.world <= GCCVAE_DP_MAX_RANKS && a.rank >= 0 && a.rank < a.world, "dp_reduce_adam: bad rank %d/%d",
              a.rank, a.world);
  for (int q = 0; q < a.world; ++q)
    GCC_REQUIRE(a.grad[q] && a.sync[q] && (uintptr_t)a.grad[q] % 16 == 0, "dp_reduce_adam: bad peer pointer %d", q);
  GCC_REQUIRE(a.param && a.m && a.v && a.step_state, "dp_reduce_adam: null pointer");
  GCC_REQUIRE(a.i0 >= 0 && a.n > a.i0 && a.n_zero >= a.n && a.i0 % 4 == 0 && a.n % 4 == 0 && a.n_zero % 4 == 0,
              "dp_reduce_adam: range [%lld, %lld) / %lld must be 16-byte aligned", a.i0, a.n, a.n_zero);
  GCC_REQUIRE(((uintptr_t)a.param | (uintptr_t)a.m | (uintptr_t)a.v) % 16 == 0, "dp_reduce_adam: buffers must be 16-byte aligned");
  GCC_REQUIRE(a.result_ring == nullptr || (a.result_c && a.ring_slots > 0 && (a.loss_index >= 0 || a.result_loss)),
              "dp_reduce_adam: bad result ring");
  long long blocks = (a.n_zero - a.i0 + DP_THREADS * 4 - 1) / (DP_THREADS * 4);
  if (blocks > 296) blocks = 296;   // every block waits at the barriers: the whole grid must fit on the device at once
  if (blocks < 1) blocks = 1;
  if (a.push) {
    GCC_REQUIRE(a.recv_stride >= (a.n - a.i0) + 4 && a.recv_stride % 4 == 0, "dp_reduce_adam: receive slots of %lld floats are too small",
                a.recv_stride);
    for (int q = 0; q < a.world; ++q)
      GCC_REQUIRE(a.recv[q] && (uintptr_t)a.recv[q] % 16 == 0, "dp_reduce_adam: bad receive buffer %d", q);
    dp_push_adam_kernel<<<(int)blocks, DP_THREADS, 0, (cudaStream_t)stream>>>(a);
  } else {
    dp_reduce_adam_kernel<<<(int)blocks, DP_THREADS, 0, (cudaStream_t)stream>>>(a);
  }
  GCC_CHECK_LAUNCH("dp_reduce_adam");
  return GCCVAE_OK;
}
