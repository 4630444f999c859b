// Evaluation-time gather of the per-clip logits across the GPUs of one node (the reference gathers predictions
// on the host, train.py:150-170 concatenates them batch by batch) as PUSHES over NVLink peer memory instead of a collective.
//
// Every rank owns one peer-mapped block (torch symmetric memory: the same layout on every rank, all W base pointers known to all):
//
//     table[2][W][n]   fp32   slot s = step & 1; row block r = the n = clips x 21 logits rank r computed in that step
//     flags[2][W]      u32    flags[s][r] = step + 1 once rank r's block of that step is complete in THIS rank's table
//
// logits_push (one CTA per peer, launched right behind the fusion head): stores this rank's logits straight into every peer's table
// (16-byte st.global over NVLink), then releases the peer's flag at system scope.  Nothing waits for anybody: a rank that is ahead
// just leaves its block in the others' tables.  logits_wait (one CTA, at the END of the step, behind the SFormer kernel that
// follows the head in the captured graph): spins until all W flags of this step's slot show step + 1 — by then the pushes are
// usually hundreds of microseconds old — and advances the step counter.  The hard per-step rendezvous of an all-gather (NCCL in
// the stream: 76 us per step on 8 B200s, mostly rank skew) becomes a bounded-lag pipeline: rank A may finish step i while a peer is
// still anywhere behind its own push of step i.
//
// Two slots are enough: A pushes step i+2 only after its wait of step i+1, which saw B's push of step i+1, which B issued (stream
// order) after everything B did with the table of step i.  The spin is bounded (timeout -> error word, never a hung GPU).
#include "avf_common.cuh"
#include "avf_internal.h"

namespace avf {
namespace {

__device__ __forceinline__ void st_release_sys_u32(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__host__ __device__ inline size_t peer_flags_offset(int world, size_t n) { return (size_t(2) * world * n * sizeof(float) + 127) / 128 * 128; }

// grid = W (peer), block = 256.  state[0] = step counter (advanced by logits_wait_kernel of the same stream).
__global__ void __launch_bounds__(256) logits_push_kernel(const float* __restrict__ logits, size_t n, const unsigned long long* __restrict__ peer_base, int world,
                                                          int rank, const uint32_t* __restrict__ state) {
  const uint32_t step = state[0];
  const uint32_t slot = step & 1u;
  uint8_t* base = reinterpret_cast<uint8_t*>(peer_base[blockIdx.x]);
  float* dst = reinterpret_cast<float*>(base) + (size_t(slot) * world + rank) * n;
  if ((n & 3u) == 0 && (reinterpret_cast<uintptr_t>(logits) & 15u) == 0) {
    const float4* s4 = reinterpret_cast<const float4*>(logits);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (size_t i = threadIdx.x; i < n / 4; i += blockDim.x) d4[i] = s4[i];
  } else {
    for (size_t i = threadIdx.x; i < n; i += blockDim.x) dst[i] = logits[i];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    uint32_t* flags = reinterpret_cast<uint32_t*>(base + peer_flags_offset(world, n));
    st_release_sys_u32(flags + slot * world + rank, step + 1u);
  }
}

// grid = 1, block = 32 * ceil(W / 32).  state[0] = step counter, state[1] = error word (0 = fine, 1 + r = rank r's block timed out).
__global__ void logits_wait_kernel(const uint8_t* __restrict__ my_base, size_t n, int world, uint32_t* state, unsigned long long timeout_ns) {
  const uint32_t step = state[0];
  const uint32_t slot = step & 1u;
  const uint32_t* flags = reinterpret_cast<const uint32_t*>(my_base + peer_flags_offset(world, n)) + slot * world;
  if (int(threadIdx.x) < world) {
    const unsigned long long t0 = global_timer_ns();
    uint32_t spins = 0;
    while (int32_t(ld_acquire_sys_u32(flags + threadIdx.x) - (step + 1u)) < 0) {
      if ((++spins & 1023u) == 0 && global_timer_ns() - t0 > timeout_ns) {
        atomicCAS(&state[1], 0u, 1u + threadIdx.x);
        break;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) state[0] = step + 1u;
}

}  // namespace

size_t peer_gather_bytes(int world, size_t n) { return peer_flags_offset(world, n) + size_t(2) * world * sizeof(uint32_t); }

int logits_push(const float* logits, size_t n, const unsigned long long* peer_base, int world, int rank, const uint32_t* state, cudaStream_t st) {
  logits_push_kernel<<<world, 256, 0, st>>>(logits, n, peer_base, world, rank, state);
  AVF_LAUNCH_CHECK("logits_push_kernel");
  return 0;
}

int logits_wait(const void* my_base, size_t n, int world, uint32_t* state, unsigned long long timeout_ns, cudaStream_t st) {
  logits_wait_kernel<<<1, 32 * ceil_div(world, 32), 0, st>>>(static_cast<const uint8_t*>(my_base), n, world, state, timeout_ns);
  AVF_LAUNCH_CHECK("logits_wait_kernel");
  return 0;
}

}  // namespace avf
