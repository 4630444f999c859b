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
#include <algorithm>
#include <cmath>

#include "avf_adam.cuh"
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

// ---- gradient all-reduce (sum) of the flat fp32 bucket over peer memory -------------------------------------------------------
// Training data parallelism: after the backward pass every rank holds its shard's gradient sum in a flat
// bucket of n floats (41.7 MB for the hot path).  The bucket lives in a peer-mapped block, followed by two flag rows:
//
//     grad[n_pad] fp32 | enter[W] u32 | exit[W] u32            n_pad = n rounded up to a multiple of 4 * W
//
// One kernel, two-shot, in place: (1) tell every peer "my bucket is complete" and wait for theirs; (2) rank r sums slice r of all W
// buckets (16-byte loads straight from the peers' memory, ranks added in the fixed order 0..W-1, so the result is deterministic and
// identical everywhere) and stores the sum into slice r of ALL buckets; (3) tell every peer "my slice is in your bucket, I have stopped
// reading yours" and wait for theirs.  Per GPU (W-1)/W of the bucket crosses NVLink once in each direction and nothing is staged,
// against NCCL's ring / tree of the same 41.7 MB at 0.25 ms on 8 B200s.  Spins are bounded (timeout -> error word, message, trap).
// Bucket loads: weak ld.global.cg (L2 only).  They are ordered behind the acquire of the peers' "bucket complete" flags (ld.acquire.sys
// + bar.sync), which is all the memory model asks for, and — unlike volatile / .relaxed.sys loads, which ptxas issued in dependent
// groups of 4 + 2 + 2 — the eight of them go out back to back, one NVLink round trip per element.
__device__ __forceinline__ float4 ld_bucket_f4(const float4* p) { return __ldcg(p); }

__device__ __forceinline__ bool spin_until(const uint32_t* flag, uint32_t seq, unsigned long long timeout_ns) {
  const unsigned long long t0 = global_timer_ns();
  uint32_t spins = 0;
  while (int32_t(ld_acquire_sys_u32(flag) - seq) < 0) {
    if ((++spins & 1023u) == 0 && global_timer_ns() - t0 > timeout_ns) return false;
  }
  return true;
}

// A peer that never shows up: gradients that were not summed must not reach the optimiser silently.  Record who was missing, say so,
// and stop the kernel with an error the host sees at its next synchronisation (the training step cannot be saved anyway).
__device__ __noinline__ void give_up(uint32_t* state, uint32_t code, int rank, const char* where) {
  atomicCAS(&state[1], 0u, code);
  __threadfence_system();
  printf("avf: gradient all-reduce on rank %d: rank %u did not reach the %s of the reduction in time\n", rank, (code - 1u) % 100u, where);
  __trap();
}

constexpr int AR_THREADS = 256;
constexpr int AR_MAX_WORLD = 16;

// state: [0] sequence number of the last completed reduction, [1] error word, [2] CTAs of this launch that finished their slice
// The optimiser step that follows the reduction, fused into the same kernel (avf_adam_allreduce_step): once the sums of all slices have
// landed in this rank's bucket every CTA runs the Adam update of its share of the WHOLE bucket (weights are replicated: every rank
// updates all of them, with identical inputs and therefore identical results).
struct FusedAdam {
  float *p, *m, *v;
  __nv_bfloat16* shadow;
  size_t n;              // elements of the bucket that carry parameters (n <= n_pad)
  AdamParams a;
};

template <int W, bool ADAM>
__global__ void __launch_bounds__(AR_THREADS) grad_allreduce_kernel(const unsigned long long* __restrict__ peer_base, size_t n_pad, int world_rt, int rank,
                                                                    uint32_t* state, unsigned long long timeout_ns, FusedAdam ad) {
  const int world = W > 0 ? W : world_rt;
  const uint32_t seq = state[0] + 1u;
  __shared__ unsigned long long base[AR_MAX_WORLD];
  if (int(threadIdx.x) < world) base[threadIdx.x] = peer_base[threadIdx.x];
  __syncthreads();
  const size_t flag_off = n_pad * sizeof(float);
  uint32_t* my_enter = reinterpret_cast<uint32_t*>(base[rank] + flag_off);
  uint32_t* my_exit = my_enter + world;

  // (1) my bucket is complete (stream order: the backward kernels are done); wait until everybody's is
  if (blockIdx.x == 0 && int(threadIdx.x) < world) {
    __threadfence_system();
    st_release_sys_u32(reinterpret_cast<uint32_t*>(base[threadIdx.x] + flag_off) + rank, seq);
  }
  if (int(threadIdx.x) < world && !spin_until(my_enter + threadIdx.x, seq, timeout_ns)) give_up(state, 1u + threadIdx.x, rank, "entry");
  __syncthreads();

  // (2) reduce my slice of all buckets, store it everywhere
  const size_t slice4 = n_pad / 4 / world;
  const size_t lo4 = slice4 * rank;
  constexpr int NW = W > 0 ? W : AR_MAX_WORLD;          // loads per element
  constexpr int U = NW >= 8 ? 1 : 8 / NW;               // elements per thread and trip: eight 16-byte loads in flight whatever W is
  const size_t stride = size_t(gridDim.x) * AR_THREADS;
  float4* buck[NW];                                     // slice `rank` of every bucket, in registers
#pragma unroll
  for (int r = 0; r < NW; ++r) buck[r] = reinterpret_cast<float4*>(base[r < world ? r : 0]) + lo4;
  for (size_t i0 = size_t(blockIdx.x) * AR_THREADS + threadIdx.x; i0 < slice4; i0 += stride * U) {
    float4 v[U][NW];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t i = i0 + u * stride;
#pragma unroll
      for (int r = 0; r < NW; ++r)
        if (r < world && i < slice4) v[u][r] = ld_bucket_f4(buck[r] + i);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t i = i0 + u * stride;
      if (i >= slice4) break;
      float4 acc = v[u][0];
#pragma unroll
      for (int r = 1; r < NW; ++r)
        if (r < world) { acc.x += v[u][r].x; acc.y += v[u][r].y; acc.z += v[u][r].z; acc.w += v[u][r].w; }
#pragma unroll
      for (int r = 0; r < NW; ++r)
        if (r < world) buck[r][i] = acc;
    }
  }
  __syncthreads();

  // (3) the last CTA of this rank to finish tells every peer; CTA 0 waits for everybody's slice to have landed here
  if (threadIdx.x == 0) {
    __threadfence_system();
    const uint32_t prev = atomicAdd(&state[2], 1u);
    if (prev == gridDim.x - 1) {
      state[2] = 0;
      __threadfence_system();
      for (int p = 0; p < world; ++p) st_release_sys_u32(reinterpret_cast<uint32_t*>(base[p] + flag_off) + world + rank, seq);
    }
  }
  if (blockIdx.x == 0 || ADAM) {
    if (int(threadIdx.x) < world && !spin_until(my_exit + threadIdx.x, seq, timeout_ns)) give_up(state, 101u + threadIdx.x, rank, "exit");
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0) state[0] = seq;
  }
  if constexpr (ADAM) {
    // (4) Adam over the whole (now summed) bucket; 16-byte groups, the bucket length is a multiple of 4 (FusedAdam pads every parameter to 8)
    const float4* g4 = reinterpret_cast<const float4*>(base[rank]);
    float4 *p4 = reinterpret_cast<float4*>(ad.p), *m4 = reinterpret_cast<float4*>(ad.m), *v4 = reinterpret_cast<float4*>(ad.v);
    for (size_t i = size_t(blockIdx.x) * AR_THREADS + threadIdx.x; i < ad.n / 4; i += size_t(gridDim.x) * AR_THREADS) {
      const float4 g = __ldcg(g4 + i);
      float4 pv = p4[i], mv = m4[i], vv = v4[i];
      adam_update(pv.x, g.x, mv.x, vv.x, ad.a);
      adam_update(pv.y, g.y, mv.y, vv.y, ad.a);
      adam_update(pv.z, g.z, mv.z, vv.z, ad.a);
      adam_update(pv.w, g.w, mv.w, vv.w, ad.a);
      p4[i] = pv; m4[i] = mv; v4[i] = vv;
      if (ad.shadow != nullptr) {
        __nv_bfloat162 lo = __floats2bfloat162_rn(pv.x, pv.y), hi = __floats2bfloat162_rn(pv.z, pv.w);
        uint2 o;
        o.x = *reinterpret_cast<uint32_t*>(&lo);
        o.y = *reinterpret_cast<uint32_t*>(&hi);
        reinterpret_cast<uint2*>(ad.shadow)[i] = o;
      }
    }
  }
}

}  // namespace

size_t peer_allreduce_pad(int world, size_t n) { return (n + size_t(4) * world - 1) / (size_t(4) * world) * (size_t(4) * world); }
size_t peer_allreduce_bytes(int world, size_t n) { return peer_allreduce_pad(world, n) * sizeof(float) + size_t(2) * world * sizeof(uint32_t); }

namespace {
template <bool ADAM>
int launch_allreduce(const unsigned long long* peer_base, size_t n, int world, int rank, uint32_t* state, unsigned long long timeout_ns, const FusedAdam& ad,
                     cudaStream_t st) {
  const size_t n_pad = peer_allreduce_pad(world, n);
  const size_t slice4 = n_pad / 4 / world;
  // Up to two CTAs per SM: ptxas keeps four of a thread's eight 16-byte loads in flight at a time, so the bytes in flight (2 x 148 x 256 x
  // 64 B = 4.8 MB) come from the CTA count.  ALL CTAs must be co-resident — the last one to finish its slice releases the exit flags the
  // others (all of them in the fused-Adam form) spin on — so the grid is bounded by what the occupancy calculator says fits at once.
  int per_sm = 0;
  const void* fn = nullptr;
  switch (world) {
    case 2: fn = reinterpret_cast<const void*>(&grad_allreduce_kernel<2, ADAM>); break;
    case 4: fn = reinterpret_cast<const void*>(&grad_allreduce_kernel<4, ADAM>); break;
    case 8: fn = reinterpret_cast<const void*>(&grad_allreduce_kernel<8, ADAM>); break;
    default: fn = reinterpret_cast<const void*>(&grad_allreduce_kernel<0, ADAM>); break;
  }
  AVF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, AR_THREADS, 0));
  AVF_REQUIRE(per_sm >= 1, AVF_EINVAL, "grad_allreduce: the kernel does not fit on an SM");
  const size_t want = ADAM ? std::max<size_t>((slice4 + AR_THREADS - 1) / AR_THREADS, (ad.n / 4 + AR_THREADS - 1) / AR_THREADS) : (slice4 + AR_THREADS - 1) / AR_THREADS;
  int grid = int(std::min<size_t>(size_t(std::min(per_sm, 2) * sm_count_of_current_device()), want));
  if (grid < 1) grid = 1;
  switch (world) {
    case 2: grad_allreduce_kernel<2, ADAM><<<grid, AR_THREADS, 0, st>>>(peer_base, n_pad, world, rank, state, timeout_ns, ad); break;
    case 4: grad_allreduce_kernel<4, ADAM><<<grid, AR_THREADS, 0, st>>>(peer_base, n_pad, world, rank, state, timeout_ns, ad); break;
    case 8: grad_allreduce_kernel<8, ADAM><<<grid, AR_THREADS, 0, st>>>(peer_base, n_pad, world, rank, state, timeout_ns, ad); break;
    default: grad_allreduce_kernel<0, ADAM><<<grid, AR_THREADS, 0, st>>>(peer_base, n_pad, world, rank, state, timeout_ns, ad); break;
  }
  AVF_LAUNCH_CHECK("grad_allreduce_kernel");
  return 0;
}
}  // namespace

int grad_allreduce(const unsigned long long* peer_base, size_t n, int world, int rank, uint32_t* state, unsigned long long timeout_ns, cudaStream_t st) {
  return launch_allreduce<false>(peer_base, n, world, rank, state, timeout_ns, FusedAdam{}, st);
}

// All-reduce (sum) of the peer-mapped gradient buckets + Adam / AdamW on the replicated parameters in ONE kernel; grad_scale = 1 / world
// (the mean of the data-parallel shards) unless the caller says otherwise.
int adam_allreduce_step(const unsigned long long* peer_base, size_t n, int world, int rank, uint32_t* state, unsigned long long timeout_ns, float* p, float* m,
                        float* v, void* shadow, float lr, float b1, float b2, float eps, float wd, int step, int decoupled, float grad_scale, cudaStream_t st) {
  AVF_REQUIRE(p && m && v, AVF_EINVAL, "adam_allreduce_step: null pointer");
  AVF_REQUIRE(step >= 1, AVF_EINVAL, "adam_allreduce_step: step=%d (1-based)", step);
  AVF_REQUIRE((n & 3u) == 0, AVF_EINVAL, "adam_allreduce_step: bucket length %zu is not a multiple of 4", n);
  AVF_REQUIRE(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(shadow)) & 15) == 0,
              AVF_EINVAL, "adam_allreduce_step: buckets must be 16-byte aligned");
  const double bc1 = 1.0 - pow(double(b1), step), bc2 = 1.0 - pow(double(b2), step);
  FusedAdam ad;
  ad.p = p; ad.m = m; ad.v = v; ad.shadow = static_cast<__nv_bfloat16*>(shadow); ad.n = n;
  ad.a = AdamParams{lr, b1, b2, eps, wd, float(1.0 / bc1), float(1.0 / sqrt(bc2)), grad_scale, decoupled};
  return launch_allreduce<true>(peer_base, n, world, rank, state, timeout_ns, ad, st);
}

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
