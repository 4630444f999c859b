// bf16 linear layer on the 5th-gen tensor cores:  C[M,N] = epi(A[M,K] * W[N,K]^T).
//
// Both operands are K-major (activations row-major, nn.Linear weights [out,in] row-major), so they go
// through TMA (128B swizzle, 64-column boxes) straight into the layout tcgen05.mma reads.  One CTA owns
// one 128 x BN output tile: warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2..5 = epilogue (TMEM -> registers -> bias / tanh-GELU / fp32 residual -> global).  The fp32
// accumulator never leaves TMEM until the epilogue.  Two CTAs fit per SM so one tile's epilogue
// overlaps the other's main loop.
//
// Replaces the cuBLAS calls behind nn.Linear at models/heads.py:192,195,212,215 of the reference.
#include <cuda.h>

#include <mutex>

#include "avf_common.cuh"

namespace avf {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;        // one 128-byte swizzle span of bf16
constexpr int UMMA_K = 16;

template <int BN, int STAGES>
struct GemmCfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int W_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + W_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int SMEM_BYTES = BAR_OFF + 256 + 1024;   // barriers + slack for 1024B alignment
  static constexpr int TMEM_COLS = 2 * BN;                  // double-buffered accumulator; power of two >= 32
};

// Epilogue math on one 32-column chunk held in registers, then the store of that chunk.
template <typename OutT>
__device__ __forceinline__ void epilogue_chunk(const uint32_t (&v)[32], int row, int col0, OutT* C, int ldc,
                                               const float* __restrict__ bias, const float* res, int ld_res, int flags) {
  float f[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
  if (flags & AVF_EPI_BIAS) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(bias + col0 + j));
      f[j] += b.x; f[j + 1] += b.y; f[j + 2] += b.z; f[j + 3] += b.w;
    }
  }
  if (flags & AVF_EPI_GELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = gelu_tanh<true>(f[j]);
  }
  if (flags & AVF_EPI_RESIDUAL) {
    const float* rp = res + size_t(row) * ld_res + col0;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 r = *reinterpret_cast<const float4*>(rp + j);
      f[j] += r.x; f[j + 1] += r.y; f[j + 2] += r.z; f[j + 3] += r.w;
    }
  }
  OutT* cp = C + size_t(row) * ldc + col0;
  if constexpr (sizeof(OutT) == 4) {
#pragma unroll
    for (int j = 0; j < 32; j += 4)
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(cp) + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
  } else {
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      uint4 o;
      o.x = pack_bf16x2(f[j], f[j + 1]);
      o.y = pack_bf16x2(f[j + 2], f[j + 3]);
      o.z = pack_bf16x2(f[j + 4], f[j + 5]);
      o.w = pack_bf16x2(f[j + 6], f[j + 7]);
      *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(cp) + j) = o;
    }
  }
}

// Persistent kernel: grid = min(#tiles, #SMs); every CTA walks tiles t = blockIdx.x, +gridDim.x, ...
// (n-block fastest, so CTAs that run concurrently share the A tile in L2).  The TMEM accumulator is double
// buffered (2 x BN columns): the MMA warp fills buffer (t+1)&1 while the 8 epilogue warps drain buffer t&1.
template <int BN, int STAGES, typename OutT>
__global__ void __launch_bounds__(384, 1)
gemm_umma_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w,
                 OutT* C, int ldc, const float* __restrict__ bias,
                 const float* res, int ld_res, int M, int N, int K, int flags) {
  using Cfg = GemmCfg<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFF);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* acc_full = empty_bar + STAGES;       // [2]
  uint64_t* acc_empty = acc_full + 2;            // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kblocks = K / BK;
  const int n_blocks = N / BN;
  const int n_tiles = n_blocks * ((M + BM - 1) / BM);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_w);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], 8);               // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int m0 = (tile / n_blocks) * BM, n0 = (tile % n_blocks) * BN;
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          mbar_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
          uint8_t* st = smem + s * Cfg::STAGE_BYTES;
          tma_load_2d(st, &tm_a, &full_bar[s], kb * BK, m0);
          tma_load_2d(st + Cfg::A_BYTES, &tm_w, &full_bar[s], kb * BK, n0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN);
      uint32_t it = 0, lt = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++lt) {
        const uint32_t buf = lt & 1, aph = (lt >> 1) & 1;
        mbar_wait(&acc_empty[buf], aph ^ 1);      // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + buf * BN;
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * Cfg::STAGE_BYTES);
          const uint64_t da = make_desc_sw128_kmajor(a_addr);
          const uint64_t db = make_desc_sw128_kmajor(a_addr + Cfg::A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advance along K inside the 128B swizzle span: +32 bytes per UMMA_K (>>4 in descriptor units)
            umma_bf16(tmem_d, da + uint64_t(k * 2), db + uint64_t(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);             // frees the smem stage when these MMAs have read it
        }
        umma_commit(&acc_full[buf]);              // accumulator complete
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int half = (warp - 4) >> 2;             // which half of the tile's columns
    constexpr int HALF_COLS = BN / 2;
    uint32_t lt = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++lt) {
      const int m0 = (tile / n_blocks) * BM, n0 = (tile % n_blocks) * BN;
      const uint32_t buf = lt & 1, aph = (lt >> 1) & 1;
      const int row = m0 + q * 32 + lane;
      mbar_wait(&acc_full[buf], aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + buf * BN + half * HALF_COLS;
#pragma unroll 1
      for (int c0 = 0; c0 < HALF_COLS; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + uint32_t(c0), v);
        tmem_ld_wait();
        if (c0 + 32 >= HALF_COLS) {               // last chunk is in registers: hand the buffer back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[buf]);
        }
        if (row < M) epilogue_chunk<OutT>(v, row, n0 + half * HALF_COLS + c0, C, ldc, bias, res, ld_res, flags);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------
// host: tensor maps
// ---------------------------------------------------------------------------------------------
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

}  // namespace

// 2-D bf16 row-major [rows, cols] (row stride ld elements), box = [box_rows x 64 cols], 128B swizzle.
int make_tmap_bf16_2d(CUtensorMap* tm, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  AVF_REQUIRE(fn != nullptr, AVF_ENODEVICE, "cuTensorMapEncodeTiled is not available from the driver");
  AVF_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (ld * 2) % 16 == 0, AVF_EINVAL,
              "TMA operand must be 16-byte aligned with a 16-byte multiple row pitch (ptr=%p ld=%llu)", ptr,
              (unsigned long long)ld);
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  AVF_REQUIRE(r == CUDA_SUCCESS, AVF_EINVAL, "cuTensorMapEncodeTiled failed with CUresult %d", int(r));
  return 0;
}

static int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

template <int BN, int STAGES, typename OutT>
static int launch_gemm(const void* a, int lda, const void* w, OutT* c, int ldc, const float* bias,
                       const float* res, int ld_res, int m, int n, int k, int flags, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, STAGES>;
  CUtensorMap ta, tw;
  int e = make_tmap_bf16_2d(&ta, a, m, k, lda, BM);
  if (e) return e;
  e = make_tmap_bf16_2d(&tw, w, n, k, k, BN);
  if (e) return e;
  auto kern = gemm_umma_kernel<BN, STAGES, OutT>;
  static bool configured = false;
  if (!configured) {
    AVF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured = true;
  }
  const int n_tiles = (n / BN) * ceil_div(m, BM);
  kern<<<min(n_tiles, sm_count()), 384, Cfg::SMEM_BYTES, stream>>>(ta, tw, c, ldc, bias, res, ld_res, m, n, k, flags);
  AVF_LAUNCH_CHECK("gemm_umma_kernel");
  return 0;
}

// C = epi(A W^T) with bf16 operands; c_mode selects fp32 / bf16 output.
int linear_umma(const void* a, int lda, const void* w, const float* bias, const float* res, int ld_res, void* c,
                int ldc, int c_mode, int m, int n, int k, int flags, cudaStream_t stream) {
  AVF_REQUIRE(m > 0 && n > 0 && k > 0, AVF_EINVAL, "linear: empty problem m=%d n=%d k=%d", m, n, k);
  AVF_REQUIRE(k % BK == 0, AVF_EUNSUPPORTED, "linear(bf16): K=%d must be a multiple of %d", k, BK);
  AVF_REQUIRE(n % 64 == 0, AVF_EUNSUPPORTED, "linear(bf16): N=%d must be a multiple of 64", n);
  AVF_REQUIRE(ldc % 8 == 0 && (!(flags & AVF_EPI_RESIDUAL) || ld_res % 4 == 0), AVF_EINVAL,
              "linear(bf16): ldc=%d / ld_res=%d break vector alignment", ldc, ld_res);
  // widest tile that still gives every SM at least two tiles; small problems get narrower tiles for parallelism
  const int m_tiles = ceil_div(m, BM), sms = sm_count();
  int bn = 64;
  if (n % 256 == 0 && m_tiles * (n / 256) >= 2 * sms) bn = 256;
  else if (n % 128 == 0 && m_tiles * (n / 128) >= sms) bn = 128;
  else if (n % 128 == 0 && n / 128 * m_tiles >= sms / 2 && n >= 512) bn = 128;
#define AVF_GEMM(BN_, ST_)                                                                                                         \
  if (bn == BN_) {                                                                                                                 \
    if (c_mode == AVF_BF16)                                                                                                        \
      return launch_gemm<BN_, ST_, __nv_bfloat16>(a, lda, w, static_cast<__nv_bfloat16*>(c), ldc, bias, res, ld_res, m, n, k, flags, stream); \
    return launch_gemm<BN_, ST_, float>(a, lda, w, static_cast<float*>(c), ldc, bias, res, ld_res, m, n, k, flags, stream);          \
  }
  AVF_GEMM(256, 4)
  AVF_GEMM(128, 6)
  AVF_GEMM(64, 8)
#undef AVF_GEMM
  return AVF_EUNSUPPORTED;
}

}  // namespace avf
