// Fused pre-LN encoder stack for dim 256 / 8 heads x 32 (SFormer, fusion head): ONE persistent kernel runs
//   x + to_out(softmax(q k^T / sqrt(dh)) v)   and   x + W2 gelu(W1 LN(x) + b1) + b2        (models/heads.py:164-256)
// for `depth` layers on 128-row tiles (floor(128 / n_tok) whole sequences per tile) without touching HBM in between.
//
//   * the fp32 residual stream of the tile lives in TMEM columns [0,256): the out-projection and the second MLP GEMM
//     ACCUMULATE onto it (tcgen05.mma with accumulate=1 on a tile pre-loaded with x + bias), so both residual adds
//     and both bias adds cost nothing;
//   * LayerNorm runs on the row workers (one TMEM lane = one token row, two threads per row) and writes the bf16
//     A operand straight into the 128B-swizzled K-major layout tcgen05.mma reads;
//   * attention is per head on the tensor cores: [Q|K|V]_h = LN(x) Wqkv_h^T (N=96), S = Q K^T over the whole tile
//     (block-diagonal: a row only uses the columns of its own sequence), softmax in registers, P written back as
//     bf16 INTO the S columns of TMEM and used as the A operand of O = P V (V is an MN-major B operand);
//   * weights are streamed from L2 by TMA through a 4-slot ring, activations enter / leave as whole NCHW frames
//     by bulk copies (SFormer, models/vformer.py:245-259) or as fp32 rows.
//
// Warp roles: 0 = TMA producer (weights; also the next tile's input frames), 1 = TMEM allocator + MMA issuer (one thread),
// 2..9 = row workers (warp w owns TMEM lanes 32*(w%4)..+31; warps 2-5 / 6-9 split the columns).  320 threads -> 200 registers each.
#include <cuda.h>

#include "avf_common.cuh"
#include "avf_internal.h"

namespace avf {

int make_tmap_bf16_2d(CUtensorMap* tm, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows);

namespace {

constexpr int DIM = 256, HEADS = 8, DH = 32;
constexpr int MAX_DEPTH = 3;
constexpr int NUM_THREADS = 320;
constexpr int WORKER_T0 = 64;          // first worker thread
constexpr int RING = 4, SLOT_BYTES = 16384;

// shared memory map (offsets from a 1024-byte aligned base)
constexpr int OFF_A0 = 0;                          // 64 KB  LN output [128 x 256] bf16, 4 K-panels; also the NCHW input staging
constexpr int OFF_A1 = 65536;                      // 64 KB  attention output / 2 x GELU(hidden chunk) / NCHW output staging
constexpr int OFF_Q = 131072;                      // 8 KB   Q_h [128 x 32] K-major SW64
constexpr int OFF_K = OFF_Q + 8192;                // 8 KB   K_h
constexpr int OFF_V = OFF_K + 8192;                // 2x8 KB V_h [128 tok x 32] MN-major SW64, double buffered
constexpr int OFF_RING = OFF_V + 16384;            // 4 x 16 KB weight ring
constexpr int OFF_XCH = OFF_RING + RING * SLOT_BYTES;   // 2 KB  row-pair exchange [2][128][2] floats
constexpr int OFF_BAR = OFF_XCH + 2048;
constexpr int SMEM_USED = OFF_BAR + 256;           // 231,680
constexpr int SMEM_ALLOC = 232448;                 // 227 KB: everything the SM has

// TMEM columns
constexpr uint32_t TM_X = 0, TM_D1 = 256, TM_O = 352, TM_S = 384, TM_H0 = 256, TM_H1 = 384;

enum {
  B_RING_FULL = 0, B_RING_EMPTY = 4, B_X0_FULL = 8, B_A0_FREE, B_A0_READY, B_D1_FULL, B_STAGED, B_S_FULL, B_P_READY, B_O_FULL,
  B_O_DRAINED, B_X1_FULL, B_HACC_FULL, B_HACC_FULL1, B_H_READY, B_H_READY1, B_HBUF_FREE, B_HBUF_FREE1, B_X2_FULL, NUM_BARS
};

enum { IO_NCHW_BF16 = 0, IO_ROWS_F32 = 1 };

struct LayerArgs {
  CUtensorMap tm_qkv, tm_out, tm_w1, tm_w2;
  const float *ln1_g, *ln1_b, *b_out, *ln2_g, *ln2_b, *b_ff1, *b_ff2;
  uint64_t pad_;
};

struct FusedArgs {
  LayerArgs layer[MAX_DEPTH];
  const void* in;
  void* out;
  const float* pos;          // [n_tok, 256] or nullptr
  int ld_in, ld_out;         // IO_ROWS_F32 row strides (elements)
  int n_seq, n_tok, spt, n_tiles, n_chunks, depth;
};

__device__ __forceinline__ void bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
               "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void bulk_store_1d(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(reinterpret_cast<uint64_t>(gdst)), "r"(smem_u32(smem_src)),
               "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ uint64_t desc_sw64(uint32_t addr) { return make_desc(addr, 16, 512, 4); }

// ---------------------------------------------------------------------------------------------
// row workers
// ---------------------------------------------------------------------------------------------
struct Worker {
  uint8_t* smem;
  uint64_t* bars;
  uint32_t tl;          // TMEM address of this warp's lane quarter, column 0
  int lane, q, g, row;
  uint32_t xslot;

  __device__ __forceinline__ float exchange(float mine) {      // value of the other thread that owns this row
    float* s = reinterpret_cast<float*>(smem + OFF_XCH) + xslot * 256;
    xslot ^= 1;
    s[row * 2 + g] = mine;
    bar_sync(1 + q, 64);
    return s[row * 2 + (g ^ 1)];
  }
  __device__ __forceinline__ void arrive(int bar) {             // one arrive per warp, after every lane's fences
    fence_proxy_async_smem();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&bars[bar]);
  }
  // Statistics are accumulated on chunks as shifted sums (shift = the thread's first element) and merged with the
  // partner thread by Chan's formula, so one sweep over the row is enough and nothing cancels catastrophically.
  struct Stats {
    float shift, s, ss;
    bool have;
    __device__ __forceinline__ void add(const float (&x)[32]) {
      if (!have) {
        shift = x[0];
        have = true;
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float d = x[j] - shift;
        s += d;
        ss = fmaf(d, d, ss);
      }
    }
  };
  __device__ __forceinline__ void finish_stats(const Stats& st, float& mean, float& rstd) {
    const float mean_g = st.shift + st.s * (1.f / 128.f);
    const float m2_g = fmaxf(st.ss - st.s * st.s * (1.f / 128.f), 0.f);
    const float mean_o = exchange(mean_g), m2_o = exchange(m2_g);
    const float dm = mean_g - mean_o;
    mean = 0.5f * (mean_g + mean_o);
    rstd = rsqrtf((m2_g + m2_o + dm * dm * 64.f) * (1.f / DIM) + 1e-5f);
  }
  // Sweep 1 of a LayerNorm whose input already sits in TMEM (x1 after the out-projection, x2 after the MLP).
  __device__ __forceinline__ void stats_from_tmem(float& mean, float& rstd) {
    Stats st{0.f, 0.f, 0.f, false};
#pragma unroll 1
    for (int c0 = 0; c0 < 128; c0 += 32) {
      float x[32];
      tmem_ld32(tl + TM_X + g * 128 + c0, reinterpret_cast<uint32_t(&)[32]>(x));
      tmem_ld_wait();
      st.add(x);
    }
    finish_stats(st, mean, rstd);
  }
  // Sweep 2: LN(x) -> A0 (bf16, K-major SW128 panels), x + next_bias -> TMEM; then signal the MMA thread.
  __device__ __forceinline__ void normalize_from_tmem(float mean, float rstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                      const float* __restrict__ next_bias) {
    const int cbase = g * 128;
#pragma unroll 1
    for (int c0 = 0; c0 < 128; c0 += 32) {
      float x[32];
      tmem_ld32(tl + TM_X + cbase + c0, reinterpret_cast<uint32_t(&)[32]>(x));
      tmem_ld_wait();
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        const int col = cbase + c0 + ch * 8;
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + col)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + col + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + col)), b1 = __ldg(reinterpret_cast<const float4*>(beta + col + 4));
        const float* y = &x[ch * 8];
        uint4 pk;
        pk.x = pack_bf16x2(fmaf((y[0] - mean) * rstd, g0.x, b0.x), fmaf((y[1] - mean) * rstd, g0.y, b0.y));
        pk.y = pack_bf16x2(fmaf((y[2] - mean) * rstd, g0.z, b0.z), fmaf((y[3] - mean) * rstd, g0.w, b0.w));
        pk.z = pack_bf16x2(fmaf((y[4] - mean) * rstd, g1.x, b1.x), fmaf((y[5] - mean) * rstd, g1.y, b1.y));
        pk.w = pack_bf16x2(fmaf((y[6] - mean) * rstd, g1.z, b1.z), fmaf((y[7] - mean) * rstd, g1.w, b1.w));
        const int panel = col >> 6, chunk = (col & 63) >> 3;
        *reinterpret_cast<uint4*>(smem + OFF_A0 + panel * 16384 + row * 128 + ((chunk ^ (row & 7)) << 4)) = pk;
      }
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(next_bias + cbase + c0 + j));
        x[j] += b.x; x[j + 1] += b.y; x[j + 2] += b.z; x[j + 3] += b.w;
      }
      tmem_st32(tl + TM_X + cbase + c0, reinterpret_cast<const uint32_t(&)[32]>(x));
    }
    tmem_st_wait();
    arrive(B_A0_READY);
  }
};

template <int IO>
__device__ void worker_main(const FusedArgs& a, uint8_t* smem, uint64_t* bars, uint32_t tmem) {
  Worker w;
  w.smem = smem;
  w.bars = bars;
  const int wt = threadIdx.x - WORKER_T0;
  w.lane = wt & 31;
  w.q = (threadIdx.x >> 5) & 3;          // TMEM lane quarter = warp index mod 4
  w.g = wt >> 7;
  w.row = w.q * 32 + w.lane;
  w.tl = tmem + (uint32_t(w.q * 32) << 16);
  w.xslot = 0;
  const int q = w.q, g = w.g, row = w.row;
  const uint32_t tl = w.tl;
  const int n_tok = a.n_tok, spt = a.spt, rows_full = spt * n_tok;
  const int kblocks = (rows_full + 15) >> 4;                       // 16-column blocks of S / P in use
  // softmax geometry: this row attends to columns [lo, hi); this warp's rows need blocks [b_lo, b_hi)
  const int sq = min(row / n_tok, spt - 1);
  const int lo = sq * n_tok, hi = lo + n_tok;
  const int t_in_seq = row - (row / n_tok) * n_tok;
  const int r0 = q * 32, r1 = min(q * 32 + 31, rows_full - 1);
  const int lo_w = min(r0 / n_tok, spt - 1) * n_tok, hi_w = (min(r1 / n_tok, spt - 1) + 1) * n_tok;
  const int b_lo = lo_w >> 4, b_hi = (hi_w + 15) >> 4;
  const int nb0 = (b_hi - b_lo + 1) >> 1;
  const int my_b0 = g == 0 ? b_lo : b_lo + nb0;
  const int my_nb = g == 0 ? nb0 : (b_hi - b_lo) - nb0;             // <= 4
  const float sm_scale = 1.4426950408889634f * rsqrtf(float(DH));
  uint32_t n_x0 = 0, n_x1 = 0, n_x2 = 0, n_hacc[2] = {0, 0}, n_hfree[2] = {0, 0};

  for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    const int seqs_here = min(spt, a.n_seq - tile * spt);
    const int rows_here = seqs_here * n_tok;
    const size_t grow = size_t(tile) * spt * n_tok + row;           // global token row (IO_ROWS_F32)
    // ---- tile input: x (+ pos) -> TMEM, statistics of LN1 on the way --------------------------------------
    float mean, rstd;
    {
      Worker::Stats st{0.f, 0.f, 0.f, false};
      const bool valid = row < rows_here;
      if constexpr (IO == IO_NCHW_BF16) mbar_wait(&bars[B_X0_FULL], (n_x0++) & 1);
      const int fr = row / n_tok;
      const __nv_bfloat16* src16 = reinterpret_cast<const __nv_bfloat16*>(smem + OFF_A0) + size_t(fr * DIM + g * 128) * n_tok + t_in_seq;
      const float* src32 = static_cast<const float*>(a.in) + grow * a.ld_in + g * 128;
      const float* pp = a.pos + t_in_seq * DIM + g * 128;
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 32) {
        float x[32];
        if (valid) {
          if constexpr (IO == IO_NCHW_BF16) {
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = __bfloat162float(src16[(c0 + j) * n_tok]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 x4 = *reinterpret_cast<const float4*>(src32 + c0 + j);
              x[j] = x4.x; x[j + 1] = x4.y; x[j + 2] = x4.z; x[j + 3] = x4.w;
            }
          }
          if (a.pos != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 p4 = __ldg(reinterpret_cast<const float4*>(pp + c0 + j));
              x[j] += p4.x; x[j + 1] += p4.y; x[j + 2] += p4.z; x[j + 3] += p4.w;
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] = 0.f;
        }
        st.add(x);
        tmem_st32(tl + TM_X + g * 128 + c0, reinterpret_cast<const uint32_t(&)[32]>(x));
      }
      tmem_st_wait();
      w.finish_stats(st, mean, rstd);
      if constexpr (IO == IO_NCHW_BF16) bar_sync(5, 256);   // every worker has read its part of the staged frames out of A0
    }

    for (int l = 0; l < a.depth; ++l) {
      const LayerArgs& L = a.layer[l];
      w.normalize_from_tmem(mean, rstd, L.ln1_g, L.ln1_b, L.b_out);

      // ---- attention: per head  E1 (QKV -> smem), E3 of the previous head (O -> A1), E2 (softmax) -------------
      float inv_l = 0.f;
#pragma unroll 1
      for (int h = 0; h <= HEADS; ++h) {
        if (h < HEADS) {
          mbar_wait(&bars[B_D1_FULL], h & 1);
          tc_fence_after();
          const uint32_t sw = uint32_t((row >> 1) & 3);
          uint8_t* vbuf = smem + OFF_V + (h & 1) * 8192 + row * 64;
          if (g == 0) {          // Q (cols 0..31) and the first half of K (cols 32..47)
            uint32_t r[32], r2[16];
            tmem_ld32(tl + TM_D1, r);
            tmem_ld16(tl + TM_D1 + 32, r2);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              uint4 pk;
              pk.x = pack_bf16x2(__uint_as_float(r[c * 8 + 0]), __uint_as_float(r[c * 8 + 1]));
              pk.y = pack_bf16x2(__uint_as_float(r[c * 8 + 2]), __uint_as_float(r[c * 8 + 3]));
              pk.z = pack_bf16x2(__uint_as_float(r[c * 8 + 4]), __uint_as_float(r[c * 8 + 5]));
              pk.w = pack_bf16x2(__uint_as_float(r[c * 8 + 6]), __uint_as_float(r[c * 8 + 7]));
              *reinterpret_cast<uint4*>(smem + OFF_Q + row * 64 + ((uint32_t(c) ^ sw) << 4)) = pk;
            }
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              uint4 pk;
              pk.x = pack_bf16x2(__uint_as_float(r2[c * 8 + 0]), __uint_as_float(r2[c * 8 + 1]));
              pk.y = pack_bf16x2(__uint_as_float(r2[c * 8 + 2]), __uint_as_float(r2[c * 8 + 3]));
              pk.z = pack_bf16x2(__uint_as_float(r2[c * 8 + 4]), __uint_as_float(r2[c * 8 + 5]));
              pk.w = pack_bf16x2(__uint_as_float(r2[c * 8 + 6]), __uint_as_float(r2[c * 8 + 7]));
              *reinterpret_cast<uint4*>(smem + OFF_K + row * 64 + ((uint32_t(c) ^ sw) << 4)) = pk;
            }
          } else {               // second half of K (cols 48..63) and V (cols 64..95)
            uint32_t r[32], r2[16];
            tmem_ld16(tl + TM_D1 + 48, r2);
            tmem_ld32(tl + TM_D1 + 64, r);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              uint4 pk;
              pk.x = pack_bf16x2(__uint_as_float(r2[c * 8 + 0]), __uint_as_float(r2[c * 8 + 1]));
              pk.y = pack_bf16x2(__uint_as_float(r2[c * 8 + 2]), __uint_as_float(r2[c * 8 + 3]));
              pk.z = pack_bf16x2(__uint_as_float(r2[c * 8 + 4]), __uint_as_float(r2[c * 8 + 5]));
              pk.w = pack_bf16x2(__uint_as_float(r2[c * 8 + 6]), __uint_as_float(r2[c * 8 + 7]));
              *reinterpret_cast<uint4*>(smem + OFF_K + row * 64 + ((uint32_t(c + 2) ^ sw) << 4)) = pk;
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              uint4 pk;
              pk.x = pack_bf16x2(__uint_as_float(r[c * 8 + 0]), __uint_as_float(r[c * 8 + 1]));
              pk.y = pack_bf16x2(__uint_as_float(r[c * 8 + 2]), __uint_as_float(r[c * 8 + 3]));
              pk.z = pack_bf16x2(__uint_as_float(r[c * 8 + 4]), __uint_as_float(r[c * 8 + 5]));
              pk.w = pack_bf16x2(__uint_as_float(r[c * 8 + 6]), __uint_as_float(r[c * 8 + 7]));
              *reinterpret_cast<uint4*>(vbuf + ((uint32_t(c) ^ sw) << 4)) = pk;
            }
          }
          w.arrive(B_STAGED);
        }
        if (h > 0) {             // E3 of head h-1: O / l -> bf16 -> A1 columns [(h-1)*32, +32)
          mbar_wait(&bars[B_O_FULL], (h - 1) & 1);
          tc_fence_after();
          uint32_t r[16];
          tmem_ld16(tl + TM_O + g * 16, r);
          tmem_ld_wait();
          const int col = (h - 1) * DH + g * 16;
          const int panel = col >> 6, chunk = (col & 63) >> 3;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint4 pk;
            pk.x = pack_bf16x2(__uint_as_float(r[c * 8 + 0]) * inv_l, __uint_as_float(r[c * 8 + 1]) * inv_l);
            pk.y = pack_bf16x2(__uint_as_float(r[c * 8 + 2]) * inv_l, __uint_as_float(r[c * 8 + 3]) * inv_l);
            pk.z = pack_bf16x2(__uint_as_float(r[c * 8 + 4]) * inv_l, __uint_as_float(r[c * 8 + 5]) * inv_l);
            pk.w = pack_bf16x2(__uint_as_float(r[c * 8 + 6]) * inv_l, __uint_as_float(r[c * 8 + 7]) * inv_l);
            *reinterpret_cast<uint4*>(smem + OFF_A1 + panel * 16384 + row * 128 + (((chunk + c) ^ (row & 7)) << 4)) = pk;
          }
          w.arrive(B_O_DRAINED);
        }
        if (h < HEADS) {         // E2: masked softmax of this row over its own sequence, P (bf16) over the S columns
          mbar_wait(&bars[B_S_FULL], h & 1);
          tc_fence_after();
          float s[4][16];
#pragma unroll
          for (int b = 0; b < 4; ++b)
            if (b < my_nb) tmem_ld16(tl + TM_S + (my_b0 + b) * 16, reinterpret_cast<uint32_t(&)[16]>(s[b]));
          tmem_ld_wait();
          float mx = -INFINITY;
#pragma unroll
          for (int b = 0; b < 4; ++b) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int col = (my_b0 + b) * 16 + j;
              const bool ok = (b < my_nb) && col >= lo && col < hi;
              s[b][j] = ok ? s[b][j] * sm_scale : -INFINITY;
              mx = fmaxf(mx, s[b][j]);
            }
          }
          mx = fmaxf(mx, w.exchange(mx));      // also orders every S load of the row pair before any P store
          float sum = 0.f;
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            if (b < my_nb) {
              uint32_t pk[8];
#pragma unroll
              for (int j = 0; j < 16; j += 2) {
                const float p0 = fast_exp2(s[b][j] - mx), p1 = fast_exp2(s[b][j + 1] - mx);
                sum += p0 + p1;
                pk[j >> 1] = pack_bf16x2(p0, p1);
              }
              tmem_st8(tl + TM_S + (my_b0 + b) * 8, pk);
            }
          }
          {                      // P columns the MMA reads but no row of this warp uses: zeros
            const uint32_t z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            const int zb0 = g == 0 ? 0 : b_hi, zb1 = g == 0 ? b_lo : kblocks;
            for (int b = zb0; b < zb1; ++b) tmem_st8(tl + TM_S + b * 8, z);
          }
          sum += w.exchange(sum);
          inv_l = 1.f / sum;
          tmem_st_wait();
          w.arrive(B_P_READY);
        }
      }

      // ---- LN2 on x1 = x + attention (accumulated in TMEM by the out-projection) ------------------------------
      mbar_wait(&bars[B_X1_FULL], (n_x1++) & 1);
      tc_fence_after();
      w.stats_from_tmem(mean, rstd);
      w.normalize_from_tmem(mean, rstd, L.ln2_g, L.ln2_b, L.b_ff2);

      // ---- MLP: per 128-column chunk of the hidden layer, bias + tanh-GELU -> bf16 A operand ---------------------
#pragma unroll 1
      for (int c = 0; c < a.n_chunks; ++c) {
        const int b = c & 1;
        mbar_wait(&bars[B_HACC_FULL + b], (n_hacc[b]++) & 1);
        if (c >= 2) mbar_wait(&bars[B_HBUF_FREE + b], (n_hfree[b]++) & 1);
        tc_fence_after();
        const float* bias = L.b_ff1 + c * 128 + g * 64;
        uint8_t* dst = smem + OFF_A1 + b * 32768 + g * 16384 + row * 128;
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(tl + (b ? TM_H1 : TM_H0) + g * 64 + c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + c0 + ch * 8)), b1 = __ldg(reinterpret_cast<const float4*>(bias + c0 + ch * 8 + 4));
            uint4 pk;
            pk.x = pack_bf16x2(gelu_tanh<true>(__uint_as_float(r[ch * 8 + 0]) + b0.x), gelu_tanh<true>(__uint_as_float(r[ch * 8 + 1]) + b0.y));
            pk.y = pack_bf16x2(gelu_tanh<true>(__uint_as_float(r[ch * 8 + 2]) + b0.z), gelu_tanh<true>(__uint_as_float(r[ch * 8 + 3]) + b0.w));
            pk.z = pack_bf16x2(gelu_tanh<true>(__uint_as_float(r[ch * 8 + 4]) + b1.x), gelu_tanh<true>(__uint_as_float(r[ch * 8 + 5]) + b1.y));
            pk.w = pack_bf16x2(gelu_tanh<true>(__uint_as_float(r[ch * 8 + 6]) + b1.z), gelu_tanh<true>(__uint_as_float(r[ch * 8 + 7]) + b1.w));
            *reinterpret_cast<uint4*>(dst + (((c0 >> 3) + ch) ^ (row & 7)) * 16) = pk;
          }
        }
        w.arrive(B_H_READY + b);
      }

      // ---- x2 = x1 + MLP, accumulated in TMEM by the second MLP GEMM --------------------------------------------
      mbar_wait(&bars[B_X2_FULL], (n_x2++) & 1);
      tc_fence_after();
      if (l + 1 < a.depth) w.stats_from_tmem(mean, rstd);
    }

    // ---- tile output ----------------------------------------------------------------------------
    {
      const bool valid = row < rows_here;
      const int fr = row / n_tok;
      __nv_bfloat16* dst16 = reinterpret_cast<__nv_bfloat16*>(smem + OFF_A1) + size_t(fr * DIM + g * 128) * n_tok + t_in_seq;
      float* dst32 = static_cast<float*>(a.out) + grow * a.ld_out + g * 128;
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 32) {
        float x[32];
        tmem_ld32(tl + TM_X + g * 128 + c0, reinterpret_cast<uint32_t(&)[32]>(x));
        tmem_ld_wait();
        if (valid) {
          if constexpr (IO == IO_NCHW_BF16) {
#pragma unroll
            for (int j = 0; j < 32; ++j) dst16[(c0 + j) * n_tok] = __float2bfloat16_rn(x[j]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst32 + c0 + j) = make_float4(x[j], x[j + 1], x[j + 2], x[j + 3]);
          }
        }
      }
      if constexpr (IO == IO_NCHW_BF16) {
        fence_proxy_async_smem();
        bar_sync(5, 256);
        if (threadIdx.x == WORKER_T0) {
          __nv_bfloat16* gdst = static_cast<__nv_bfloat16*>(a.out) + size_t(tile) * spt * n_tok * DIM;
          bulk_store_1d(gdst, smem + OFF_A1, uint32_t(rows_here) * DIM * 2);
          bulk_wait_read0();     // A1 is written again by the next tile's attention (ordered by the bar_sync after its input sweep)
        }
      }
    }
  }
  if constexpr (IO == IO_NCHW_BF16) {
    if (threadIdx.x == WORKER_T0) bulk_wait_all0();
  }
}

// ---------------------------------------------------------------------------------------------
// weight producer (one thread): must issue slots in exactly the order the MMA thread consumes them
// ---------------------------------------------------------------------------------------------
template <int IO>
__device__ void producer_main(const FusedArgs& a, uint8_t* smem, uint64_t* bars) {
  uint32_t it = 0;
  // Input frames of the NEXT tile go into A0 as soon as the MMAs of the current tile are done with it (B_A0_FREE, one
  // completion per tile).  Polled between weight slots so that this thread never blocks on it.
  int load_tile = blockIdx.x;
  uint32_t n_free = 0;
  bool need_free = false;
  auto poll_loader = [&]() {
    if (IO != IO_NCHW_BF16 || load_tile >= a.n_tiles) return;
    if (need_free) {
      if (!mbar_try_wait(&bars[B_A0_FREE], n_free & 1)) return;
      ++n_free;
    }
    const int seqs_here = min(a.spt, a.n_seq - load_tile * a.spt);
    const uint32_t bytes = uint32_t(seqs_here) * a.n_tok * DIM * 2;
    mbar_expect_tx(&bars[B_X0_FULL], bytes);
    bulk_load_1d(smem + OFF_A0, static_cast<const __nv_bfloat16*>(a.in) + size_t(load_tile) * a.spt * a.n_tok * DIM, bytes, &bars[B_X0_FULL]);
    load_tile += gridDim.x;
    need_free = true;
  };
  auto slot = [&](uint32_t bytes) -> uint8_t* {
    const uint32_t s = it % RING, ph = (it / RING) & 1;
    poll_loader();
    const long long t0 = clock64();
    while (!mbar_try_wait(&bars[B_RING_EMPTY + s], ph ^ 1)) {
      poll_loader();
      if (clock64() - t0 > 4000000000LL) {
        printf("avf: fused encoder producer timed out (block %d slot %u)\n", blockIdx.x, it);
        __trap();
      }
    }
    mbar_expect_tx(&bars[B_RING_FULL + s], bytes);
    return smem + OFF_RING + s * SLOT_BYTES;
  };
  for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    for (int l = 0; l < a.depth; ++l) {
      const LayerArgs& L = a.layer[l];
      for (int h = 0; h < HEADS; ++h)
        for (int kp = 0; kp < 4; ++kp) {
          uint8_t* d = slot(3 * 32 * 128);
          uint64_t* fb = &bars[B_RING_FULL + it % RING];
          for (int s3 = 0; s3 < 3; ++s3) tma_load_2d(d + s3 * 4096, &L.tm_qkv, fb, kp * 64, s3 * (HEADS * DH) + h * DH);
          ++it;
        }
      for (int kp = 0; kp < 4; ++kp)
        for (int nh = 0; nh < 2; ++nh) {
          uint8_t* d = slot(SLOT_BYTES);
          tma_load_2d(d, &L.tm_out, &bars[B_RING_FULL + it % RING], kp * 64, nh * 128);
          ++it;
        }
      auto ff1 = [&](int c) {
        for (int kp = 0; kp < 4; ++kp) {
          uint8_t* d = slot(SLOT_BYTES);
          tma_load_2d(d, &L.tm_w1, &bars[B_RING_FULL + it % RING], kp * 64, c * 128);
          ++it;
        }
      };
      auto ff2 = [&](int c) {
        for (int kp = 0; kp < 2; ++kp)
          for (int nh = 0; nh < 2; ++nh) {
            uint8_t* d = slot(SLOT_BYTES);
            tma_load_2d(d, &L.tm_w2, &bars[B_RING_FULL + it % RING], c * 128 + kp * 64, nh * 128);
            ++it;
          }
      };
      ff1(0);
      if (a.n_chunks > 1) ff1(1);
      for (int c = 0; c < a.n_chunks; ++c) {
        ff2(c);
        if (c + 2 < a.n_chunks) ff1(c + 2);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// MMA issuer (one thread)
// ---------------------------------------------------------------------------------------------
__device__ void mma_main(const FusedArgs& a, uint8_t* smem, uint64_t* bars, uint32_t tmem) {
  const uint32_t a0 = smem_u32(smem + OFF_A0), a1 = smem_u32(smem + OFF_A1);
  const uint32_t qs = smem_u32(smem + OFF_Q), ks = smem_u32(smem + OFF_K), vs = smem_u32(smem + OFF_V);
  const uint32_t ring = smem_u32(smem + OFF_RING);
  const int rows_full = a.spt * a.n_tok;
  const int kmax = ((rows_full + 15) >> 4) << 4;
  const uint32_t id_qkv = make_idesc_bf16(128, 96), id_s = make_idesc_bf16(128, kmax), id_pv = make_idesc_bf16(128, DH, 0, 1),
                 id_128 = make_idesc_bf16(128, 128);
  uint32_t it = 0, n_a0 = 0, n_hready[2] = {0, 0};
  auto slot_wait = [&]() -> uint32_t {
    const uint32_t s = it % RING, ph = (it / RING) & 1;
    mbar_wait(&bars[B_RING_FULL + s], ph);
    tc_fence_after();
    return ring + s * SLOT_BYTES;
  };
  auto slot_release = [&]() {
    umma_commit(&bars[B_RING_EMPTY + it % RING]);
    ++it;
  };
  auto qkv = [&]() {                          // D1[128 x 96] = LN(x) [Wq_h; Wk_h; Wv_h]^T
    for (int kp = 0; kp < 4; ++kp) {
      const uint32_t sb = slot_wait();
      const uint64_t da = make_desc_sw128_kmajor(a0 + kp * 16384), db = make_desc_sw128_kmajor(sb);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(tmem + TM_D1, da + uint64_t(k * 2), db + uint64_t(k * 2), id_qkv, (kp | k) != 0 ? 1u : 0u);
      slot_release();
    }
    umma_commit(&bars[B_D1_FULL]);
  };
  bool last_layer = false;
  auto ff1 = [&](int c) {                     // H[c&1][128 x 128] = LN2(x) W1[c*128.., :]^T
    const uint32_t d = tmem + ((c & 1) ? TM_H1 : TM_H0);
    for (int kp = 0; kp < 4; ++kp) {
      const uint32_t sb = slot_wait();
      const uint64_t da = make_desc_sw128_kmajor(a0 + kp * 16384), db = make_desc_sw128_kmajor(sb);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(d, da + uint64_t(k * 2), db + uint64_t(k * 2), id_128, (kp | k) != 0 ? 1u : 0u);
      slot_release();
    }
    umma_commit(&bars[B_HACC_FULL + (c & 1)]);
    if (c == a.n_chunks - 1 && last_layer) umma_commit(&bars[B_A0_FREE]);
  };

  for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    for (int l = 0; l < a.depth; ++l) {
      last_layer = l == a.depth - 1;
      mbar_wait(&bars[B_A0_READY], (n_a0++) & 1);
      tc_fence_after();
      qkv();
      for (int h = 0; h < HEADS; ++h) {
        mbar_wait(&bars[B_STAGED], h & 1);
        tc_fence_after();
        {                                     // S[128 x kmax] = Q_h K_h^T
          const uint64_t da = desc_sw64(qs), db = desc_sw64(ks);
          umma_bf16(tmem + TM_S, da, db, id_s, 0u);
          umma_bf16(tmem + TM_S, da + 2, db + 2, id_s, 1u);
          umma_commit(&bars[B_S_FULL]);
        }
        if (h + 1 < HEADS) qkv();
        mbar_wait(&bars[B_P_READY], h & 1);
        if (h > 0) mbar_wait(&bars[B_O_DRAINED], (h - 1) & 1);
        tc_fence_after();
        for (int k = 0; k < kmax / 16; ++k)   // O[128 x 32] = P V_h   (A from TMEM, B MN-major)
          umma_bf16_ts(tmem + TM_O, tmem + TM_S + uint32_t(k * 8), desc_sw64(vs + (h & 1) * 8192 + k * 1024), id_pv, k != 0 ? 1u : 0u);
        umma_commit(&bars[B_O_FULL]);
      }
      mbar_wait(&bars[B_O_DRAINED], (HEADS - 1) & 1);
      tc_fence_after();
      for (int kp = 0; kp < 4; ++kp)          // x += attn Wout^T  (x + b_out was stored by the workers)
        for (int nh = 0; nh < 2; ++nh) {
          const uint32_t sb = slot_wait();
          const uint64_t da = make_desc_sw128_kmajor(a1 + kp * 16384), db = make_desc_sw128_kmajor(sb);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem + TM_X + nh * 128, da + uint64_t(k * 2), db + uint64_t(k * 2), id_128, 1u);
          slot_release();
        }
      umma_commit(&bars[B_X1_FULL]);

      mbar_wait(&bars[B_A0_READY], (n_a0++) & 1);
      tc_fence_after();
      ff1(0);
      if (a.n_chunks > 1) ff1(1);
      for (int c = 0; c < a.n_chunks; ++c) {
        const int b = c & 1;
        mbar_wait(&bars[B_H_READY + b], (n_hready[b]++) & 1);
        tc_fence_after();
        for (int kp = 0; kp < 2; ++kp)        // x += gelu(H_c) W2[:, c*128..]^T
          for (int nh = 0; nh < 2; ++nh) {
            const uint32_t sb = slot_wait();
            const uint64_t da = make_desc_sw128_kmajor(a1 + b * 32768 + kp * 16384), db = make_desc_sw128_kmajor(sb);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(tmem + TM_X + nh * 128, da + uint64_t(k * 2), db + uint64_t(k * 2), id_128, 1u);
            slot_release();
          }
        if (c + 2 < a.n_chunks) {
          umma_commit(&bars[B_HBUF_FREE + b]);
          ff1(c + 2);
        }
      }
      umma_commit(&bars[B_X2_FULL]);
    }
  }
}

template <int IO>
__global__ void __launch_bounds__(NUM_THREADS, 1) encoder_fused_kernel(const __grid_constant__ FusedArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  if (threadIdx.x == 0 && (smem - smem_raw) + SMEM_USED > SMEM_ALLOC) {
    printf("avf: encoder_fused_kernel: dynamic shared memory base is not 1024-byte aligned enough\n");
    __trap();
  }
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NUM_BARS);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int l = 0; l < a.depth; ++l) {
      tma_prefetch_desc(&a.layer[l].tm_qkv);
      tma_prefetch_desc(&a.layer[l].tm_out);
      tma_prefetch_desc(&a.layer[l].tm_w1);
      tma_prefetch_desc(&a.layer[l].tm_w2);
    }
    for (int i = 0; i < NUM_BARS; ++i) {
      const bool by_workers = i == B_A0_READY || i == B_STAGED || i == B_P_READY || i == B_O_DRAINED || i == B_H_READY || i == B_H_READY1;
      mbar_init(&bars[i], by_workers ? 8 : 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) producer_main<IO>(a, smem, bars);
  } else if (warp == 1) {
    if (lane == 0) mma_main(a, smem, bars, tmem);
  } else {
    worker_main<IO>(a, smem, bars, tmem);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

int sm_count_cached() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

}  // namespace

bool encoder_fused_supported(const avf_stack_shape* s) {
  return s->dim == DIM && s->heads == HEADS && s->dim_head == DH && s->mlp_dim % 128 == 0 && s->mlp_dim >= 128 && s->n_tok >= 1 &&
         s->n_tok <= 64 && s->depth >= 1 && s->depth <= MAX_DEPTH;
}

// io_kind 0: in/out are NCHW bf16 maps [n_seq, 256, n_tok] (pos required); 1: fp32 token rows with strides ld_in / ld_out.
int encoder_fused(int io_kind, const avf_stack_shape* s, const avf_layer_weights* L, const void* in, int ld_in, void* out, int ld_out,
                  const float* pos, cudaStream_t st) {
  AVF_REQUIRE(encoder_fused_supported(s), AVF_EUNSUPPORTED, "fused encoder: unsupported shape dim=%d heads=%d dh=%d mlp=%d n_tok=%d depth=%d",
              s->dim, s->heads, s->dim_head, s->mlp_dim, s->n_tok, s->depth);
  AVF_REQUIRE(io_kind == IO_ROWS_F32 || pos != nullptr, AVF_EINVAL, "fused encoder: NCHW input needs the positional embedding");
  AVF_REQUIRE(io_kind == IO_NCHW_BF16 || (ld_in % 4 == 0 && ld_out % 4 == 0), AVF_EINVAL, "fused encoder: row strides must be multiples of 4");
  static thread_local FusedArgs a;      // ~2 KB of tensor maps + pointers, passed by value (__grid_constant__) per launch
  static_assert(sizeof(FusedArgs) <= 4000, "kernel parameter space");
  a.in = in; a.out = out; a.pos = pos; a.ld_in = ld_in; a.ld_out = ld_out;
  a.n_seq = s->n_seq; a.n_tok = s->n_tok; a.spt = 128 / s->n_tok;
  a.n_tiles = ceil_div(s->n_seq, a.spt);
  a.n_chunks = s->mlp_dim / 128; a.depth = s->depth;
  const int inner = HEADS * DH;
  for (int l = 0; l < s->depth; ++l) {
    LayerArgs& A = a.layer[l];
    int e;
    if ((e = make_tmap_bf16_2d(&A.tm_qkv, L[l].w_qkv, 3 * inner, DIM, DIM, 32))) return e;
    if ((e = make_tmap_bf16_2d(&A.tm_out, L[l].w_out, DIM, inner, inner, 128))) return e;
    if ((e = make_tmap_bf16_2d(&A.tm_w1, L[l].w_ff1, s->mlp_dim, DIM, DIM, 128))) return e;
    if ((e = make_tmap_bf16_2d(&A.tm_w2, L[l].w_ff2, DIM, s->mlp_dim, s->mlp_dim, 128))) return e;
    A.ln1_g = L[l].ln1_gamma; A.ln1_b = L[l].ln1_beta; A.b_out = L[l].b_out;
    A.ln2_g = L[l].ln2_gamma; A.ln2_b = L[l].ln2_beta; A.b_ff1 = L[l].b_ff1; A.b_ff2 = L[l].b_ff2;
  }
  static bool configured = false;
  if (!configured) {
    AVF_CUDA(cudaFuncSetAttribute(encoder_fused_kernel<IO_NCHW_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_ALLOC));
    AVF_CUDA(cudaFuncSetAttribute(encoder_fused_kernel<IO_ROWS_F32>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_ALLOC));
    configured = true;
  }
  const int grid = min(a.n_tiles, sm_count_cached());
  if (io_kind == IO_NCHW_BF16)
    encoder_fused_kernel<IO_NCHW_BF16><<<grid, NUM_THREADS, SMEM_ALLOC, st>>>(a);
  else
    encoder_fused_kernel<IO_ROWS_F32><<<grid, NUM_THREADS, SMEM_ALLOC, st>>>(a);
  AVF_LAUNCH_CHECK("encoder_fused_kernel");
  return 0;
}

}  // namespace avf
