// Fused pre-LN encoder stack, dim 256 / 8 heads x 32, for sequences of 33..64 tokens (the SFormer: 49 tokens per frame,
// models/vformer.py:245-259) — second generation of avf_layer_fused.cu, same mathematics (models/heads.py:164-256), same
// TMEM-resident fp32 residual stream, same weight rings; what changed is WHO does the per-row work and in what order:
//
//   * SIXTEEN row-worker warps instead of eight: four threads per token row, each owning 64 of the 256 columns in the
//     LayerNorm / GELU / tile input / tile output sweeps.  The sweeps are latency-bound (TMEM and shared-memory round trips,
//     MUFU), so twice the warps per scheduler is close to twice the throughput.
//   * attention runs as TWO INDEPENDENT HEAD CHAINS: the worker threads are split into chain 0 (even heads) and chain 1 (odd
//     heads), eight warps = two threads per row each.  A head is a serial chain  QKV -> stage Q/K/V -> S = QK^T -> softmax -> PV
//     -> stage O -> out-projection  of three tensor-core / CUDA-core round trips; with one chain the tensor pipe and the workers
//     take turns waiting for each other (avf_layer_fused.cu: 27 k of 64 k cycles per tile).  Two chains half a head apart fill
//     each other's gaps.
//   * COMPACT SCORES make the second chain fit in TMEM.  A tile holds two sequences in two 64-row slots, and a row only needs
//     the 64 keys of its own sequence:  S[128 x 64] = Qm0 K[0:64]^T + Qm1 K[64:128]^T  with Qm_s = Q_h with the rows of the OTHER
//     sequence zero (two staged copies of Q_h whose zero halves are written once per layer).  S is then 64 columns instead of
//     128, and the 128 score columns of TMEM hold one buffer per chain.  P_h (bf16, 128 keys = 64 columns, zero outside the
//     row's own sequence) overwrites its own S_h and is the TMEM A operand of O = P V.
//
// TMEM columns: [0,256) fp32 residual stream X | [256,352) D1 = [Q|K|V]_h | [352,384) O_h | [384,448) S/P chain 0 | [448,512) S/P
// chain 1; in the MLP phase [256,384) / [384,512) are the two hidden-layer accumulators.
//
// Order of the MMA thread inside a layer (fixed; every wait is for an event that only depends on MMAs already issued):
//   QKV(0) | S(0) QKV(1) | S(1) QKV(2) PV(0) | S(2) QKV(3) out(0) PV(1) | ... | out(6) PV(7) | out(7)
// Hazards: K and Q staging are single (E1(h+1) follows D1(h+1), issued behind S(h)); V is one buffer per chain (a chain stages
// head h+2 only after it has drained O(h), i.e. after PV(h)); O and its staging buffer are single (PV(h) is issued behind
// out(h-1), and E3(h) follows PV(h)); D1 is single (QKV(h+1) is issued after E1(h) has read it).
//
// Warp roles: 0 = TMA producer (out-projection / MLP weights, next tile's frames), 1 = TMEM allocator + MMA issuer,
// 2 = TMA producer of the QKV ring, 3..18 = row workers (thread (row, g): g = 2 * chain + half).
#include <cuda.h>

#include "avf_common.cuh"
#include "avf_fused_helpers.cuh"
#include "avf_internal.h"

namespace avf {

int make_tmap_bf16_2d(CUtensorMap* tm, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows, uint32_t box_cols = 64);

namespace {
using namespace fused;

constexpr int DIM = 256, HEADS = 8, DH = 32;
constexpr int MAX_DEPTH = 3, MAX_MLP = 1024;
constexpr int NT = 4;                      // threads per token row
constexpr int WORKER_T0 = 96;              // first worker thread
constexpr int NUM_WORKERS = 128 * NT;
constexpr int NUM_THREADS = WORKER_T0 + NUM_WORKERS;
constexpr int CW = DIM / NT;               // residual-stream columns per thread in the sweeps
constexpr int CHAIN_WARPS = NUM_WORKERS / 64;
// Main weight ring, 16 KB slots, as in avf_layer_fused.cu: slots 2-4 are free in every phase (out-projection slices), the MLP
// weights rotate over all nine (slots 0-1 alias the Q/K/V staging, slots 5-8 the 64 KB of A1, idle in the MLP phase).
constexpr int RING = 9, RING_LO = 2, RING_HI = 5, SLOT_BYTES = 16384;
constexpr int QRING = 4, QSLOT_BYTES = 12288;   // QKV ring: one slot = one 64-wide K panel of [Wq_h; Wk_h; Wv_h]

// shared memory map (offsets from a 1024-byte aligned base)
constexpr int OFF_A0 = 0;                          // 64 KB  LN output [128 x 256] bf16, 4 K-panels; also the NCHW input staging
constexpr int OFF_A1 = 65536;                      // 64 KB  attention phase: O staging, Qm1, QKV ring; NCHW output staging; ring slots 5-8
constexpr int OFF_OST = OFF_A1;                    //   8 KB  O_h / l  [128 x 32] K-major SW64
constexpr int OFF_Q1 = OFF_A1 + 8192;              //   8 KB  Qm1: Q_h rows of sequence 1, rows 0-63 zero
constexpr int OFF_QRING = OFF_A1 + 16384;          //   4 x 12 KB QKV weight ring
constexpr int OFF_Q = 131072;                      // 8 KB   Qm0: Q_h rows of sequence 0, rows 64-127 zero  [128 x 32] K-major SW64
constexpr int OFF_K = OFF_Q + 8192;                // 8 KB   K_h
constexpr int OFF_V = OFF_K + 8192;                // 2x8 KB V_h [128 tok x 32] MN-major SW64, one buffer per chain
constexpr int OFF_RING = OFF_V + 16384;            // 3 x 16 KB weight ring (slots 2-4)
static_assert(OFF_RING == OFF_Q + RING_LO * SLOT_BYTES, "ring slots are contiguous from OFF_Q");
constexpr int OFF_XCH = OFF_RING + (RING_HI - RING_LO) * SLOT_BYTES;     // row exchange [2][128][4] floats
constexpr int OFF_SUMS = OFF_XCH + 2 * 128 * 4 * 4;                      // partial softmax row sums [2 chains][128][2] floats
constexpr int OFF_VEC = OFF_SUMS + 2 * 128 * 2 * 4;                      // per-layer vectors
__host__ __device__ constexpr int slot_offset(uint32_t s) { return s < RING_HI ? OFF_Q + int(s) * SLOT_BYTES : OFF_A1 + int(s - RING_HI) * SLOT_BYTES; }
constexpr int V_BOUT = 0, V_BFF2 = 256, V_BFF1 = 512;     // fp32
constexpr int V_B16 = V_BFF1 + MAX_MLP;                   // then bf16 copies of the LayerNorm affine vectors: [ln1_g | ln1_b | ln2_g | ln2_b] x 256
constexpr int OFF_BAR = OFF_VEC + (V_B16 + 512) * 4;
constexpr int SMEM_USED = OFF_BAR + 512;
constexpr int SMEM_ALLOC = SMEM_USED + 1024;       // slack for the 1024-byte alignment of the base
static_assert(SMEM_ALLOC <= 232448, "shared memory budget");

// TMEM columns
constexpr uint32_t TM_X = 0, TM_D1 = 256, TM_O = 352, TM_S = 384, TM_H0 = 256, TM_H1 = 384;

enum {
  B_RING_FULL = 0, B_RING_EMPTY = RING, B_X0_FULL = 2 * RING, B_A0_FREE, B_A0_READY,
  B_D1_FULL, B_STAGED = B_D1_FULL + 2, B_S_FULL = B_STAGED + 2, B_P_READY = B_S_FULL + 2, B_O_FULL = B_P_READY + 2, B_O_DRAINED = B_O_FULL + 2,
  B_X1_FULL = B_O_DRAINED + 2, B_HACC_FULL, B_HACC_FULL1, B_H_READY, B_H_READY1, B_X2_FULL,
  B_QR_FULL, B_QR_EMPTY = B_QR_FULL + QRING, B_A1_FREE = B_QR_EMPTY + QRING, B_OUT_READ, B_QKV_FREE, NUM_BARS
};
static_assert(NUM_BARS * 8 + 8 <= 512, "barrier block");

// named barriers (bar.sync): 0 = __syncthreads, 1..4 = the four threads of the rows of one lane quarter (128 threads),
// 5..12 = the two threads of a row inside one head chain (64 threads), 13 = all workers
constexpr int NB_ROW4 = 1, NB_ROW2 = 5, NB_ALL = 13;

enum { IO_NCHW_BF16 = 0, IO_ROWS_F32 = 1 };
constexpr int POS_LD = 64;   // row stride of the channel-major positional table

// worker phases (chain 0's first thread) / MMA thread phases of the -DAVF_FUSED_PROF build; same slots as avf_layer_fused.cu
enum { PW_INPUT = 0, PW_VEC, PW_LN1, PW_WAIT_D1, PW_E1, PW_WAIT_O, PW_E3, PW_WAIT_S, PW_E2, PW_WAIT_X1, PW_LN2, PW_WAIT_HACC, PW_GELU, PW_WAIT_X2,
       PW_OUTPUT, PW_TILES, PW_E2_LD, PW_E2_EXP, PW_E2_XCH, PW_E2_ST,
       PM_WAIT_A0 = 32, PM_QKV, PM_WAIT_STAGED, PM_S, PM_WAIT_P, PM_PV, PM_WAIT_OD7, PM_OUT, PM_WAIT_A0B, PM_FF1, PM_WAIT_H, PM_FF2, PM_RINGWAIT };

struct LayerArgs {
  CUtensorMap tm_qkv, tm_out, tm_w1, tm_w2;     // tm_out: box [256 rows x 32 cols], 64B swizzle (one head's K slice of Wout)
  const float *ln1_g, *ln1_b, *b_out, *ln2_g, *ln2_b, *b_ff1, *b_ff2;
  uint64_t pad_;
};

// A tile is 128 token rows = two sequences in two 64-row slots: row r belongs to sequence r / 64, token r % 64.
struct FusedArgs {
  LayerArgs layer[MAX_DEPTH];
  const void* in;
  void* out;
  const float* pos;          // [n_tok, 256] or nullptr (IO_ROWS_F32)
  const float* pos_t;        // IO_NCHW_BF16: the same table as [256 / 4][POS_LD tokens][4 channels]
  int ld_in, ld_out;         // IO_ROWS_F32 row strides (elements)
  int n_seq, n_tok, n_tiles, n_chunks, depth;
};

// ---------------------------------------------------------------------------------------------
// row workers: thread (row, g) owns columns [g*CW, g*CW+CW) of token row `row` (= TMEM lane) in the sweeps; in the attention
// phase g = 2 * chain + half
// ---------------------------------------------------------------------------------------------
struct Worker {
  uint8_t* smem;
  uint64_t* bars;
  const float* vec;     // per-layer vectors in shared memory
  uint32_t tl;          // TMEM address of this warp's lane quarter, column 0
  int lane, q, g, row, chain, half;
  uint32_t xslot;

  // all-gather of one float between the four threads of a row
  __device__ __forceinline__ float exchange4_sum(float mine) {
    float* s = reinterpret_cast<float*>(smem + OFF_XCH) + xslot * (128 * 4);
    xslot ^= 1;
    s[row * 4 + g] = mine;
    bar_sync(NB_ROW4 + q, 32 * NT);
    const float4 f = *reinterpret_cast<const float4*>(s + row * 4);
    return (f.x + f.y) + (f.z + f.w);
  }
  // maximum over the two threads of a row inside one head chain
  __device__ __forceinline__ float exchange2_max(float mine) {
    float* s = reinterpret_cast<float*>(smem + OFF_XCH) + xslot * (128 * 4);
    xslot ^= 1;
    s[row * 4 + g] = mine;
    bar_sync(NB_ROW2 + chain * 4 + q, 64);
    const float2 f = *reinterpret_cast<const float2*>(s + row * 4 + chain * 2);
    return fmaxf(f.x, f.y);
  }
  __device__ __forceinline__ void arrive(int bar) {             // one arrive per warp, after every lane's fences
    fence_proxy_async_smem();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&bars[bar]);
  }
  __device__ __forceinline__ void arrive_tmem_only(int bar) {   // the phase wrote TMEM only (no shared-memory operand for the async proxy)
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&bars[bar]);
  }
  // Statistics are accumulated as shifted sums (shift = the thread's first element) and merged across the row's threads
  // by Chan's formula, so one sweep over the row is enough and nothing cancels catastrophically.
  struct Stats {
    float shift, s, ss;
    __device__ __forceinline__ void add(const float (&x)[32], bool first) {
      if (first) shift = x[0];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float d = x[j] - shift;
        s += d;
        ss = fmaf(d, d, ss);
      }
    }
  };
  __device__ __forceinline__ void finish_stats(const Stats& st, float& mean, float& rstd) {
    constexpr float n = float(CW);
    const float mean_g = st.shift + st.s * (1.f / n);
    const float m2_g = fmaxf(st.ss - st.s * st.s * (1.f / n), 0.f);
    mean = exchange4_sum(mean_g) * (1.f / NT);
    const float dm = mean_g - mean;
    const float m2 = exchange4_sum(fmaf(dm * dm, n, m2_g));
    rstd = rsqrtf(m2 * (1.f / DIM) + 1e-5f);
  }
  // Sweep 1 of a LayerNorm whose input already sits in TMEM (x1 after the out-projection, x2 after the MLP).
  __device__ __forceinline__ void stats_from_tmem(float& mean, float& rstd) {
    Stats st{0.f, 0.f, 0.f};
#pragma unroll
    for (int c0 = 0; c0 < CW; c0 += 32) {
      float x[32];
      tmem_ld32(tl + TM_X + g * CW + c0, reinterpret_cast<uint32_t(&)[32]>(x));
      tmem_ld_wait();
      st.add(x, c0 == 0);
    }
    finish_stats(st, mean, rstd);
  }
  // Sweep 2: LN(x) -> A0 (bf16, K-major SW128 panels), x + next_bias -> TMEM; then signal the MMA thread.
  // (x - mean) * rstd in fp32, the affine part on bf16 pairs (gamma / beta pre-rounded to bf16 in shared memory).
  __device__ __forceinline__ void normalize_from_tmem(float mean, float rstd, int ln_sel, int v_next_bias) {
    const int cbase = g * CW;
    const float nmr = -mean * rstd;
#pragma unroll 1
    for (int c0 = 0; c0 < CW; c0 += 32) {
      float x[32];
      tmem_ld32(tl + TM_X + cbase + c0, reinterpret_cast<uint32_t(&)[32]>(x));
      tmem_ld_wait();
      const int col0 = cbase + c0;
      uint8_t* dst = smem + OFF_A0 + (col0 >> 6) * 16384 + row * 128;
      const int chunk0 = (col0 & 63) >> 3;
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        const __nv_bfloat16* v16 = reinterpret_cast<const __nv_bfloat16*>(vec + V_B16) + ln_sel * 512 + col0 + ch * 8;
        const uint4 gm = *reinterpret_cast<const uint4*>(v16), bt = *reinterpret_cast<const uint4*>(v16 + 256);
        uint4 pk;
        pk.x = fma_bf16x2(cvt_bf16x2(fmaf(x[ch * 8], rstd, nmr), fmaf(x[ch * 8 + 1], rstd, nmr)), gm.x, bt.x);
        pk.y = fma_bf16x2(cvt_bf16x2(fmaf(x[ch * 8 + 2], rstd, nmr), fmaf(x[ch * 8 + 3], rstd, nmr)), gm.y, bt.y);
        pk.z = fma_bf16x2(cvt_bf16x2(fmaf(x[ch * 8 + 4], rstd, nmr), fmaf(x[ch * 8 + 5], rstd, nmr)), gm.z, bt.z);
        pk.w = fma_bf16x2(cvt_bf16x2(fmaf(x[ch * 8 + 6], rstd, nmr), fmaf(x[ch * 8 + 7], rstd, nmr)), gm.w, bt.w);
        *reinterpret_cast<uint4*>(dst + (((chunk0 + ch) ^ (row & 7)) << 4)) = pk;
      }
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b = *reinterpret_cast<const float4*>(vec + v_next_bias + col0 + j);
        x[j] += b.x; x[j + 1] += b.y; x[j + 2] += b.z; x[j + 3] += b.w;
      }
      tmem_st32(tl + TM_X + col0, reinterpret_cast<const uint32_t(&)[32]>(x));
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
  w.vec = reinterpret_cast<const float*>(smem + OFF_VEC);
  const int wt = threadIdx.x - WORKER_T0;
  w.lane = wt & 31;
  w.q = (threadIdx.x >> 5) & 3;          // TMEM lane quarter = warp index mod 4
  w.g = wt >> 7;                         // 0..3
  w.chain = w.g >> 1;
  w.half = w.g & 1;
  w.row = w.q * 32 + w.lane;
  w.tl = tmem + (uint32_t(w.q * 32) << 16);
  w.xslot = 0;
  const int g = w.g, row = w.row, chain = w.chain, half = w.half;
  const uint32_t tl = w.tl;
  const float* vec = w.vec;
  const int n_tok = a.n_tok;
  const int seq_in_tile = row >> 6, t_in_seq = row & 63;
  // softmax geometry (compact scores): this row attends to score columns [0, n_tok) of its chain's buffer; the two threads of
  // the row split the 8-column chunks [0, c_hi)
  const int c_hi = (n_tok + 7) >> 3;
  constexpr int MAXC = 4;
  const int per = (c_hi + 1) >> 1;
  const int my_c0 = half * per;
  const int my_nc = max(0, min(per, c_hi - my_c0));
  const float sm_scale = 1.4426950408889634f * rsqrtf(float(DH));
  const uint32_t tm_s = TM_S + uint32_t(chain) * 64;        // this chain's score / P buffer
  const uint32_t tm_p = tm_s + uint32_t(seq_in_tile) * 32;  // where this row's P (its own sequence's 64 keys) starts
  float* const sums = reinterpret_cast<float*>(smem + OFF_SUMS) + chain * 256 + row * 2;
  uint32_t n_x0 = 0, n_x1 = 0, n_x2 = 0, n_hacc0 = 0, n_hacc1 = 0;
  bool vec_loaded = false;
  Prof pf;
  pf.start(blockIdx.x == 0 && wt == 0);

  for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    const int seqs_here = min(2, a.n_seq - tile * 2);
    const bool valid = seq_in_tile < seqs_here && t_in_seq < n_tok;
    const size_t grow = (size_t(tile) * 2 + seq_in_tile) * n_tok + t_in_seq;     // global token row (IO_ROWS_F32)

    // ---- tile input: x (+ pos) -> TMEM, statistics of LN1 on the way --------------------------------------
    float mean, rstd;
    {
      Worker::Stats st{0.f, 0.f, 0.f};
      // Branch-free: padding rows (and sequences beyond the end of the batch) read a valid location, i.e. they carry a copy of a
      // real token.  Nothing ever looks at them: as keys they are outside every row's softmax window, as rows they are not stored.
      const int t_safe = valid ? t_in_seq : 0, seq_safe = valid ? seq_in_tile : 0;
      const float* pp = (IO == IO_ROWS_F32 && a.pos != nullptr) ? a.pos + t_safe * DIM + g * CW : nullptr;
      const float4* ppt = reinterpret_cast<const float4*>(a.pos_t) + size_t(g * CW / 4) * POS_LD + t_safe;
      if constexpr (IO == IO_NCHW_BF16) mbar_wait(&bars[B_X0_FULL], (n_x0++) & 1);
      const __nv_bfloat16* src16 = reinterpret_cast<const __nv_bfloat16*>(smem + OFF_A0) + size_t(seq_safe * DIM + g * CW) * n_tok + t_safe;
      const float* src32 = static_cast<const float*>(a.in) + ((size_t(tile) * 2 + seq_safe) * n_tok + t_safe) * a.ld_in + g * CW;
#pragma unroll 1
      for (int c0 = 0; c0 < CW; c0 += 32) {
        float x[32];
        if constexpr (IO == IO_NCHW_BF16) {
          float4 p[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) p[j] = __ldg(ppt + (c0 / 4 + j) * POS_LD);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            x[4 * j] = __bfloat162float(src16[(c0 + 4 * j) * n_tok]) + p[j].x;
            x[4 * j + 1] = __bfloat162float(src16[(c0 + 4 * j + 1) * n_tok]) + p[j].y;
            x[4 * j + 2] = __bfloat162float(src16[(c0 + 4 * j + 2) * n_tok]) + p[j].z;
            x[4 * j + 3] = __bfloat162float(src16[(c0 + 4 * j + 3) * n_tok]) + p[j].w;
          }
        } else {
          float4 p4[8];
          if (pp != nullptr) {
#pragma unroll
            for (int j = 0; j < 8; ++j) p4[j] = __ldg(reinterpret_cast<const float4*>(pp + c0) + j);
          }
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 x4 = *reinterpret_cast<const float4*>(src32 + c0 + j);
            x[j] = x4.x; x[j + 1] = x4.y; x[j + 2] = x4.z; x[j + 3] = x4.w;
          }
          if (pp != nullptr) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              x[4 * j] += p4[j].x; x[4 * j + 1] += p4[j].y; x[4 * j + 2] += p4[j].z; x[4 * j + 3] += p4[j].w;
            }
          }
        }
        st.add(x, c0 == 0);
        tmem_st32(tl + TM_X + g * CW + c0, reinterpret_cast<const uint32_t(&)[32]>(x));
      }
      tmem_st_wait();
      w.finish_stats(st, mean, rstd);
      if constexpr (IO == IO_NCHW_BF16) {
        // the previous tile's output store must have finished READING A1 before this tile's attention writes it again: waited
        // for here (one thread, ordered for everybody by the barrier below) instead of right after issuing it
        if (threadIdx.x == WORKER_T0) {
          bulk_wait_read0();
          mbar_arrive(&bars[B_OUT_READ]);   // ... and before the QKV ring (which lives in A1) is refilled for this tile
        }
        bar_sync(NB_ALL, NUM_WORKERS);      // every worker has read its part of the staged frames out of A0
      }
    }
    pf.mark(PW_INPUT);

    for (int l = 0; l < a.depth; ++l) {
      const LayerArgs& L = a.layer[l];
      if (a.depth > 1 || !vec_loaded) {          // per-layer vectors -> shared memory (once per CTA when there is one layer)
        if (vec_loaded) bar_sync(NB_ALL, NUM_WORKERS);  // previous layer's readers are done
        float* vs = reinterpret_cast<float*>(smem + OFF_VEC);
        for (int i = wt; i < DIM; i += NUM_WORKERS) {
          vs[V_BOUT + i] = L.b_out[i]; vs[V_BFF2 + i] = L.b_ff2[i];
          __nv_bfloat16* v16 = reinterpret_cast<__nv_bfloat16*>(vs + V_B16);
          v16[i] = __float2bfloat16_rn(L.ln1_g[i]); v16[256 + i] = __float2bfloat16_rn(L.ln1_b[i]);
          v16[512 + i] = __float2bfloat16_rn(L.ln2_g[i]); v16[768 + i] = __float2bfloat16_rn(L.ln2_b[i]);
        }
        for (int i = wt; i < a.n_chunks * 128; i += NUM_WORKERS) vs[V_BFF1 + i] = L.b_ff1[i];
        bar_sync(NB_ALL, NUM_WORKERS);
        vec_loaded = true;
      }
      pf.mark(PW_VEC);
      // The halves of the two masked Q copies that stay zero through the eight heads: rows 64-127 of Qm0, rows 0-63 of Qm1
      // (the MLP weights of the previous layer / tile and the tile output went through these buffers).  Made visible to the tensor
      // core by the fence of the A0_READY arrive below.
      {
        uint8_t* zp = (wt < 256) ? smem + OFF_Q + 4096 + wt * 16 : smem + OFF_Q1 + (wt - 256) * 16;
        *reinterpret_cast<uint4*>(zp) = make_uint4(0u, 0u, 0u, 0u);
      }
      w.normalize_from_tmem(mean, rstd, 0, V_BOUT);
      pf.mark(PW_LN1);

      // ---- attention: this chain's four heads (chain 0: 0, 2, 4, 6; chain 1: 1, 3, 5, 7) ---------------------------
#pragma unroll 1
      for (int i = 0; i < HEADS / 2; ++i) {
        const uint32_t par = uint32_t(i & 1);
        float sum_g;
        {                        // E1: D1 = [Q|K|V]_h as 12 chunks of 8 columns, 6 per thread: half 0 -> Q0-3 K0-1, half 1 -> K2-3 V0-3
          mbar_wait(&bars[B_D1_FULL + chain], par);
          tc_fence_after();
          pf.mark(PW_WAIT_D1);
          constexpr int NCH = 6;
          const uint32_t sw = uint32_t((row >> 1) & 3);
          uint32_t r[NCH * 8];
          tmem_ld32(tl + TM_D1 + half * 48, reinterpret_cast<uint32_t(&)[32]>(r[0]));
          tmem_ld16(tl + TM_D1 + half * 48 + 32, reinterpret_cast<uint32_t(&)[16]>(r[32]));
          tmem_ld_wait();
          uint8_t* const qb = smem + (seq_in_tile ? OFF_Q1 : OFF_Q);
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
            const int id = half * NCH + c;                    // 0..3 Q, 4..7 K, 8..11 V   (compile-time per half after unrolling)
            uint8_t* base = id < 4 ? qb : (id < 8 ? smem + OFF_K : smem + OFF_V + chain * 8192);
            *reinterpret_cast<uint4*>(base + row * 64 + ((uint32_t(id & 3) ^ sw) << 4)) = pack8u(&r[c * 8]);
          }
          w.arrive(B_STAGED + chain);
          pf.mark(PW_E1);
        }
        {                        // E2: softmax of this row over the keys of its own sequence, P (bf16) over the S columns
          mbar_wait(&bars[B_S_FULL + chain], par);
          tc_fence_after();
          pf.mark(PW_WAIT_S);
          float s[MAXC][8];
#pragma unroll
          for (int c = 0; c < MAXC; ++c)
            if (c < my_nc) tmem_ld8(tl + tm_s + (my_c0 + c) * 8, reinterpret_cast<uint32_t(&)[8]>(s[c]));
          tmem_ld_wait();
          pf.mark(PW_E2_LD);
          float mloc = -INFINITY;
#pragma unroll
          for (int c = 0; c < MAXC; ++c) {
            if (c < my_nc) {
              const int col0 = (my_c0 + c) * 8;
              if (col0 + 8 > n_tok) {        // boundary chunk: mask per column
#pragma unroll
                for (int j = 0; j < 8; ++j) s[c][j] = (col0 + j < n_tok) ? s[c][j] : -INFINITY;
              }
              mloc = fmaxf(mloc, fmaxf(fmaxf(fmaxf(s[c][0], s[c][1]), fmaxf(s[c][2], s[c][3])), fmaxf(fmaxf(s[c][4], s[c][5]), fmaxf(s[c][6], s[c][7]))));
            }
          }
          // the row maximum first (one exchange; it also orders every S load of the row before any P store — P is written over
          // the S columns), then 2^(s - max) straight to its final bf16 value, two per MUFU instruction; the partial row sums
          // (fp32, of the rounded values the MMA will see) meet again in E3 through shared memory, no second barrier
          pf.mark(PW_E2_EXP);
          const float mrow = w.exchange2_max(mloc);
          pf.mark(PW_E2_XCH);
          const float nml = -mrow * sm_scale;            // every row has at least one valid column: mrow is finite
          sum_g = 0.f;
#pragma unroll
          for (int c = 0; c < MAXC; ++c) {
            if (c < my_nc) {
              uint32_t e[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                e[j] = ex2_bf16x2(cvt_bf16x2(fmaf(s[c][2 * j], sm_scale, nml), fmaf(s[c][2 * j + 1], sm_scale, nml)));
                sum_g += __uint_as_float(e[j] << 16) + __uint_as_float(e[j] & 0xFFFF0000u);
              }
              tmem_st4(tl + tm_p + (my_c0 + c) * 4, e[0], e[1], e[2], e[3]);
            }
          }
          sums[half] = sum_g;
          // P columns the MMA reads (64 = 128 keys) but this row does not own: zeros.  Thread `half` clears TMEM columns
          // [32 half, 32 half + 32) of the buffer outside the row's window [32 seq, 32 seq + 4 c_hi) — columns whose scores it has
          // read itself (half != seq: its own chunks; half == seq: the tail behind the last chunk, which nobody reads).
          if (half != seq_in_tile) {
            tmem_st16_zero(tl + tm_s + half * 32);
            tmem_st16_zero(tl + tm_s + half * 32 + 16);
          } else {
            zero_p_columns(tl + tm_s, half * 32 + c_hi * 4, half * 32 + 32);
          }
          tmem_st_wait();
          pf.mark(PW_E2_ST);
          w.arrive_tmem_only(B_P_READY + chain);
          pf.mark(PW_E2);
        }
        {                        // E3: O / l -> bf16 -> staging buffer, columns [half*16, +16) (A operand of the per-head out-projection)
          mbar_wait(&bars[B_O_FULL + chain], par);
          tc_fence_after();
          pf.mark(PW_WAIT_O);
          uint32_t r[16];
          tmem_ld16(tl + TM_O + half * 16, r);
          tmem_ld_wait();
          float inv_l;
          {   // l = own + partner's partial row sum (the partner wrote it before its P_READY arrive, which PV -> O_FULL follows)
            const float2 ps = *reinterpret_cast<const float2*>(sums);
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv_l) : "f"(ps.x + ps.y));
          }
          uint8_t* ob = smem + OFF_OST + row * 64;
          const uint32_t sw = uint32_t((row >> 1) & 3);
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            float y[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) y[j] = __uint_as_float(r[c * 8 + j]) * inv_l;
            *reinterpret_cast<uint4*>(ob + ((uint32_t(half * 2 + c) ^ sw) << 4)) = pack8(y);
          }
          w.arrive(B_O_DRAINED + chain);
          pf.mark(PW_E3);
        }
      }

      // ---- LN2 on x1 = x + attention (accumulated in TMEM by the out-projection) ------------------------------
      mbar_wait(&bars[B_X1_FULL], (n_x1++) & 1);
      tc_fence_after();
      pf.mark(PW_WAIT_X1);
      w.stats_from_tmem(mean, rstd);
      w.normalize_from_tmem(mean, rstd, 1, V_BFF2);
      pf.mark(PW_LN2);

      // ---- MLP: per 128-column chunk of the hidden layer, bias + tanh-GELU -> bf16 A operand ---------------------
      // The bf16 result goes back INTO the accumulator's TMEM columns: thread g turns its fp32 columns [32 g, 32 g + 32) of the
      // chunk into 16 packed columns at [32 g, 32 g + 16) — columns it has already read — and the second MLP GEMM takes its A
      // operand from TMEM (16-wide k-step kk sits at column 32 (kk / 2) + 8 (kk % 2)).
#pragma unroll 1
      for (int c = 0; c < a.n_chunks; ++c) {
        const int b = c & 1;
        if (b == 0) mbar_wait(&bars[B_HACC_FULL], (n_hacc0++) & 1);
        else mbar_wait(&bars[B_HACC_FULL1], (n_hacc1++) & 1);
        tc_fence_after();
        pf.mark(PW_WAIT_HACC);
        const uint32_t hbase = tl + (b ? TM_H1 : TM_H0) + g * 32;
        const float* bias = vec + V_BFF1 + c * 128 + g * 32;
        uint32_t r[32], yk[16];
        tmem_ld32(hbase, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(bias + j);
          yk[j / 2] = gelu_bf16x2(cvt_bf16x2(__uint_as_float(r[j]) + b4.x, __uint_as_float(r[j + 1]) + b4.y));
          yk[j / 2 + 1] = gelu_bf16x2(cvt_bf16x2(__uint_as_float(r[j + 2]) + b4.z, __uint_as_float(r[j + 3]) + b4.w));
        }
        tmem_st16(hbase, yk);
        tmem_st_wait();
        w.arrive_tmem_only(B_H_READY + b);
        pf.mark(PW_GELU);
      }

      // ---- x2 = x1 + MLP, accumulated in TMEM by the second MLP GEMM --------------------------------------------
      mbar_wait(&bars[B_X2_FULL], (n_x2++) & 1);
      tc_fence_after();
      pf.mark(PW_WAIT_X2);
      if (l + 1 < a.depth) w.stats_from_tmem(mean, rstd);
    }

    // ---- tile output ----------------------------------------------------------------------------
    {
      __nv_bfloat16* dst16 = reinterpret_cast<__nv_bfloat16*>(smem + OFF_A1) + size_t(seq_in_tile * DIM + g * CW) * n_tok + t_in_seq;
      float* dst32 = static_cast<float*>(a.out) + grow * a.ld_out + g * CW;
#pragma unroll 1
      for (int c0 = 0; c0 < CW; c0 += 32) {
        float x[32];
        tmem_ld32(tl + TM_X + g * CW + c0, reinterpret_cast<uint32_t(&)[32]>(x));
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
        bar_sync(NB_ALL, NUM_WORKERS);
        if (threadIdx.x == WORKER_T0) {
          __nv_bfloat16* gdst = static_cast<__nv_bfloat16*>(a.out) + size_t(tile) * 2 * n_tok * DIM;
          bulk_store_1d(gdst, smem + OFF_A1, uint32_t(seqs_here) * n_tok * DIM * 2);   // read completion: see the next tile's input sweep
        }
      }
    }
    pf.mark(PW_OUTPUT);
    pf.count(PW_TILES);
  }
  pf.flush(0, 32);
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
    const int seqs_here = min(2, a.n_seq - load_tile * 2);
    const uint32_t bytes = uint32_t(seqs_here) * a.n_tok * DIM * 2;
    mbar_expect_tx(&bars[B_X0_FULL], bytes);
    bulk_load_1d(smem + OFF_A0, static_cast<const __nv_bfloat16*>(a.in) + size_t(load_tile) * 2 * a.n_tok * DIM, bytes, &bars[B_X0_FULL]);
    load_tile += gridDim.x;
    need_free = true;
  };
  uint32_t pbits = 0, n_gate = 0;
  auto poll_wait = [&](uint64_t* bar, uint32_t parity) {
    poll_loader();
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
      poll_loader();
      if (clock64() - t0 > 4000000000LL) {
        printf("avf: fused SFormer producer timed out (block %d item %u)\n", blockIdx.x, it);
        __trap();
      }
    }
  };
  auto load = [&](uint32_t s, const CUtensorMap* tm, int c0, int c1) {
    poll_wait(&bars[B_RING_EMPTY + s], ((pbits >> s) & 1u) ^ 1u);
    pbits ^= 1u << s;
    mbar_expect_tx(&bars[B_RING_FULL + s], SLOT_BYTES);
    tma_load_2d(smem + slot_offset(s), tm, &bars[B_RING_FULL + s], c0, c1);
    ++it;
  };
  for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    for (int l = 0; l < a.depth; ++l) {
      const LayerArgs& L = a.layer[l];
      for (int h = 0; h < HEADS; ++h) load(RING_LO + h % (RING_HI - RING_LO), &L.tm_out, h * DH, 0);   // Wout[:, 32h .. 32h+32): [256 x 32], 64B swizzle
      poll_wait(&bars[B_QKV_FREE], (n_gate++) & 1);          // the attention MMAs are done with the Q/K/V staging and A1 = slots 0-1, 5-8
      uint32_t m = 0;
      auto ff1 = [&](int c) {
        for (int kp = 0; kp < 4; ++kp) load((m++) % RING, &L.tm_w1, kp * 64, c * 128);
      };
      auto ff2 = [&](int c) {
        for (int kp = 0; kp < 2; ++kp)
          for (int nh = 0; nh < 2; ++nh) load((m++) % RING, &L.tm_w2, c * 128 + kp * 64, nh * 128);
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

// QKV weight producer (one thread of its own warp): the 4 x 12 KB ring inside A1
template <int IO>
__device__ void qkv_producer_main(const FusedArgs& a, uint8_t* smem, uint64_t* bars) {
  uint32_t qit = 0, n_free = 0, n_read = 0;
  bool first = true;
  for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    // A1 doubles as the NCHW output staging: the previous tile's bulk store must have read it (signalled after this tile's input sweep)
    if constexpr (IO == IO_NCHW_BF16) mbar_wait(&bars[B_OUT_READ], (n_read++) & 1);
    for (int l = 0; l < a.depth; ++l) {
      const LayerArgs& L = a.layer[l];
      if (!first) mbar_wait(&bars[B_A1_FREE], (n_free++) & 1);   // the previous layer's MLP weights (ring slots 5-8 live in A1) are consumed
      first = false;
      for (int h = 0; h < HEADS; ++h)
        for (int kp = 0; kp < 4; ++kp) {
          const uint32_t s = qit % QRING, ph = (qit / QRING) & 1;
          mbar_wait(&bars[B_QR_EMPTY + s], ph ^ 1);
          uint64_t* fb = &bars[B_QR_FULL + s];
          uint8_t* d = smem + OFF_QRING + s * QSLOT_BYTES;
          mbar_expect_tx(fb, QSLOT_BYTES);
          for (int s3 = 0; s3 < 3; ++s3) tma_load_2d(d + s3 * 4096, &L.tm_qkv, fb, kp * 64, s3 * (HEADS * DH) + h * DH);
          ++qit;
        }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// MMA issuer: executed by ALL 32 lanes of warp 1 with warp-uniform values; only the elected lane issues tcgen05.mma / commit
// ---------------------------------------------------------------------------------------------
__device__ void mma_main(const FusedArgs& a, uint8_t* smem, uint64_t* bars, uint32_t tmem) {
  const bool leader = elect_one();
  const uint32_t a0 = smem_u32(smem + OFF_A0);
  const uint32_t q0s = smem_u32(smem + OFF_Q), q1s = smem_u32(smem + OFF_Q1), ks = smem_u32(smem + OFF_K), vs = smem_u32(smem + OFF_V);
  const uint32_t smem0 = smem_u32(smem), qring = smem_u32(smem + OFF_QRING), ost = smem_u32(smem + OFF_OST);
  const uint32_t id_qkv = make_idesc_bf16(128, 96), id_s = make_idesc_bf16(128, 64), id_pv = make_idesc_bf16(128, DH, 0, 1),
                 id_128 = make_idesc_bf16(128, 128), id_256 = make_idesc_bf16(128, 256);
  uint32_t cbits = 0, m = 0, qit = 0, n_a0 = 0, n_hready[2] = {0, 0};
  Prof pf;
  pf.start(blockIdx.x == 0 && leader);
  int ring_phase = PM_QKV;
  auto slot_wait = [&](uint32_t s) -> uint32_t {     // same slot sequence as producer_main
    pf.mark(ring_phase);
    mbar_wait(&bars[B_RING_FULL + s], (cbits >> s) & 1u);
    tc_fence_after();
    pf.mark(PM_RINGWAIT);
    return smem0 + uint32_t(slot_offset(s));
  };
  auto slot_release = [&](uint32_t s) {
    if (leader) umma_commit(&bars[B_RING_EMPTY + s]);
    cbits ^= 1u << s;
  };
  auto qkv = [&](int h) {                     // D1[128 x 96] = LN(x) [Wq_h; Wk_h; Wv_h]^T, weights from the QKV ring
    for (int kp = 0; kp < 4; ++kp) {
      const uint32_t s = qit % QRING, ph = (qit / QRING) & 1;
      pf.mark(ring_phase);
      mbar_wait(&bars[B_QR_FULL + s], ph);
      tc_fence_after();
      pf.mark(PM_RINGWAIT);
      const uint64_t da = make_desc_sw128_kmajor(a0 + kp * 16384), db = make_desc_sw128_kmajor(qring + s * QSLOT_BYTES);
#pragma unroll
      for (int k = 0; k < 4; ++k) if (leader) umma_bf16(tmem + TM_D1, da + uint64_t(k * 2), db + uint64_t(k * 2), id_qkv, (kp | k) != 0 ? 1u : 0u);
      if (leader) umma_commit(&bars[B_QR_EMPTY + s]);
      ++qit;
    }
    if (leader) umma_commit(&bars[B_D1_FULL + (h & 1)]);
  };
  auto scores = [&](int h) {                  // S_h[128 x 64] = Qm0 K[0:64]^T + Qm1 K[64:128]^T   (compact: own sequence's keys only)
    const uint32_t d = tmem + TM_S + uint32_t(h & 1) * 64;
    const uint64_t da0 = desc_sw64(q0s), da1 = desc_sw64(q1s), db0 = desc_sw64(ks), db1 = desc_sw64(ks + 4096);
    if (leader) umma_bf16(d, da0, db0, id_s, 0u);
    if (leader) umma_bf16(d, da0 + 2, db0 + 2, id_s, 1u);
    if (leader) umma_bf16(d, da1, db1, id_s, 1u);
    if (leader) umma_bf16(d, da1 + 2, db1 + 2, id_s, 1u);
    if (leader) umma_commit(&bars[B_S_FULL + (h & 1)]);
  };
  auto pv = [&](int h) {                      // O[128 x 32] = P_h V_h   (A from TMEM: 128 keys = 64 columns, B MN-major)
    const uint32_t c = uint32_t(h & 1);
    for (int k = 0; k < 8; ++k)
      if (leader) umma_bf16_ts(tmem + TM_O, tmem + TM_S + c * 64 + uint32_t(k * 8), desc_sw64(vs + c * 8192 + k * 1024), id_pv, k != 0 ? 1u : 0u);
    if (leader) umma_commit(&bars[B_O_FULL + c]);
  };
  auto outproj = [&](int h) {                 // x += (O_h / l) Wout[:, 32h .. 32h+32)^T   (x + b_out was stored by the workers)
    const uint32_t s = RING_LO + h % (RING_HI - RING_LO), sb = slot_wait(s);
    const uint64_t da = desc_sw64(ost), db = desc_sw64(sb);
    if (leader) umma_bf16(tmem + TM_X, da, db, id_256, 1u);
    if (leader) umma_bf16(tmem + TM_X, da + 2, db + 2, id_256, 1u);
    slot_release(s);
  };
  auto wait_head = [&](int bar0, int h) {     // the event of head h on its chain's barrier: four completions per layer and chain
    mbar_wait(&bars[bar0 + (h & 1)], uint32_t(h >> 1) & 1u);
    tc_fence_after();
  };
  bool last_layer = false;
  auto ff1 = [&](int c) {                     // H[c&1][128 x 128] = LN2(x) W1[c*128.., :]^T
    const uint32_t d = tmem + ((c & 1) ? TM_H1 : TM_H0);
    for (int kp = 0; kp < 4; ++kp) {
      const uint32_t s = (m++) % RING, sb = slot_wait(s);
      const uint64_t da = make_desc_sw128_kmajor(a0 + kp * 16384), db = make_desc_sw128_kmajor(sb);
#pragma unroll
      for (int k = 0; k < 4; ++k) if (leader) umma_bf16(d, da + uint64_t(k * 2), db + uint64_t(k * 2), id_128, (kp | k) != 0 ? 1u : 0u);
      slot_release(s);
    }
    if (leader) umma_commit(&bars[B_HACC_FULL + (c & 1)]);
    if (c == a.n_chunks - 1 && last_layer) if (leader) umma_commit(&bars[B_A0_FREE]);
  };

  for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    for (int l = 0; l < a.depth; ++l) {
      last_layer = l == a.depth - 1;
      mbar_wait(&bars[B_A0_READY], (n_a0++) & 1);
      tc_fence_after();
      pf.mark(PM_WAIT_A0);
      ring_phase = PM_QKV;
      qkv(0);
      pf.mark(PM_QKV);
      for (int h = 0; h < HEADS; ++h) {
        wait_head(B_STAGED, h);
        pf.mark(PM_WAIT_STAGED);
        scores(h);
        pf.mark(PM_S);
        if (h + 1 < HEADS) qkv(h + 1);
        pf.mark(PM_QKV);
        if (h >= 1) {                         // the other chain's head h-1: its softmax ran next to the staging of head h
          wait_head(B_P_READY, h - 1);
          pf.mark(PM_WAIT_P);
          if (h >= 2) {                       // O(h-2) is out of TMEM and staged: its slice of the out-projection, then O is free for PV(h-1)
            wait_head(B_O_DRAINED, h - 2);
            pf.mark(PM_WAIT_OD7);
            ring_phase = PM_OUT;
            outproj(h - 2);
            pf.mark(PM_OUT);
            ring_phase = PM_QKV;
          }
          pv(h - 1);
          pf.mark(PM_PV);
        }
      }
      wait_head(B_P_READY, HEADS - 1);
      pf.mark(PM_WAIT_P);
      wait_head(B_O_DRAINED, HEADS - 2);
      pf.mark(PM_WAIT_OD7);
      ring_phase = PM_OUT;
      outproj(HEADS - 2);
      pf.mark(PM_OUT);
      pv(HEADS - 1);
      pf.mark(PM_PV);
      wait_head(B_O_DRAINED, HEADS - 1);
      pf.mark(PM_WAIT_OD7);
      outproj(HEADS - 1);
      if (leader) umma_commit(&bars[B_X1_FULL]);
      if (leader) umma_commit(&bars[B_QKV_FREE]);
      m = 0;
      pf.mark(PM_OUT);

      mbar_wait(&bars[B_A0_READY], (n_a0++) & 1);
      tc_fence_after();
      pf.mark(PM_WAIT_A0B);
      ring_phase = PM_FF1;
      ff1(0);
      if (a.n_chunks > 1) ff1(1);
      for (int c = 0; c < a.n_chunks; ++c) {
        const int b = c & 1;
        pf.mark(PM_FF1);
        mbar_wait(&bars[B_H_READY + b], (n_hready[b]++) & 1);
        tc_fence_after();
        pf.mark(PM_WAIT_H);
        ring_phase = PM_FF2;
        for (int kp = 0; kp < 2; ++kp)        // x += gelu(H_c) W2[:, c*128..]^T
          for (int nh = 0; nh < 2; ++nh) {
            const uint32_t s = (m++) % RING, sb = slot_wait(s);
            const uint32_t hb = tmem + (b ? TM_H1 : TM_H0);
            const uint64_t db = make_desc_sw128_kmajor(sb);
#pragma unroll
            for (int k = 0; k < 4; ++k) {     // 16-wide k-step kk = 4 kp + k of gelu(H_c): thread group kk / 2 packed it at column 32 (kk / 2) + 8 (kk % 2)
              const uint32_t kk = uint32_t(kp * 4 + k);
              if (leader) umma_bf16_ts(tmem + TM_X + nh * 128, hb + 32u * (kk >> 1) + 8u * (kk & 1u), db + uint64_t(k * 2), id_128, 1u);
            }
            slot_release(s);
          }
        pf.mark(PM_FF2);
        ring_phase = PM_FF1;
        if (c + 2 < a.n_chunks) ff1(c + 2);     // overwrites H_c: in order behind the MMAs above that read it
      }
      if (leader) umma_commit(&bars[B_X2_FULL]);
      if (leader) umma_commit(&bars[B_A1_FREE]);
    }
  }
  pf.flush(32, 64);
}

template <int IO>
__global__ void __launch_bounds__(NUM_THREADS, 1) sformer_fused_kernel(const __grid_constant__ FusedArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  if (threadIdx.x == 0 && (smem - smem_raw) + SMEM_USED > SMEM_ALLOC) {
    printf("avf: sformer_fused_kernel: dynamic shared memory base is not 1024-byte aligned enough\n");
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
      uint32_t count = 1;
      if (i == B_A0_READY || i == B_H_READY || i == B_H_READY1) count = NUM_WORKERS / 32;                                   // all worker warps
      if ((i >= B_STAGED && i < B_STAGED + 2) || (i >= B_P_READY && i < B_P_READY + 2) || (i >= B_O_DRAINED && i < B_O_DRAINED + 2))
        count = CHAIN_WARPS;                                                                                                // one chain's warps
      mbar_init(&bars[i], count);
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
  pdl_wait();        // the prologue above overlapped the previous kernel; its results are needed from here on
  pdl_trigger();

  if (warp == 0) {
    if (lane == 0) producer_main<IO>(a, smem, bars);
  } else if (warp == 1) {
    mma_main(a, smem, bars, tmem);
  } else if (warp == 2) {
    if (lane == 0) qkv_producer_main<IO>(a, smem, bars);
  } else {
    worker_main<IO>(a, smem, bars, tmem);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// pos [n_tok, 256] -> pos_t [256 / 4][POS_LD][4] (one tiny launch in front of the fused kernel; 64 KB)
__global__ void __launch_bounds__(256) pos_transpose2_kernel(const float* __restrict__ pos, float* __restrict__ pos_t, int n_tok) {
  pdl_wait();
  pdl_trigger();
  const int c = blockIdx.x * 4 + (threadIdx.x & 3), t = threadIdx.x >> 2;
  pos_t[(blockIdx.x * POS_LD + t) * 4 + (threadIdx.x & 3)] = t < n_tok ? pos[t * DIM + c] : 0.f;
}

}  // namespace

int sformer_fused_prof_read(unsigned long long* out64, int reset) {
#ifdef AVF_FUSED_PROF
  AVF_CUDA(cudaDeviceSynchronize());
  AVF_CUDA(cudaMemcpyFromSymbol(out64, g_prof, sizeof(unsigned long long) * 64));
  if (reset) {
    static unsigned long long zeros[64] = {0};
    AVF_CUDA(cudaMemcpyToSymbol(g_prof, zeros, sizeof(zeros)));
  }
  return 0;
#else
  (void)out64; (void)reset;
  return AVF_EUNSUPPORTED;
#endif
}

// Shapes this kernel takes over from avf_layer_fused.cu: the same stack family, sequences that need a 64-row slot.
bool sformer_fused_supported(const avf_stack_shape* s) {
  return s->dim == DIM && s->heads == HEADS && s->dim_head == DH && s->mlp_dim % 128 == 0 && s->mlp_dim >= 128 && s->mlp_dim <= MAX_MLP && s->n_tok > 32 &&
         s->n_tok <= 64 && s->depth >= 1 && s->depth <= MAX_DEPTH;
}

// io_kind 0: in/out are NCHW bf16 maps [n_seq, 256, n_tok] (pos required); 1: fp32 token rows with strides ld_in / ld_out.
int sformer_fused(int io_kind, const avf_stack_shape* s, const avf_layer_weights* L, const void* in, int ld_in, void* out, int ld_out,
                  const float* pos, void* scratch, cudaStream_t st) {
  AVF_REQUIRE(sformer_fused_supported(s), AVF_EUNSUPPORTED, "fused SFormer: unsupported shape dim=%d heads=%d dh=%d mlp=%d n_tok=%d depth=%d",
              s->dim, s->heads, s->dim_head, s->mlp_dim, s->n_tok, s->depth);
  AVF_REQUIRE(io_kind == IO_ROWS_F32 || (pos != nullptr && scratch != nullptr), AVF_EINVAL,
              "fused SFormer: NCHW input needs the positional embedding and the scratch buffer");
  AVF_REQUIRE(io_kind == IO_NCHW_BF16 || (ld_in % 4 == 0 && ld_out % 4 == 0), AVF_EINVAL, "fused SFormer: row strides must be multiples of 4");
  static thread_local FusedArgs a;      // ~2 KB of tensor maps + pointers, passed by value (__grid_constant__) per launch
  static_assert(sizeof(FusedArgs) <= 4000, "kernel parameter space");
  a.in = in; a.out = out; a.pos = pos; a.pos_t = static_cast<const float*>(scratch); a.ld_in = ld_in; a.ld_out = ld_out;
  if (io_kind == IO_NCHW_BF16) {
    launch_pdl(pos_transpose2_kernel, DIM / 4, 256, 0, st, pos, static_cast<float*>(scratch), s->n_tok);
    AVF_LAUNCH_CHECK("pos_transpose2_kernel");
  }
  a.n_seq = s->n_seq; a.n_tok = s->n_tok;
  a.n_tiles = ceil_div(s->n_seq, 2);
  a.n_chunks = s->mlp_dim / 128; a.depth = s->depth;
  const int inner = HEADS * DH;
  for (int l = 0; l < s->depth; ++l) {
    LayerArgs& A = a.layer[l];
    int e;
    if ((e = make_tmap_bf16_2d(&A.tm_qkv, L[l].w_qkv, 3 * inner, DIM, DIM, 32))) return e;
    if ((e = make_tmap_bf16_2d(&A.tm_out, L[l].w_out, DIM, inner, inner, 256, DH))) return e;
    if ((e = make_tmap_bf16_2d(&A.tm_w1, L[l].w_ff1, s->mlp_dim, DIM, DIM, 128))) return e;
    if ((e = make_tmap_bf16_2d(&A.tm_w2, L[l].w_ff2, DIM, s->mlp_dim, s->mlp_dim, 128))) return e;
    A.ln1_g = L[l].ln1_gamma; A.ln1_b = L[l].ln1_beta; A.b_out = L[l].b_out;
    A.ln2_g = L[l].ln2_gamma; A.ln2_b = L[l].ln2_beta; A.b_ff1 = L[l].b_ff1; A.b_ff2 = L[l].b_ff2;
  }
  static bool configured = false;
  if (!configured) {
    AVF_CUDA(cudaFuncSetAttribute(sformer_fused_kernel<IO_NCHW_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_ALLOC));
    AVF_CUDA(cudaFuncSetAttribute(sformer_fused_kernel<IO_ROWS_F32>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_ALLOC));
    configured = true;
  }
  int n_sm = 148, dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n_sm <= 0) n_sm = 148;
  const int cap = sm_cap();
  const int grid = min(a.n_tiles, cap > 0 ? min(cap, n_sm) : n_sm);
  if (io_kind == IO_NCHW_BF16)
    launch_pdl(sformer_fused_kernel<IO_NCHW_BF16>, grid, NUM_THREADS, SMEM_ALLOC, st, a);
  else
    launch_pdl(sformer_fused_kernel<IO_ROWS_F32>, grid, NUM_THREADS, SMEM_ALLOC, st, a);
  AVF_LAUNCH_CHECK("sformer_fused_kernel");
  return 0;
}

}  // namespace avf
