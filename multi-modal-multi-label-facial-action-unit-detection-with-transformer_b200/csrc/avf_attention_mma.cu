// Small-sequence attention on the tensor cores (bf16 in/out, fp32 softmax), models/heads.py:221-237.
//
// Sequences on this path are 9..49 tokens with 32- or 64-wide heads: one (sequence, head) problem is a
// handful of 16x8x16 MMAs, far below a 128-row tcgen05 tile, so each WARP owns one problem and keeps the
// whole score matrix in registers (flash-style: S never touches memory).  K and V of the head are staged in
// a per-warp padded shared-memory slab (conflict-free fragment loads / ldmatrix.trans); Q fragments come
// straight from global memory.  qkv [rows, 3*H*dh] with columns q|k|v head-major; out [rows, H*dh].
// (The 49-token SFormer additionally has a fully fused tcgen05 path; this kernel serves every stack.)
#include "avf_common.cuh"
#include "avf_internal.h"

namespace avf {

namespace {

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t& r0, uint32_t& r1, const void* smem_row_ptr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(smem_u32(smem_row_ptr)));
}

// 4 x 4 transpose of 32-bit words inside every quad of lanes: lane t enters with words 0..3 of ITS 16-byte chunk and leaves
// with word t of the chunks of lanes 0..3 (and back: the transpose is its own inverse).
__device__ __forceinline__ void quad_transpose32(uint32_t (&p)[4], int lane) {
  const bool b0 = lane & 1, b1 = lane & 2;
  uint32_t r;
  r = __shfl_xor_sync(0xffffffffu, b0 ? p[0] : p[1], 1); if (b0) p[0] = r; else p[1] = r;
  r = __shfl_xor_sync(0xffffffffu, b0 ? p[2] : p[3], 1); if (b0) p[2] = r; else p[3] = r;
  r = __shfl_xor_sync(0xffffffffu, b1 ? p[0] : p[2], 2); if (b1) p[0] = r; else p[2] = r;
  r = __shfl_xor_sync(0xffffffffu, b1 ? p[1] : p[3], 2); if (b1) p[1] = r; else p[3] = r;
}

template <int DH, int NP>    // NP = tokens padded to a multiple of 16 (16 / 32 / 48 / 64)
__global__ void __launch_bounds__(DH == 32 ? 128 : 64) attention_mma_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out,
                                                            int n_problems, int n_tok, int heads, float scale_log2e, int q_rows) {
  pdl_wait();      // programmatic dependent launch: everything above is launch overhead hidden behind the previous kernel
  pdl_trigger();
  // q_rows: only the first q_rows query rows of every sequence are computed and written (compactly: out [n_seq * q_rows, inner]);
  // q_rows = n_tok is plain self-attention, q_rows = 1 is the TFormer's last layer, of which only the cls row is consumed.
  constexpr int PITCH = DH + 8;                       // elements; (DH*2+16) bytes keeps 8 consecutive rows on distinct banks
  constexpr int WARPS = DH == 32 ? 4 : 2;             // keeps the static slab under 48 KB
  __shared__ __align__(16) __nv_bfloat16 smem[WARPS][2][NP][PITCH];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int prob = blockIdx.x * WARPS + warp;         // = seq * heads + head
  if (prob >= n_problems) return;
  const int seq = prob / heads, head = prob - seq * heads;
  const int inner = heads * DH;
  const size_t row0 = size_t(seq) * n_tok;
  const __nv_bfloat16* base = qkv + row0 * (3 * inner) + head * DH;
  __nv_bfloat16 (*Ks)[PITCH] = smem[warp][0];
  __nv_bfloat16 (*Vs)[PITCH] = smem[warp][1];

  // stage K and V (zero rows beyond n_tok: 0 * garbage must not become NaN in P*V)
  constexpr int CH = DH / 8;                          // 16-byte chunks per row
  // (all global loads are issued before the first shared-memory store: the warp has nothing else to hide their latency with)
  constexpr int NLD = NP * CH / 32;
  static_assert(NP * CH % 32 == 0, "whole warps of 16-byte chunks");
  uint4 kreg[NLD], vreg[NLD];
#pragma unroll
  for (int u = 0; u < NLD; ++u) {
    const int i = lane + u * 32, r = i / CH, c = i - r * CH;
    kreg[u] = make_uint4(0, 0, 0, 0);
    vreg[u] = make_uint4(0, 0, 0, 0);
    if (r < n_tok) {
      const __nv_bfloat16* p = base + size_t(r) * (3 * inner) + c * 8;
      kreg[u] = *reinterpret_cast<const uint4*>(p + inner);
      vreg[u] = *reinterpret_cast<const uint4*>(p + 2 * inner);
    }
  }
#pragma unroll
  for (int u = 0; u < NLD; ++u) {
    const int i = lane + u * 32, r = i / CH, c = i - r * CH;
    *reinterpret_cast<uint4*>(&Ks[r][c * 8]) = kreg[u];
    *reinterpret_cast<uint4*>(&Vs[r][c * 8]) = vreg[u];
  }
  __syncwarp();

  const int g = lane >> 2, t = lane & 3;
#pragma unroll 1
  for (int mt = 0; mt < NP / 16; ++mt) {
    if (mt * 16 >= q_rows) break;
    const int r_lo = mt * 16 + g, r_hi = r_lo + 8;
    // Q fragments (A operand, row-major 16x16 per k-step) from global: lane t of a quad fetches the 16-byte chunks t, t+4, ... of
    // its two rows (the quad reads 64 contiguous bytes per row and instruction instead of 4 x 4 bytes), then the quad transposes
    // 4-byte pieces so that every lane holds columns 2t, 2t+1 of each 8-column chunk — the m16n8k16 A layout.
    uint32_t qf[DH / 16][4];
#pragma unroll
    for (int grp = 0; grp < DH / 32; ++grp) {
      uint32_t lo[4] = {0u, 0u, 0u, 0u}, hi[4] = {0u, 0u, 0u, 0u};
      if (r_lo < n_tok) {
        const uint4 v = *reinterpret_cast<const uint4*>(base + size_t(r_lo) * (3 * inner) + (grp * 4 + t) * 8);
        lo[0] = v.x; lo[1] = v.y; lo[2] = v.z; lo[3] = v.w;
      }
      if (r_hi < n_tok) {
        const uint4 v = *reinterpret_cast<const uint4*>(base + size_t(r_hi) * (3 * inner) + (grp * 4 + t) * 8);
        hi[0] = v.x; hi[1] = v.y; hi[2] = v.z; hi[3] = v.w;
      }
      quad_transpose32(lo, lane);
      quad_transpose32(hi, lane);
#pragma unroll
      for (int j = 0; j < 4; ++j) {               // piece t of chunk grp*4 + j = columns (grp*4 + j)*8 + 2t, +1
        const int ks = (grp * 4 + j) >> 1, half = (grp * 4 + j) & 1;
        qf[ks][half * 2] = lo[j];
        qf[ks][half * 2 + 1] = hi[j];
      }
    }
    // S = Q K^T  (B fragment: B[k][n] = K[n][k] -> one 32-bit load per register)
    float s[NP / 8][4];
#pragma unroll
    for (int nt = 0; nt < NP / 8; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < DH / 16; ++ks) {
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(&Ks[nt * 8 + g][ks * 16 + 2 * t]);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(&Ks[nt * 8 + g][ks * 16 + 8 + 2 * t]);
        mma_bf16_16816(s[nt], qf[ks], b0, b1);
      }
    }
    // softmax over keys, rows r_lo (regs 0,1) and r_hi (regs 2,3); columns >= n_tok masked
    float m_lo = -INFINITY, m_hi = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < NP / 8; ++nt) {
      const int c = nt * 8 + 2 * t;
      if (c >= n_tok) { s[nt][0] = -INFINITY; s[nt][2] = -INFINITY; }
      if (c + 1 >= n_tok) { s[nt][1] = -INFINITY; s[nt][3] = -INFINITY; }
      m_lo = fmaxf(m_lo, fmaxf(s[nt][0], s[nt][1]));
      m_hi = fmaxf(m_hi, fmaxf(s[nt][2], s[nt][3]));
    }
    m_lo = fmaxf(m_lo, __shfl_xor_sync(0xffffffffu, m_lo, 1));
    m_lo = fmaxf(m_lo, __shfl_xor_sync(0xffffffffu, m_lo, 2));
    m_hi = fmaxf(m_hi, __shfl_xor_sync(0xffffffffu, m_hi, 1));
    m_hi = fmaxf(m_hi, __shfl_xor_sync(0xffffffffu, m_hi, 2));
    const float o_lo = m_lo * scale_log2e, o_hi = m_hi * scale_log2e;
    float l_lo = 0.f, l_hi = 0.f;
#pragma unroll
    for (int nt = 0; nt < NP / 8; ++nt) {
      s[nt][0] = exp2f(fmaf(s[nt][0], scale_log2e, -o_lo));
      s[nt][1] = exp2f(fmaf(s[nt][1], scale_log2e, -o_lo));
      s[nt][2] = exp2f(fmaf(s[nt][2], scale_log2e, -o_hi));
      s[nt][3] = exp2f(fmaf(s[nt][3], scale_log2e, -o_hi));
      l_lo += s[nt][0] + s[nt][1];
      l_hi += s[nt][2] + s[nt][3];
    }
    l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1);
    l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
    l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1);
    l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
    // O = P V : the C fragments of two adjacent score tiles are exactly one A fragment
    float o[DH / 8][4];
#pragma unroll
    for (int nt = 0; nt < DH / 8; ++nt) o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f;
#pragma unroll
    for (int kt = 0; kt < NP / 16; ++kt) {
      uint32_t pa[4];
      pa[0] = pack_bf16x2(s[2 * kt][0], s[2 * kt][1]);
      pa[1] = pack_bf16x2(s[2 * kt][2], s[2 * kt][3]);
      pa[2] = pack_bf16x2(s[2 * kt + 1][0], s[2 * kt + 1][1]);
      pa[3] = pack_bf16x2(s[2 * kt + 1][2], s[2 * kt + 1][3]);
#pragma unroll
      for (int nt = 0; nt < DH / 8; ++nt) {
        uint32_t b0, b1;
        ldmatrix_x2_trans(b0, b1, &Vs[kt * 16 + (lane & 15)][nt * 8]);
        mma_bf16_16816(o[nt], pa, b0, b1);
      }
    }
    const float i_lo = 1.f / l_lo, i_hi = 1.f / l_hi;
    __nv_bfloat16* olo = out + (size_t(seq) * q_rows + r_lo) * inner + head * DH;
    __nv_bfloat16* ohi = out + (size_t(seq) * q_rows + r_hi) * inner + head * DH;
#pragma unroll
    for (int grp = 0; grp < DH / 32; ++grp) {     // the same transpose backwards: lane t stores the whole 16-byte chunk grp*4 + t
      uint32_t lo[4], hi[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        lo[j] = pack_bf16x2(o[grp * 4 + j][0] * i_lo, o[grp * 4 + j][1] * i_lo);
        hi[j] = pack_bf16x2(o[grp * 4 + j][2] * i_hi, o[grp * 4 + j][3] * i_hi);
      }
      quad_transpose32(lo, lane);
      quad_transpose32(hi, lane);
      if (r_lo < q_rows) *reinterpret_cast<uint4*>(olo + (grp * 4 + t) * 8) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      if (r_hi < q_rows) *reinterpret_cast<uint4*>(ohi + (grp * 4 + t) * 8) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    }
  }
}

template <int DH, int NP>
int launch(const void* qkv, void* out, int n_seq, int n_tok, int heads, int q_rows, cudaStream_t st) {
  const int n_problems = n_seq * heads;
  const float scale_log2e = 1.4426950408889634f / sqrtf(float(DH));
  constexpr int kWarps = DH == 32 ? 4 : 2;
  launch_pdl(attention_mma_kernel<DH, NP>, ceil_div(n_problems, kWarps), kWarps * 32, 0, st, static_cast<const __nv_bfloat16*>(qkv), static_cast<__nv_bfloat16*>(out),
                                                                      n_problems, n_tok, heads, scale_log2e, q_rows);
  AVF_LAUNCH_CHECK("attention_mma_kernel");
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Backward (training): dQ, dK, dV of one (sequence, head) per warp, all five contractions on mma.sync.
//   pass 1 (rows = queries): S = Q K^T -> P, dP = dO V^T, delta = rowsum(P o dP), dS = P o (dP - delta) / sqrt(dh),
//                            dQ = dS K; the softmax statistics (max, 1/sum) and delta go to shared memory;
//   pass 2 (rows = keys):    S^T = K Q^T -> P^T from the saved statistics, dP^T = V dO^T, dS^T likewise,
//                            dV = P^T dO,  dK = dS^T Q.
// Recomputing the scores in the transposed orientation keeps every A operand a row-major fragment straight out of the
// accumulator registers (no transposes through shared memory).  Q, K, V, dO of the head sit in per-warp padded slabs.
// ---------------------------------------------------------------------------------------------
template <int DH, int NP>
struct BwdCfg {
  static constexpr int PITCH = DH + 8;
  static constexpr int SLAB_BYTES = 4 * NP * PITCH * 2 + 3 * NP * 4;
  static constexpr int WARPS = SLAB_BYTES * 2 <= 46 * 1024 ? 2 : 1;
};

template <int DH, int NP>
__global__ void __launch_bounds__(BwdCfg<DH, NP>::WARPS * 32) attention_bwd_mma_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                                                       const __nv_bfloat16* __restrict__ dout,
                                                                                       __nv_bfloat16* __restrict__ dqkv, int n_problems,
                                                                                       int n_tok, int heads, float scale) {
  pdl_wait();      // programmatic dependent launch: everything above is launch overhead hidden behind the previous kernel
  pdl_trigger();
  using Cfg = BwdCfg<DH, NP>;
  constexpr int PITCH = Cfg::PITCH, WARPS = Cfg::WARPS;
  __shared__ __align__(16) __nv_bfloat16 smem[WARPS][4][NP][PITCH];
  __shared__ float stats[WARPS][3][NP];               // row max (in exp2 units), 1 / row sum, delta
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int prob = blockIdx.x * WARPS + warp;
  if (prob >= n_problems) return;
  const int seq = prob / heads, head = prob - seq * heads;
  const int inner = heads * DH;
  const size_t row0 = size_t(seq) * n_tok;
  const __nv_bfloat16* base = qkv + row0 * (3 * inner) + head * DH;
  const __nv_bfloat16* dobase = dout + row0 * inner + head * DH;
  __nv_bfloat16* dbase = dqkv + row0 * (3 * inner) + head * DH;
  __nv_bfloat16 (*Qs)[PITCH] = smem[warp][0];
  __nv_bfloat16 (*Ks)[PITCH] = smem[warp][1];
  __nv_bfloat16 (*Vs)[PITCH] = smem[warp][2];
  __nv_bfloat16 (*Os)[PITCH] = smem[warp][3];
  float* st_m = stats[warp][0];
  float* st_il = stats[warp][1];
  float* st_d = stats[warp][2];
  const float scale_log2e = scale * 1.4426950408889634f;

  constexpr int CH = DH / 8;
  for (int i = lane; i < NP * CH; i += 32) {
    const int r = i / CH, c = i - r * CH;
    uint4 q = make_uint4(0, 0, 0, 0), k = q, v = q, o = q;
    if (r < n_tok) {
      const __nv_bfloat16* p = base + size_t(r) * (3 * inner) + c * 8;
      q = *reinterpret_cast<const uint4*>(p);
      k = *reinterpret_cast<const uint4*>(p + inner);
      v = *reinterpret_cast<const uint4*>(p + 2 * inner);
      o = *reinterpret_cast<const uint4*>(dobase + size_t(r) * inner + c * 8);
    }
    *reinterpret_cast<uint4*>(&Qs[r][c * 8]) = q;
    *reinterpret_cast<uint4*>(&Ks[r][c * 8]) = k;
    *reinterpret_cast<uint4*>(&Vs[r][c * 8]) = v;
    *reinterpret_cast<uint4*>(&Os[r][c * 8]) = o;
  }
  __syncwarp();

  const int g = lane >> 2, t = lane & 3;
  auto load_a = [&](uint32_t (&f)[DH / 16][4], __nv_bfloat16 (*M)[PITCH], int r_lo) {
#pragma unroll
    for (int ks = 0; ks < DH / 16; ++ks) {
      f[ks][0] = *reinterpret_cast<const uint32_t*>(&M[r_lo][ks * 16 + 2 * t]);
      f[ks][1] = *reinterpret_cast<const uint32_t*>(&M[r_lo + 8][ks * 16 + 2 * t]);
      f[ks][2] = *reinterpret_cast<const uint32_t*>(&M[r_lo][ks * 16 + 8 + 2 * t]);
      f[ks][3] = *reinterpret_cast<const uint32_t*>(&M[r_lo + 8][ks * 16 + 8 + 2 * t]);
    }
  };
  // C[16 x NP] = A (fragments of 16 rows) * M^T  with M [NP x DH] row-major in shared memory
  auto mma_abt = [&](float (&c)[NP / 8][4], const uint32_t (&a)[DH / 16][4], __nv_bfloat16 (*M)[PITCH]) {
#pragma unroll
    for (int nt = 0; nt < NP / 8; ++nt) {
      c[nt][0] = c[nt][1] = c[nt][2] = c[nt][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < DH / 16; ++ks) {
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(&M[nt * 8 + g][ks * 16 + 2 * t]);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(&M[nt * 8 + g][ks * 16 + 8 + 2 * t]);
        mma_bf16_16816(c[nt], a[ks], b0, b1);
      }
    }
  };
  // D[16 x DH] = X[16 x NP] (fp32 accumulator-layout registers, rounded to bf16) * M [NP x DH]
  auto mma_xm = [&](float (&d)[DH / 8][4], const float (&x)[NP / 8][4], __nv_bfloat16 (*M)[PITCH]) {
#pragma unroll
    for (int nt = 0; nt < DH / 8; ++nt) d[nt][0] = d[nt][1] = d[nt][2] = d[nt][3] = 0.f;
#pragma unroll
    for (int kt = 0; kt < NP / 16; ++kt) {
      uint32_t pa[4];
      pa[0] = pack_bf16x2(x[2 * kt][0], x[2 * kt][1]);
      pa[1] = pack_bf16x2(x[2 * kt][2], x[2 * kt][3]);
      pa[2] = pack_bf16x2(x[2 * kt + 1][0], x[2 * kt + 1][1]);
      pa[3] = pack_bf16x2(x[2 * kt + 1][2], x[2 * kt + 1][3]);
#pragma unroll
      for (int nt = 0; nt < DH / 8; ++nt) {
        uint32_t b0, b1;
        ldmatrix_x2_trans(b0, b1, &M[kt * 16 + (lane & 15)][nt * 8]);
        mma_bf16_16816(d[nt], pa, b0, b1);
      }
    }
  };
  auto store_rows = [&](const float (&d)[DH / 8][4], __nv_bfloat16* dst_col0, int r_lo, float mul) {
    __nv_bfloat16* lo = dst_col0 + size_t(r_lo) * (3 * inner) + 2 * t;
    __nv_bfloat16* hi = lo + size_t(8) * (3 * inner);
#pragma unroll
    for (int nt = 0; nt < DH / 8; ++nt) {
      if (r_lo < n_tok) *reinterpret_cast<uint32_t*>(lo + nt * 8) = pack_bf16x2(d[nt][0] * mul, d[nt][1] * mul);
      if (r_lo + 8 < n_tok) *reinterpret_cast<uint32_t*>(hi + nt * 8) = pack_bf16x2(d[nt][2] * mul, d[nt][3] * mul);
    }
  };

  // ---- pass 1: rows = queries ------------------------------------------------------------------------------
#pragma unroll 1
  for (int mt = 0; mt < NP / 16; ++mt) {
    if (mt * 16 >= n_tok) break;
    const int r_lo = mt * 16 + g;
    uint32_t af[DH / 16][4];
    float s[NP / 8][4], dp[NP / 8][4];
    load_a(af, Qs, r_lo);
    mma_abt(s, af, Ks);
    load_a(af, Os, r_lo);
    mma_abt(dp, af, Vs);
    float m_lo = -INFINITY, m_hi = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < NP / 8; ++nt) {
      const int c = nt * 8 + 2 * t;
      if (c >= n_tok) { s[nt][0] = -INFINITY; s[nt][2] = -INFINITY; }
      if (c + 1 >= n_tok) { s[nt][1] = -INFINITY; s[nt][3] = -INFINITY; }
      m_lo = fmaxf(m_lo, fmaxf(s[nt][0], s[nt][1]));
      m_hi = fmaxf(m_hi, fmaxf(s[nt][2], s[nt][3]));
    }
    m_lo = fmaxf(m_lo, __shfl_xor_sync(0xffffffffu, m_lo, 1));
    m_lo = fmaxf(m_lo, __shfl_xor_sync(0xffffffffu, m_lo, 2));
    m_hi = fmaxf(m_hi, __shfl_xor_sync(0xffffffffu, m_hi, 1));
    m_hi = fmaxf(m_hi, __shfl_xor_sync(0xffffffffu, m_hi, 2));
    const float o_lo = m_lo * scale_log2e, o_hi = m_hi * scale_log2e;
    float l_lo = 0.f, l_hi = 0.f;
#pragma unroll
    for (int nt = 0; nt < NP / 8; ++nt) {
      s[nt][0] = exp2f(fmaf(s[nt][0], scale_log2e, -o_lo));
      s[nt][1] = exp2f(fmaf(s[nt][1], scale_log2e, -o_lo));
      s[nt][2] = exp2f(fmaf(s[nt][2], scale_log2e, -o_hi));
      s[nt][3] = exp2f(fmaf(s[nt][3], scale_log2e, -o_hi));
      l_lo += s[nt][0] + s[nt][1];
      l_hi += s[nt][2] + s[nt][3];
    }
    l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1);
    l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
    l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1);
    l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
    const float i_lo = 1.f / l_lo, i_hi = 1.f / l_hi;
    float d_lo = 0.f, d_hi = 0.f;
#pragma unroll
    for (int nt = 0; nt < NP / 8; ++nt) {
      s[nt][0] *= i_lo; s[nt][1] *= i_lo; s[nt][2] *= i_hi; s[nt][3] *= i_hi;
      d_lo += s[nt][0] * dp[nt][0] + s[nt][1] * dp[nt][1];
      d_hi += s[nt][2] * dp[nt][2] + s[nt][3] * dp[nt][3];
    }
    d_lo += __shfl_xor_sync(0xffffffffu, d_lo, 1);
    d_lo += __shfl_xor_sync(0xffffffffu, d_lo, 2);
    d_hi += __shfl_xor_sync(0xffffffffu, d_hi, 1);
    d_hi += __shfl_xor_sync(0xffffffffu, d_hi, 2);
    if (t == 0) {
      st_m[r_lo] = o_lo; st_il[r_lo] = i_lo; st_d[r_lo] = d_lo;
      st_m[r_lo + 8] = o_hi; st_il[r_lo + 8] = i_hi; st_d[r_lo + 8] = d_hi;
    }
#pragma unroll
    for (int nt = 0; nt < NP / 8; ++nt) {      // dS (the 1/sqrt(dh) is applied once when dQ is stored)
      s[nt][0] *= dp[nt][0] - d_lo; s[nt][1] *= dp[nt][1] - d_lo;
      s[nt][2] *= dp[nt][2] - d_hi; s[nt][3] *= dp[nt][3] - d_hi;
    }
    float dq[DH / 8][4];
    mma_xm(dq, s, Ks);
    store_rows(dq, dbase, r_lo, scale);
  }
  __syncwarp();

  // ---- pass 2: rows = keys ------------------------------------------------------------------------------------
#pragma unroll 1
  for (int mt = 0; mt < NP / 16; ++mt) {
    if (mt * 16 >= n_tok) break;
    const int r_lo = mt * 16 + g;
    uint32_t af[DH / 16][4];
    float p[NP / 8][4], dp[NP / 8][4];
    load_a(af, Ks, r_lo);
    mma_abt(p, af, Qs);                       // S^T[key][query]
    load_a(af, Vs, r_lo);
    mma_abt(dp, af, Os);                      // dP^T[key][query]
    float ds[NP / 8][4];
#pragma unroll
    for (int nt = 0; nt < NP / 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int qi = nt * 8 + 2 * t + e;    // query index of elements e (row r_lo) and e + 2 (row r_lo + 8)
        const bool ok = qi < n_tok;
        const float m = ok ? st_m[qi] : 0.f, il = ok ? st_il[qi] : 0.f, dl = ok ? st_d[qi] : 0.f;
        const float p0 = ok ? exp2f(fmaf(p[nt][e], scale_log2e, -m)) * il : 0.f;
        const float p1 = ok ? exp2f(fmaf(p[nt][e + 2], scale_log2e, -m)) * il : 0.f;
        p[nt][e] = p0;
        p[nt][e + 2] = p1;
        ds[nt][e] = p0 * (dp[nt][e] - dl);
        ds[nt][e + 2] = p1 * (dp[nt][e + 2] - dl);
      }
    }
    float acc[DH / 8][4];
    mma_xm(acc, p, Os);                       // dV = P^T dO
    store_rows(acc, dbase + 2 * inner, r_lo, 1.f);
    mma_xm(acc, ds, Qs);                      // dK = dS^T Q / sqrt(dh)
    store_rows(acc, dbase + inner, r_lo, scale);
  }
}

template <int DH, int NP>
int launch_bwd(const void* qkv, const void* dout, void* dqkv, int n_seq, int n_tok, int heads, cudaStream_t st) {
  const int n_problems = n_seq * heads;
  constexpr int kWarps = BwdCfg<DH, NP>::WARPS;
  launch_pdl(attention_bwd_mma_kernel<DH, NP>, ceil_div(n_problems, kWarps), kWarps * 32, 0, st, static_cast<const __nv_bfloat16*>(qkv), static_cast<const __nv_bfloat16*>(dout), static_cast<__nv_bfloat16*>(dqkv), n_problems, n_tok, heads,
      1.0f / sqrtf(float(DH)));
  AVF_LAUNCH_CHECK("attention_bwd_mma_kernel");
  return 0;
}

}  // namespace

int attention_mma_bf16(const void* qkv, void* out, int n_seq, int n_tok, int heads, int dim_head, cudaStream_t st, int q_rows) {
  AVF_REQUIRE(n_tok >= 1 && n_tok <= 64, AVF_EUNSUPPORTED, "attention: n_tok=%d (1..64)", n_tok);
  if (q_rows <= 0 || q_rows > n_tok) q_rows = n_tok;
  const int np = (n_tok + 15) / 16 * 16;
#define AVF_ATT(D, P) if (dim_head == D && np == P) return launch<D, P>(qkv, out, n_seq, n_tok, heads, q_rows, st);
  AVF_ATT(32, 16) AVF_ATT(32, 32) AVF_ATT(32, 48) AVF_ATT(32, 64)
  AVF_ATT(64, 16) AVF_ATT(64, 32) AVF_ATT(64, 48) AVF_ATT(64, 64)
#undef AVF_ATT
  AVF_REQUIRE(false, AVF_EUNSUPPORTED, "attention: dim_head=%d (supported: 32, 64)", dim_head);
  return 0;
}

}  // namespace avf

namespace avf {
int attention_bwd_mma_bf16(const void* qkv, const void* dout, void* dqkv, int n_seq, int n_tok, int heads, int dim_head, cudaStream_t st) {
  AVF_REQUIRE(n_tok >= 1 && n_tok <= 64, AVF_EUNSUPPORTED, "attention_bwd: n_tok=%d (1..64)", n_tok);
  const int np = (n_tok + 15) / 16 * 16;
#define AVF_ATT(D, P) if (dim_head == D && np == P) return launch_bwd<D, P>(qkv, dout, dqkv, n_seq, n_tok, heads, st);
  AVF_ATT(32, 16) AVF_ATT(32, 32) AVF_ATT(32, 48) AVF_ATT(32, 64)
  AVF_ATT(64, 16) AVF_ATT(64, 32) AVF_ATT(64, 48) AVF_ATT(64, 64)
#undef AVF_ATT
  AVF_REQUIRE(false, AVF_EUNSUPPORTED, "attention_bwd: dim_head=%d (supported: 32, 64)", dim_head);
  return 0;
}
}  // namespace avf
