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

template <int DH, int NP>    // NP = tokens padded to a multiple of 16 (16 / 32 / 48 / 64)
__global__ void __launch_bounds__(DH == 32 ? 128 : 64) attention_mma_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out,
                                                            int n_problems, int n_tok, int heads, float scale_log2e) {
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
  for (int i = lane; i < NP * CH; i += 32) {
    const int r = i / CH, c = i - r * CH;
    uint4 kv = make_uint4(0, 0, 0, 0), vv = make_uint4(0, 0, 0, 0);
    if (r < n_tok) {
      const __nv_bfloat16* p = base + size_t(r) * (3 * inner) + c * 8;
      kv = *reinterpret_cast<const uint4*>(p + inner);
      vv = *reinterpret_cast<const uint4*>(p + 2 * inner);
    }
    *reinterpret_cast<uint4*>(&Ks[r][c * 8]) = kv;
    *reinterpret_cast<uint4*>(&Vs[r][c * 8]) = vv;
  }
  __syncwarp();

  const int g = lane >> 2, t = lane & 3;
#pragma unroll 1
  for (int mt = 0; mt < NP / 16; ++mt) {
    if (mt * 16 >= n_tok) break;
    const int r_lo = mt * 16 + g, r_hi = r_lo + 8;
    // Q fragments (A operand, row-major 16x16 per k-step) straight from global
    uint32_t qf[DH / 16][4];
#pragma unroll
    for (int ks = 0; ks < DH / 16; ++ks) {
      const __nv_bfloat16* qlo = base + size_t(r_lo) * (3 * inner) + ks * 16 + 2 * t;
      const __nv_bfloat16* qhi = base + size_t(r_hi) * (3 * inner) + ks * 16 + 2 * t;
      qf[ks][0] = r_lo < n_tok ? *reinterpret_cast<const uint32_t*>(qlo) : 0u;
      qf[ks][1] = r_hi < n_tok ? *reinterpret_cast<const uint32_t*>(qhi) : 0u;
      qf[ks][2] = r_lo < n_tok ? *reinterpret_cast<const uint32_t*>(qlo + 8) : 0u;
      qf[ks][3] = r_hi < n_tok ? *reinterpret_cast<const uint32_t*>(qhi + 8) : 0u;
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
    __nv_bfloat16* olo = out + (row0 + r_lo) * inner + head * DH + 2 * t;
    __nv_bfloat16* ohi = out + (row0 + r_hi) * inner + head * DH + 2 * t;
#pragma unroll
    for (int nt = 0; nt < DH / 8; ++nt) {
      if (r_lo < n_tok) *reinterpret_cast<uint32_t*>(olo + nt * 8) = pack_bf16x2(o[nt][0] * i_lo, o[nt][1] * i_lo);
      if (r_hi < n_tok) *reinterpret_cast<uint32_t*>(ohi + nt * 8) = pack_bf16x2(o[nt][2] * i_hi, o[nt][3] * i_hi);
    }
  }
}

template <int DH, int NP>
int launch(const void* qkv, void* out, int n_seq, int n_tok, int heads, cudaStream_t st) {
  const int n_problems = n_seq * heads;
  const float scale_log2e = 1.4426950408889634f / sqrtf(float(DH));
  constexpr int kWarps = DH == 32 ? 4 : 2;
  attention_mma_kernel<DH, NP><<<ceil_div(n_problems, kWarps), kWarps * 32, 0, st>>>(static_cast<const __nv_bfloat16*>(qkv), static_cast<__nv_bfloat16*>(out),
                                                                      n_problems, n_tok, heads, scale_log2e);
  AVF_LAUNCH_CHECK("attention_mma_kernel");
  return 0;
}

}  // namespace

int attention_mma_bf16(const void* qkv, void* out, int n_seq, int n_tok, int heads, int dim_head, cudaStream_t st) {
  AVF_REQUIRE(n_tok >= 1 && n_tok <= 64, AVF_EUNSUPPORTED, "attention: n_tok=%d (1..64)", n_tok);
  const int np = (n_tok + 15) / 16 * 16;
#define AVF_ATT(D, P) if (dim_head == D && np == P) return launch<D, P>(qkv, out, n_seq, n_tok, heads, st);
  AVF_ATT(32, 16) AVF_ATT(32, 32) AVF_ATT(32, 48) AVF_ATT(32, 64)
  AVF_ATT(64, 16) AVF_ATT(64, 32) AVF_ATT(64, 48) AVF_ATT(64, 64)
#undef AVF_ATT
  AVF_REQUIRE(false, AVF_EUNSUPPORTED, "attention: dim_head=%d (supported: 32, 64)", dim_head);
  return 0;
}

}  // namespace avf
