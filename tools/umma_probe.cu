// tools/umma_probe.cu — stand-alone check of the tcgen05 operand layouts the fused encoder-layer kernel relies on.
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/umma_probe.bin tools/umma_probe.cu
// Each test stages operands in shared memory (or TMEM) by hand, issues tcgen05.mma and compares D with a CPU product.
//
//   T1  A K-major SW128 [128x64], B K-major SW128 [96x64]                   (N=96 shape, baseline layout)
//   T2  A K-major SW64  [128x32], B K-major SW64  [128x32]                  (S = Q K^T per head)
//   T3  A from TMEM (bf16 pairs), B MN-major SW64 [K=112 x N=32]            (O = P V per head)
//   T4  A K-major SW128 [128x64], B K-major SW128 [256x64] accumulated onto a TMEM tile pre-filled by tcgen05.st
//   T5  A from TMEM, B K-major SW128 [N=32 x K=128] (V^T fallback)
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../multi-modal-multi-label-facial-action-unit-detection-with-transformer_b200/csrc/avf_common.cuh"

namespace avf {
void set_error(const char*, ...) {}
int check_cuda(cudaError_t e, const char*) { return int(e); }
void count_launch() {}
}  // namespace avf
using namespace avf;

enum { A_SMEM_SW128 = 0, A_SMEM_SW64 = 1, A_TMEM = 2, A_MN_SW128 = 3 };
enum { B_K_SW128 = 0, B_K_SW64 = 1, B_MN_SW64 = 2, B_MN_SW128 = 3 };

struct Params {
  int a_mode, b_mode, n, k;     // m = 128
  int prefill;                  // 1: D pre-filled with `init` through tcgen05.st and MMA accumulates
  uint32_t b_lbo, b_sbo;        // descriptor byte offsets for B
  uint32_t a_lbo, a_sbo;        // descriptor byte offsets for A (A_MN_SW128)
};

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

// a: [128, k] row-major bf16.  b: K-major modes: [n, k] row-major; MN-major modes: [k, n] row-major.  d: [128, n] fp32.
__global__ void __launch_bounds__(128) probe_kernel(const __nv_bfloat16* a, const __nv_bfloat16* b, const float* init, float* d, Params p) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = smem;                 // up to 64 KB
  uint8_t* sb = smem + 65536;         // up to 64 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 131072);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *slot;
  const uint32_t TM_D = tm, TM_A = tm + 256;

  // ---- stage A ----
  if (p.a_mode == A_SMEM_SW128) {          // panels of 64 columns, [128 rows x 128 B]
    for (int i = tid; i < 128 * p.k; i += 128) {
      const int r = i / p.k, c = i % p.k;
      *reinterpret_cast<__nv_bfloat16*>(sa + sw128_offset(r, c, 128)) = a[i];
    }
  } else if (p.a_mode == A_SMEM_SW64) {    // k == 32: rows of 64 B, chunk ^= (row >> 1) & 3
    for (int i = tid; i < 128 * p.k; i += 128) {
      const int r = i / p.k, c = i % p.k;
      const uint32_t off = uint32_t(r) * 64u + (uint32_t((c >> 3) ^ ((r >> 1) & 3)) << 4) + uint32_t(c & 7) * 2u;
      *reinterpret_cast<__nv_bfloat16*>(sa + off) = a[i];
    }
  } else if (p.a_mode == A_MN_SW128) {     // [k rows x 128 m] as two 64-wide panels: row = K index, 128 B of M
    for (int i = tid; i < 128 * p.k; i += 128) {
      const int r = i / p.k, c = i % p.k;   // a[r][c]: r = m, c = k
      *reinterpret_cast<__nv_bfloat16*>(sa + sw128_offset(c, r, p.k)) = a[i];
    }
  } else {                                  // TMEM: lane = row, column j holds elements (2j, 2j+1)
    const int r = tid;
    for (int c0 = 0; c0 < p.k; c0 += 16) {
      uint32_t v[8];
      for (int j = 0; j < 8; ++j) {
        __nv_bfloat162 t2 = __halves2bfloat162(a[r * p.k + c0 + 2 * j], a[r * p.k + c0 + 2 * j + 1]);
        v[j] = *reinterpret_cast<uint32_t*>(&t2);
      }
      tmem_st8(TM_A + (uint32_t(warp * 32) << 16) + uint32_t(c0 / 2), v);
    }
    tmem_st_wait();
  }
  // ---- stage B ----
  if (p.b_mode == B_K_SW128) {
    for (int i = tid; i < p.n * p.k; i += 128) {
      const int r = i / p.k, c = i % p.k;
      *reinterpret_cast<__nv_bfloat16*>(sb + sw128_offset(r, c, p.n)) = b[i];
    }
  } else if (p.b_mode == B_K_SW64) {
    for (int i = tid; i < p.n * p.k; i += 128) {
      const int r = i / p.k, c = i % p.k;
      const uint32_t off = uint32_t(r) * 64u + (uint32_t((c >> 3) ^ ((r >> 1) & 3)) << 4) + uint32_t(c & 7) * 2u;
      *reinterpret_cast<__nv_bfloat16*>(sb + off) = b[i];
    }
  } else if (p.b_mode == B_MN_SW64) {      // [k rows x n=32] : row = K index, 64 B of N
    for (int i = tid; i < p.k * p.n; i += 128) {
      const int r = i / p.n, c = i % p.n;
      const uint32_t off = uint32_t(r) * 64u + (uint32_t((c >> 3) ^ ((r >> 1) & 3)) << 4) + uint32_t(c & 7) * 2u;
      *reinterpret_cast<__nv_bfloat16*>(sb + off) = b[i];
    }
  } else {                                  // B_MN_SW128: [k rows x n=64]: row = K index, 128 B of N
    for (int i = tid; i < p.k * p.n; i += 128) {
      const int r = i / p.n, c = i % p.n;
      *reinterpret_cast<__nv_bfloat16*>(sb + sw128_offset(r, c, p.k)) = b[i];
    }
  }
  if (p.prefill) {
    const int r = tid;
    for (int c0 = 0; c0 < p.n; c0 += 8) {
      uint32_t v[8];
      for (int j = 0; j < 8; ++j) v[j] = __float_as_uint(init[r * p.n + c0 + j]);
      tmem_st8(TM_D + (uint32_t(warp * 32) << 16) + uint32_t(c0), v);
    }
    tmem_st_wait();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (tid == 0) {
    const bool b_mn = p.b_mode >= B_MN_SW64;
    const uint32_t idesc = make_idesc_bf16(128, p.n, p.a_mode == A_MN_SW128 ? 1 : 0, b_mn ? 1 : 0);
    for (int ks = 0; ks < p.k / 16; ++ks) {
      uint64_t db;
      if (p.b_mode == B_K_SW128) {
        db = make_desc_sw128_kmajor(smem_u32(sb) + (ks / 4) * p.n * 128) + uint64_t((ks % 4) * 2);
      } else if (p.b_mode == B_K_SW64) {
        db = make_desc(smem_u32(sb), 16, 512, 4) + uint64_t(ks * 2);
      } else if (p.b_mode == B_MN_SW64) {   // advance 16 K-rows = 1024 B per step
        db = make_desc(smem_u32(sb) + ks * 1024, p.b_lbo, p.b_sbo, 4);
      } else {
        db = make_desc(smem_u32(sb) + ks * 2048, p.b_lbo, p.b_sbo, 2);
      }
      const uint32_t acc = (ks > 0 || p.prefill) ? 1u : 0u;
      if (p.a_mode == A_TMEM) {
        umma_bf16_ts(TM_D, TM_A + uint32_t(ks * 8), db, idesc, acc);
      } else {
        uint64_t da;
        if (p.a_mode == A_MN_SW128) da = make_desc(smem_u32(sa) + ks * 2048, p.a_lbo, p.a_sbo, 2);
        else if (p.a_mode == A_SMEM_SW128) da = make_desc_sw128_kmajor(smem_u32(sa) + (ks / 4) * 128 * 128) + uint64_t((ks % 4) * 2);
        else da = make_desc(smem_u32(sa), 16, 512, 4) + uint64_t(ks * 2);
        umma_bf16(TM_D, da, db, idesc, acc);
      }
    }
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < p.n; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(TM_D + (uint32_t(warp * 32) << 16) + uint32_t(c0), v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) d[tid * p.n + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

static bool run(const char* name, Params p) {
  const int m = 128;
  std::vector<float> A(m * p.k), B(p.n * p.k), I(m * p.n, 0.f), ref(m * p.n);
  srand(1234);
  for (auto& x : A) x = bf((rand() % 2001 - 1000) / 500.f);
  for (auto& x : B) x = bf((rand() % 2001 - 1000) / 500.f);     // logical B[n][k]
  for (auto& x : I) x = p.prefill ? (rand() % 2001 - 1000) / 100.f : 0.f;
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < p.n; ++j) {
      double s = I[i * p.n + j];
      for (int k = 0; k < p.k; ++k) s += double(A[i * p.k + k]) * B[j * p.k + k];
      ref[i * p.n + j] = float(s);
    }
  std::vector<__nv_bfloat16> hA(m * p.k), hB(p.n * p.k);
  for (int i = 0; i < m * p.k; ++i) hA[i] = __float2bfloat16(A[i]);
  const bool b_mn = p.b_mode >= B_MN_SW64;
  for (int j = 0; j < p.n; ++j)
    for (int k = 0; k < p.k; ++k) hB[b_mn ? k * p.n + j : j * p.k + k] = __float2bfloat16(B[j * p.k + k]);
  __nv_bfloat16 *dA, *dB;
  float *dI, *dD;
  cudaMalloc(&dA, hA.size() * 2);
  cudaMalloc(&dB, hB.size() * 2);
  cudaMalloc(&dI, I.size() * 4);
  cudaMalloc(&dD, ref.size() * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dI, I.data(), I.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, ref.size() * 4);
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 132 * 1024);
  probe_kernel<<<1, 128, 132 * 1024>>>(dA, dB, dI, dD, p);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("%-58s CUDA ERROR %s\n", name, cudaGetErrorString(e));
    exit(2);
  }
  std::vector<float> out(ref.size());
  cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  for (size_t i = 0; i < out.size(); ++i) {
    maxerr = fmax(maxerr, fabs(double(out[i]) - ref[i]));
    maxref = fmax(maxref, fabs(ref[i]));
  }
  const bool ok = maxerr <= 1e-3 * maxref + 1e-4;
  printf("%-58s %s  max_err=%.4g (ref absmax %.4g)\n", name, ok ? "PASS" : "FAIL", maxerr, maxref);
  cudaFree(dA); cudaFree(dB); cudaFree(dI); cudaFree(dD);
  return ok;
}

int main() {
  run("T1 A sw128 K-major, B sw128 K-major n=96 k=64", {A_SMEM_SW128, B_K_SW128, 96, 64, 0, 0, 0});
  run("T1b A sw128 K-major, B sw128 K-major n=96 k=256", {A_SMEM_SW128, B_K_SW128, 96, 256, 0, 0, 0});
  run("T2 A sw64 K-major, B sw64 K-major n=128 k=32", {A_SMEM_SW64, B_K_SW64, 128, 32, 0, 0, 0});
  run("T2b A sw64 K-major, B sw64 K-major n=112 k=32", {A_SMEM_SW64, B_K_SW64, 112, 32, 0, 0, 0});
  const uint32_t lbos[] = {16, 512, 1024, 64}, sbos[] = {512, 1024, 16, 64};
  for (uint32_t l : lbos)
    for (uint32_t s : sbos) {
      char nm[96];
      snprintf(nm, sizeof nm, "T3 A tmem, B MN-major sw64 k=112 n=32 lbo=%u sbo=%u", l, s);
      run(nm, {A_TMEM, B_MN_SW64, 32, 112, 0, l, s});
    }
  for (uint32_t l : lbos)
    for (uint32_t s : sbos) {
      char nm[96];
      snprintf(nm, sizeof nm, "T3s A sw128, B MN-major sw64 k=64 n=32 lbo=%u sbo=%u", l, s);
      run(nm, {A_SMEM_SW128, B_MN_SW64, 32, 64, 0, l, s});
    }
  const uint32_t lbo2[] = {16, 1024, 2048, 8192}, sbo2[] = {1024, 2048, 16};
  for (uint32_t l : lbo2)
    for (uint32_t s : sbo2) {
      char nm[96];
      snprintf(nm, sizeof nm, "T3w A tmem, B MN-major sw128 k=128 n=64 lbo=%u sbo=%u", l, s);
      run(nm, {A_TMEM, B_MN_SW128, 64, 128, 0, l, s});
    }
  const uint32_t lbo3[] = {8192, 16, 1024, 4096};
  for (uint32_t l : lbo3) {
    char nm[96];
    snprintf(nm, sizeof nm, "T7 A MN-major sw128 m=128 k=64 lbo=%u, B K sw128 n=64", l);
    run(nm, {A_MN_SW128, B_K_SW128, 64, 64, 0, 0, 0, l, 1024});
    snprintf(nm, sizeof nm, "T8 A K sw128, B MN-major sw128 n=128 k=64 lbo=%u", l);
    run(nm, {A_SMEM_SW128, B_MN_SW128, 128, 64, 0, l, 1024, 0, 0});
    snprintf(nm, sizeof nm, "T9 A MN sw128, B MN sw128 n=128 k=64 lbo=%u", l);
    run(nm, {A_MN_SW128, B_MN_SW128, 128, 64, 0, l, 1024, l, 1024});
    snprintf(nm, sizeof nm, "T9b A MN sw128, B MN sw128 n=256 k=64 lbo=%u", l);
    run(nm, {A_MN_SW128, B_MN_SW128, 256, 64, 0, l, 1024, l, 1024});
  }
  run("T4 prefilled D += A sw128 * B sw128 n=256 k=64", {A_SMEM_SW128, B_K_SW128, 256, 64, 1, 0, 0});
  run("T5 A tmem, B K-major sw128 n=32 k=128 (V^T fallback)", {A_TMEM, B_K_SW128, 32, 128, 0, 0, 0});
  run("T6 A tmem, B K-major sw128 n=128 k=64", {A_TMEM, B_K_SW128, 128, 64, 0, 0, 0});
  return 0;
}
