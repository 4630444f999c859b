// tools/tmem_bw_probe.cu — tcgen05.ld / tcgen05.st throughput by shape and warp count (one CTA per SM, 1 SM measured).
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/tmem_bw_probe.bin tools/tmem_bw_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include "../multi-modal-multi-label-facial-action-unit-detection-with-transformer_b200/csrc/avf_common.cuh"
namespace avf { void set_error(const char*, ...) {} int check_cuda(cudaError_t e, const char*) { return int(e); } void count_launch() {} }
using namespace avf;

__device__ __forceinline__ void ld_32x32b_x64(uint32_t taddr, uint32_t (&r)[64]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
      "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
        "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
        "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]),
        "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]),
        "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]),
        "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ld_16x256b_x8(uint32_t taddr, uint32_t (&r)[32]) {   // 16 lanes x 256 bit, x8: 32 regs per thread
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
        "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
        "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}

// mode 0: ld 32x32b.x32 ; 1: ld 32x32b.x64 ; 2: ld 16x256b.x8 ; 3: st 32x32b.x32 ; 4: ld x32 without per-iteration wait (wait every 4)
// 5: ld 32x32b.x8
__global__ void __launch_bounds__(512) bw_kernel(int mode, int iters, long long* cycles, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot + (uint32_t((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    const uint32_t col = uint32_t(((i + warp) * 32) & 255);
    if (mode == 0) { uint32_t r[32]; tmem_ld32(tm + col, r); tmem_ld_wait(); acc += r[0] ^ r[31]; }
    else if (mode == 1) { uint32_t r[64]; ld_32x32b_x64(tm + col, r); tmem_ld_wait(); acc += r[0] ^ r[63]; }
    else if (mode == 2) { uint32_t r[32]; ld_16x256b_x8(tm + col, r); tmem_ld_wait(); acc += r[0] ^ r[31]; }
    else if (mode == 3) { uint32_t r[32]; for (int j = 0; j < 32; ++j) r[j] = acc + j; tmem_st32(tm + col, r); tmem_st_wait(); }
    else if (mode == 4) { uint32_t r[32]; tmem_ld32(tm + col, r); if ((i & 3) == 3) tmem_ld_wait(); acc += r[0]; }
    else { uint32_t r[8]; asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(tm + col) : "memory"); tmem_ld_wait(); acc += r[0] ^ r[7]; }
  }
  tmem_ld_wait();
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) *cycles = t1 - t0;
  sink[threadIdx.x] = acc;
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(slot, 512);
}

int main() {
  long long* dc; uint32_t* ds;
  cudaMalloc(&dc, 8); cudaMalloc(&ds, 4 * 512);
  const char* names[] = {"ld 32x32b.x32", "ld 32x32b.x64", "ld 16x256b.x8 (32 regs)", "st 32x32b.x32", "ld 32x32b.x32 (wait every 4)", "ld 32x32b.x8"};
  const int bytes_per_warp_inst[] = {32 * 32 * 4, 32 * 64 * 4, 32 * 32 * 4, 32 * 32 * 4, 32 * 32 * 4, 32 * 8 * 4};
  for (int mode = 0; mode < 6; ++mode)
    for (int warps : {1, 4, 8, 16}) {
      const int iters = 2000;
      bw_kernel<<<1, warps * 32>>>(mode, iters, dc, ds);
      bw_kernel<<<1, warps * 32>>>(mode, iters, dc, ds);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s warps=%d: CUDA error %s\n", names[mode], warps, cudaGetErrorString(e)); return 1; }
      long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
      printf("%-30s warps=%2d: %8lld cycles, %7.1f cyc/inst/warp, %7.1f B/clk/SM\n", names[mode], warps, c, double(c) / iters,
             double(bytes_per_warp_inst[mode]) * iters * warps / double(c));
    }
  return 0;
}
