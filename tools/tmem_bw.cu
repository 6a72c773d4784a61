// Micro-benchmark: tcgen05.ld / tcgen05.st throughput per SM as a function of the number of warps (not product code).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/build/tmem_bw tools/tmem_bw.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../multi_style_transfer_gan_b200/csrc/tcgen05.cuh"
using namespace msg::tc;

template <int MODE>   // 0: ld x32, 1: ld x16, 2: st x16 (packed), 3: ld x32 + 64 FMAs of dependent work
__global__ void k(int iters, long long* out, float* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot + ((uint32_t)((warp & 3) * 32) << 16);
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0 || MODE == 3) {
      float a[32], b[32];
      tmem_ld32(tmem + ((i * 64) & 255), a);
      tmem_ld32(tmem + ((i * 64 + 32) & 255), b);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) acc += a[j] * b[j];
    } else if (MODE == 1) {
      float a[16], b[16], c[16], d[16];
      tmem_ld16(tmem + ((i * 64) & 255), a);
      tmem_ld16(tmem + ((i * 64 + 16) & 255), b);
      tmem_ld16(tmem + ((i * 64 + 32) & 255), c);
      tmem_ld16(tmem + ((i * 64 + 48) & 255), d);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) acc += a[j] * b[j] + c[j] * d[j];
    } else {
      uint32_t r[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) r[j] = i + j;
      tmem_st16(tmem + ((i * 64) & 255), r);
      tmem_st16(tmem + ((i * 64 + 16) & 255), r);
      tmem_st16(tmem + ((i * 64 + 32) & 255), r);
      tmem_st16(tmem + ((i * 64 + 48) & 255), r);
      tmem_st_wait();
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u) : "memory");
}

int main() {
  long long* out; float* sink;
  cudaMalloc(&out, 1024 * 8); cudaMalloc(&sink, 4);
  const int iters = 4096;
  for (int mode = 0; mode < 3; ++mode)
    for (int warps : {1, 4, 8, 16, 32}) {
      if (mode == 0) k<0><<<148, warps * 32>>>(iters, out, sink);
      if (mode == 1) k<1><<<148, warps * 32>>>(iters, out, sink);
      if (mode == 2) k<2><<<148, warps * 32>>>(iters, out, sink);
      cudaError_t e = cudaDeviceSynchronize();
      long long c; cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
      const double bytes = (double)iters * warps * 32 * 64 * 4;
      printf("mode %d (%s) warps %2d: %lld cycles, %.1f B/clk/SM  (%s)\n", mode, mode == 0 ? "ld x32" : mode == 1 ? "ld x16" : "st x16", warps, c,
             bytes / c, cudaGetErrorString(e));
    }
  return 0;
}
