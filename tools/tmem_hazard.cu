// Micro-benchmark: does a tcgen05.ld wait for in-flight tcgen05.mma that accumulate into OTHER tensor-memory columns?
// One thread issues a long stream of accumulating MMAs (M=128, K=16, N=16) into columns [0, 16); the four warps of a second
// warpgroup meanwhile time tcgen05.ld (16 columns) + wait::ld at column offset `c`.  If loads are ordered behind MMAs by a
// hazard check coarser than the columns they touch, the load latency jumps for small c (motivated by csrc/msb_ring.cu, whose
// 16-column row accumulators sit side by side).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/build/tmem_hazard tools/tmem_hazard.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../multi_style_transfer_gan_b200/csrc/tcgen05.cuh"
using namespace msg::tc;

__global__ void __launch_bounds__(256, 1) k(int mma_on, int mma_n, int ld_col, int iters, long long* out, float* sink) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  __shared__ volatile int stop;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { stop = 0; mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 0) {
    if (lane == 0 && mma_on) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(mma_n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint64_t da = make_sw128_desc(base), db = make_sw128_desc(base + 16384);
      while (!stop) {
#pragma unroll
        for (int u = 0; u < 16; ++u) umma_bf16_acc(tmem, da + (u & 3) * 2, db + (u & 3) * 2, idesc);
      }
      umma_commit(smem_u32(&bar));
      mbar_wait(smem_u32(&bar), 0);
    }
  } else if (warp >= 4) {
    const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)ld_col;
    float acc = 0.f;
    __nanosleep(2000);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      float v[16];
      tmem_ld16(taddr, v);
      tmem_ld_wait();
      acc += v[0] + v[15];
    }
    const long long t1 = clock64();
    if (acc == 1234.5f) sink[0] = acc;
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (threadIdx.x == 128) { out[blockIdx.x] = t1 - t0; stop = 1; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

int main() {
  long long* d; float* sink;
  cudaMalloc(&d, 148 * 8); cudaMalloc(&sink, 4);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int iters = 2000;
  for (int n : {16, 48}) {
    for (int on : {0, 1}) {
      for (int c : {16, 32, 48, 64, 96, 128, 256, 384, 496}) {
        if (c < n) continue;
        k<<<148, 256, 56 * 1024>>>(on, n, c, iters, d, sink);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
        long long cyc; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
        printf("MMA N=%2d stream %s, tcgen05.ld x16 at column %3d: %.1f cycles per ld + wait\n", n, on ? "ON " : "off", c, (double)cyc / iters);
      }
    }
  }
  return 0;
}
