// Micro-benchmark: cost of one tcgen05.mma (M=128, K=16, bf16, cta_group::1) as a function of N when the issuing
// thread does NOTHING else (fully unrolled, operands resident in smem, no barriers inside the loop).  Separates the
// tensor pipe's own per-instruction cost from the cost of the surrounding issue loop.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/build/mma_rate tools/mma_rate.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../multi_style_transfer_gan_b200/csrc/tcgen05.cuh"

using namespace msg::tc;

template <int UNROLL>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int N, int iters, int same_k, int two_issuers, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar, bar2;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_init(smem_u32(&bar2), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0 || (two_issuers && threadIdx.x == 32)) {
    const uint32_t mybar = threadIdx.x == 0 ? smem_u32(&bar) : smem_u32(&bar2);
    const uint32_t col0 = threadIdx.x == 0 ? 0u : 128u;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t da = make_sw128_desc(base), db = make_sw128_desc(base + 16384);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const int k = (u & 3) * 2;
        const int shift = same_k * ((u >> 2) + 1) * 8;      // same_k = 1: A start address shifted by whole 128-byte rows (tap shift)
        umma_bf16_acc(tmem + col0 + (uint32_t)((u & 1) * 256), da + k + shift, db + k, idesc);
      }
    }
    umma_commit(mybar);
    mbar_wait(mybar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) cycles[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  cudaFuncSetAttribute(mma_rate_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int iters = 2000, U = 16;
  const int Ns[] = {8, 16, 32, 48, 64, 96, 128, 160, 192, 256};
  for (int mode : {0, 1, 2, 3}) {
    const int grid = 148;
    const int two = mode & 1, shifted = mode >> 1;
    for (int N : Ns) {
      if (mode && N > 128) continue;
      long long c = 0;
      mma_rate_kernel<U><<<grid, 128, 56 * 1024>>>(N, iters, shifted, two, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
      mma_rate_kernel<U><<<grid, 128, 56 * 1024>>>(N, iters, shifted, two, d);
      cudaDeviceSynchronize();
      cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
      const double per = (double)c / ((double)iters * U);
      printf("%s issuers %d  N=%3d  %.1f cycles per MMA of one issuer  (%.0f flop/clk/SM)\n", shifted ? "row-shifted A" : "aligned A", two + 1, N, per, (two + 1) * 2.0 * 128 * N * 16 / per);
    }
  }
  return 0;
}
