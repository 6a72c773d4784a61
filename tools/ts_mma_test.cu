// Micro-test of the tcgen05 building blocks the fused LocalAttention stage kernel relies on (not product code):
//   1. S = Q^T K with BOTH operands MN-major (pixel rows = K, 64-channel 128-byte rows, SWIZZLE_128B), M = N = 128, K = 16
//   2. P = exp2(S) written back to TMEM as packed fp16 with tcgen05.st (in place of S), tcgen05.wait::st
//   3. O = P V with the A operand read FROM TMEM (tcgen05.mma [d], [a], b_desc ...), B = V K-major fp16, N = 32 / 48
// and of the cost per TS-form MMA as a function of N.  Prints max errors against a host reference.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/build/ts_mma_test tools/ts_mma_test.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "../multi_style_transfer_gan_b200/csrc/tcgen05.cuh"

using namespace msg::tc;

__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

constexpr int C = 128, NV = 48;          // channels, rows of the V operand (16 pixels + 8 ones + pad)
constexpr int QK_LBO = 2048;             // next 64-channel block of a window's q / k tile (16 rows x 128 B each)
constexpr int V_KB = NV * 128;           // bytes per 64-channel K block of V

// smem image (bytes): Q [2][16][128] | K [2][16][128] | V [2][NV][128]
__global__ void __launch_bounds__(128, 1) ts_test_kernel(const uint8_t* img, int img_bytes, float* S_out, float* O_out, int NB,
                                                         int timing_iters, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  for (int i = threadIdx.x; i < img_bytes / 16; i += blockDim.x) reinterpret_cast<uint4*>(gen)[i] = reinterpret_cast<const uint4*>(img)[i];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    if (lane == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t sQ = base, sK = base + 4096, sV = base + 8192;
  const uint32_t bar_a = smem_u32(&bar);
  const uint32_t S_COL = 0, O_COL = 256;
  // ---- 1. S = Q^T K  (bf16, both MN-major)
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    umma_bf16(tmem + S_COL, make_mn_sw128_desc(sQ, QK_LBO), make_mn_sw128_desc(sK, QK_LBO), idesc, 0);
    umma_commit(bar_a);
  }
  mbar_wait(bar_a, 0);
  tc_fence_after();
  // ---- 2. P = exp2(S) -> fp16 pairs -> TMEM (cols S_COL .. S_COL+63 of the same lanes)
  const uint32_t lane_addr = ((uint32_t)(warp * 32) << 16);
  const int row = warp * 32 + lane;
  uint32_t pk[64];
  for (int h = 0; h < 4; ++h) {
    float v[32];
    tmem_ld32_sync(tmem + lane_addr + S_COL + h * 32, v);
    for (int j = 0; j < 32; ++j) S_out[row * 128 + h * 32 + j] = v[j];
    for (int j = 0; j < 16; ++j) {
      __half2 t = __floats2half2_rn(exp2f(v[2 * j]), exp2f(v[2 * j + 1]));
      pk[h * 16 + j] = *reinterpret_cast<uint32_t*>(&t);
    }
  }
  tmem_st32(tmem + lane_addr + S_COL, *reinterpret_cast<uint32_t(*)[32]>(&pk[0]));
  tmem_st32(tmem + lane_addr + S_COL + 32, *reinterpret_cast<uint32_t(*)[32]>(&pk[32]));
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // ---- 3. O = P V  (A = P from TMEM, fp16; B = V K-major fp16, N = NB)
  const uint32_t idesc2 = (1u << 4) | ((uint32_t)(NB >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  if (threadIdx.x == 0) {
    for (int ks = 0; ks < C / 16; ++ks) {
      const uint64_t db = make_sw128_desc(sV + (ks >> 2) * V_KB) + (uint64_t)((ks & 3) * 2);
      umma_ts(tmem + O_COL, tmem + S_COL + ks * 8, db, idesc2, ks != 0);
    }
    umma_commit(bar_a);
  }
  mbar_wait(bar_a, 1);
  tc_fence_after();
  {
    float v[32];
    tmem_ld32_sync(tmem + lane_addr + O_COL, v);
    for (int j = 0; j < 32; ++j) O_out[row * 64 + j] = v[j];
    tmem_ld32_sync(tmem + lane_addr + O_COL + 32, v);
    for (int j = 0; j < 32; ++j) O_out[row * 64 + 32 + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // ---- 4. timing: TS-form MMAs back to back
  if (threadIdx.x == 0 && timing_iters > 0) {
    const long long t0 = clock64();
    for (int it = 0; it < timing_iters; ++it) {
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const uint64_t db = make_sw128_desc(sV + (ks >> 2) * V_KB) + (uint64_t)((ks & 3) * 2);
        umma_ts(tmem + O_COL + (uint32_t)((ks & 1) * 64), tmem + S_COL + ks * 8, db, idesc2, 1);
      }
    }
    umma_commit(bar_a);
    mbar_wait(bar_a, 0);
    cycles[0] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

static float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
static float f16_round(float x) { return __half2float(__float2half_rn(x)); }

int main() {
  // logical data: q[p][c], k[p][c] (16 pixels, 128 channels, bf16), v[n][c] (NV rows, fp16): rows 16..23 of v are ones
  std::vector<float> q(16 * C), k(16 * C), v(NV * C);
  srand(1);
  auto rnd = [] { return (float)rand() / RAND_MAX * 2.f - 1.f; };
  for (auto& x : q) x = bf16_round(rnd() * 0.3f);
  for (auto& x : k) x = bf16_round(rnd() * 0.3f);
  for (int n = 0; n < NV; ++n)
    for (int c = 0; c < C; ++c) v[n * C + c] = n < 16 ? f16_round(rnd()) : (n < 24 ? 1.f : f16_round(rnd()));
  const int img_bytes = 4096 + 4096 + 2 * NV * 128;
  std::vector<uint8_t> img(img_bytes, 0);
  auto put_qk = [&](int off, const std::vector<float>& m) {
    for (int p = 0; p < 16; ++p)
      for (int c = 0; c < C; ++c) {
        const int b = c / 64, cc = c % 64;
        const int a = off + b * QK_LBO + (p >> 3) * 1024 + (p & 7) * 128 + (((cc >> 3) ^ (p & 7)) << 4) + (cc & 7) * 2;
        __nv_bfloat16 h = __float2bfloat16_rn(m[p * C + c]);
        memcpy(&img[a], &h, 2);
      }
  };
  put_qk(0, q);
  put_qk(4096, k);
  for (int n = 0; n < NV; ++n)
    for (int c = 0; c < C; ++c) {
      const int kb = c / 64, cc = c % 64;
      const int a = 8192 + kb * V_KB + (n >> 3) * 1024 + (n & 7) * 128 + (((cc >> 3) ^ (n & 7)) << 4) + (cc & 7) * 2;
      __half h = __float2half_rn(v[n * C + c]);
      memcpy(&img[a], &h, 2);
    }
  uint8_t* d_img; float *d_S, *d_O; long long* d_cyc;
  cudaMalloc(&d_img, img_bytes); cudaMalloc(&d_S, 128 * 128 * 4); cudaMalloc(&d_O, 128 * 64 * 4); cudaMalloc(&d_cyc, 8);
  cudaMemcpy(d_img, img.data(), img_bytes, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(ts_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  std::vector<float> S(128 * 128), O(128 * 64);
  int rc = 0;
  for (int NB : {32, 48, 16}) {
    ts_test_kernel<<<1, 128, 40 * 1024>>>(d_img, img_bytes, d_S, d_O, NB, 0, d_cyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(S.data(), d_S, S.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(O.data(), d_O, O.size() * 4, cudaMemcpyDeviceToHost);
    double eS = 0, eO = 0, mO = 0;
    for (int i = 0; i < C; ++i) {
      std::vector<float> P(C);
      for (int j = 0; j < C; ++j) {
        double s = 0;
        for (int p = 0; p < 16; ++p) s += (double)q[p * C + i] * k[p * C + j];
        eS = fmax(eS, fabs(s - S[i * 128 + j]));
        P[j] = f16_round(exp2f(S[i * 128 + j]));
      }
      for (int n = 0; n < NB; ++n) {
        double o = 0;
        for (int j = 0; j < C; ++j) o += (double)P[j] * v[n * C + j];
        eO = fmax(eO, fabs(o - O[i * 64 + n]));
        mO = fmax(mO, fabs(o));
      }
    }
    printf("N=%d: max |S - ref| = %.3e, max |O - ref| = %.3e (max |O| %.2f), row-sum col16 of row 5: %.4f\n", NB, eS, eO, mO, O[5 * 64 + 16]);
    if (eS > 1e-3 || eO > 2e-2) { printf("MISMATCH\n"); rc = 2; }
  }
  for (int NB : {16, 32, 48, 64, 128}) {
    const int iters = 2000;
    ts_test_kernel<<<148, 128, 40 * 1024>>>(d_img, img_bytes, d_S, d_O, NB > 48 ? 48 : NB, 0, d_cyc);   // warm
    cudaDeviceSynchronize();
    if (NB > 48) continue;    // V image only has 48 rows
    ts_test_kernel<<<148, 128, 40 * 1024>>>(d_img, img_bytes, d_S, d_O, NB, iters, d_cyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    long long c = 0;
    cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
    printf("TS-form MMA (A from TMEM, M=128, K=16) N=%d: %.1f cycles per MMA\n", NB, (double)c / (iters * 8.0));
  }
  return rc;
}
