// LocalAttention core (enhanced_generator.py:22-35), bf16 tensor-core forward.
//
// Per 4x4 window: S = Qh Kh^T is a [C x C] GEMM with K = 16 pixels, out = softmax_rows(S) V is
// [C x 16] with K = C.  Both contractions are far too small / too short-K for a tcgen05 + TMEM
// round trip per window (one UMMA, then TMEM->reg->TMEM for the softmax, then C/16 tiny-N UMMAs),
// so this kernel keeps each 16-row slab of S in registers, flash-attention style:
//   warp-level mma.sync m16n8k16 (bf16 in, fp32 accumulate) for S; exp on the packed accumulator fragments (half on the
//   XU pipe, half as an fp16 polynomial on the FMA pipe), which ARE the B fragments of the transposed second product
//   O^T = V^T P^T (f16 mma); an all-ones A operand yields the softmax row sums in the output's column layout.
// The window's q, k, v (16 pixels x 3C, pixel-major exactly as the qkv conv wrote them) are staged by
// cp.async into padded smem rows (pitch 2C+16 B => conflict-free ldmatrix), double-buffered across
// windows; k is rescaled in place by log2(e) / (|q_p| |k_p|) (eps 1e-12 on each norm), v converted to fp16.
// |S| <= 1, so exp needs no running max.  The C x C logits never leave the SM.
// (A one-warp-per-window variant was measured slower at every C and removed.)
#include <stdlib.h>

#include "common.cuh"
#include "la_mma.cuh"

namespace msg {
namespace {
using namespace la;

constexpr int LT_THREADS = 128;
constexpr int LT_WARPS = 4;
constexpr int P16 = 16;
#ifndef LA_CTAS_128
#define LA_CTAS_128 4   // 5 (96 registers, no spills) measured the same 0.40 ms: not occupancy-bound
#endif

template <int C, bool POLY>
__global__ void __launch_bounds__(LT_THREADS, (C == 64 ? 8 : C == 128 ? LA_CTAS_128 : 3))
local_attn_fwd_tc_kernel(const __nv_bfloat16* __restrict__ qkv, int N, int H, int W,
                         __nv_bfloat16* __restrict__ out) {
  constexpr int PITCH = 2 * C + 16;           // bytes per pixel row of one of q / k / v
  constexpr int MAT = P16 * PITCH;            // one matrix
  constexpr int BUF = 3 * MAT;                // q, k, v of one window
  constexpr int CH = C / 8;                   // 16-byte chunks per pixel per matrix
  constexpr int NCHUNK = (P16 * 3 * CH + LT_THREADS - 1) / LT_THREADS;   // cp.async copies per thread per window
  extern __shared__ __align__(16) uint8_t sm[];
  uint8_t* bufs = sm;                         // [2][BUF]
  uint8_t* os = sm + 2 * BUF;                 // [16][PITCH] output tile
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wpr = W / 4, wpi = (H / 4) * wpr;
  const int nwin = N * wpi;                   // < 2^31 (checked on the host): 32-bit window arithmetic, no 64-bit div/mod per window

  // the (pixel, chunk) -> (smem, global) mapping of this thread's copies is window-independent: compute once
  uint32_t c_dst[NCHUNK];
  int c_src[NCHUNK];
#pragma unroll
  for (int i = 0; i < NCHUNK; ++i) {
    const int c = tid + i * LT_THREADS;
    if (c < P16 * 3 * CH) {
      const int p = c / (3 * CH), cc = c - p * (3 * CH);
      const int part = cc / CH, off = cc - part * CH;
      c_dst[i] = (uint32_t)(part * MAT + p * PITCH + off * 16);
      c_src[i] = ((p >> 2) * W + (p & 3)) * (3 * C) + cc * 8;
    } else {
      c_dst[i] = 0xffffffffu;
      c_src[i] = 0;
    }
  }
  auto issue_load = [&](int wi, int b) {
    const int n = wi / wpi;
    const int r = wi - n * wpi;
    const int rr = r / wpr;
    const int h0 = rr * 4, w0 = (r - rr * wpr) * 4;
    const uint32_t dst0 = s_u32(bufs + b * BUF);
    const __nv_bfloat16* src0 = qkv + (((size_t)n * H + h0) * W + w0) * (3 * C);
#pragma unroll
    for (int i = 0; i < NCHUNK; ++i)
      if (c_dst[i] != 0xffffffffu) cpa16(dst0 + c_dst[i], src0 + c_src[i]);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  int wi = blockIdx.x;
  if (wi < nwin) issue_load(wi, 0);
  int b = 0;
  for (; wi < nwin; wi += gridDim.x, b ^= 1) {
    const int nxt = wi + (int)gridDim.x;
    if (nxt < nwin) {
      issue_load(nxt, b ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    uint8_t* qs = bufs + b * BUF;
    uint8_t* ks = qs + MAT;
    uint8_t* vs = ks + MAT;
    // ---- L2 normalisation of q and k over C (per pixel).  S = sum_p qh[i][p] kh[j][p] only needs the
    //      PRODUCT 1/(|q_p| |k_p|) per pixel p (the contraction index), so q stays untouched and k is
    //      scaled by both norms: half the smem writes.  8 threads per pixel.
    {
      const int p = tid >> 3, part = tid & 7;
      constexpr int VPT = (C + 63) / 64;      // 16-byte vectors per thread; vector v covers chunk part + 8 v, so a
                                              // quarter-warp touches 128 contiguous bytes (conflict-free LDS/STS.128)
      const bool act = part * 8 < C;          // C = 32: only 4 chunks per pixel
      const uint4* qp = reinterpret_cast<const uint4*>(qs + p * PITCH) + part;
      uint4* kp = reinterpret_cast<uint4*>(ks + p * PITCH) + part;
      uint4* vp = reinterpret_cast<uint4*>(vs + p * PITCH) + part;
      uint4 kv[VPT];
      float sq = 0.f, sk = 0.f;
#pragma unroll
      for (int v = 0; v < VPT; ++v) {
        if (act) {                              // v: bf16 -> fp16 in place (exact unless |v| > 65504 or subnormal)
          const uint4 vv = vp[8 * v];
          vp[8 * v] = make_uint4(pack_f16x2(bf_lo(vv.x), bf_hi(vv.x)), pack_f16x2(bf_lo(vv.y), bf_hi(vv.y)),
                                 pack_f16x2(bf_lo(vv.z), bf_hi(vv.z)), pack_f16x2(bf_lo(vv.w), bf_hi(vv.w)));
        }
        const uint4 qv = act ? qp[8 * v] : make_uint4(0, 0, 0, 0);
        kv[v] = act ? kp[8 * v] : make_uint4(0, 0, 0, 0);
        const uint32_t qw[4] = {qv.x, qv.y, qv.z, qv.w}, kw[4] = {kv[v].x, kv[v].y, kv[v].z, kv[v].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          sq = fmaf(bf_lo(qw[e]), bf_lo(qw[e]), fmaf(bf_hi(qw[e]), bf_hi(qw[e]), sq));
          sk = fmaf(bf_lo(kw[e]), bf_lo(kw[e]), fmaf(bf_hi(kw[e]), bf_hi(kw[e]), sk));
        }
      }
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
        sk += __shfl_xor_sync(0xffffffffu, sk, o);
      }
      // 1/(|q||k|) with the reference's eps clamp on each norm; log2(e) folded in so exp becomes ex2
      const float sc = 1.4426950408889634f * rsqrtf(fmaxf(sq, 1e-24f)) * rsqrtf(fmaxf(sk, 1e-24f));
#pragma unroll
      for (int v = 0; v < VPT; ++v) {
        uint32_t kw[4] = {kv[v].x, kv[v].y, kv[v].z, kv[v].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) kw[e] = pack_bf16x2(bf_lo(kw[e]) * sc, bf_hi(kw[e]) * sc);
        if (act) kp[8 * v] = make_uint4(kw[0], kw[1], kw[2], kw[3]);
      }
    }
    __syncthreads();
    // ---- attention rows: each warp takes 16-row slabs of S
    const uint32_t qs_a = s_u32(qs), ks_a = s_u32(ks), vs_a = s_u32(vs);
    const int mi = lane >> 3, r8 = lane & 7;
    const int g = lane >> 2, q4 = lane & 3;
    for (int rt = warp; rt < C / 16; rt += LT_WARPS) {
      const int i0 = rt * 16;
      uint32_t afr[4];
      // A = Qh^T slab: stored [pixel][channel]; matrices: (p 0-7, i0..), (p 0-7, i0+8..), (p 8-15, i0..), (p 8-15, i0+8..)
      ldsm_x4_t(qs_a + (r8 + 8 * (mi >> 1)) * PITCH + (i0 + 8 * (mi & 1)) * 2, afr);
      float acc[C / 8][4];
#pragma unroll
      for (int nt = 0; nt < C / 8; ++nt) { acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f; }
#pragma unroll
      for (int nt = 0; nt < C / 8; nt += 2) {
        uint32_t bfr[4];
        // B = Kh: matrices (p 0-7, j 8nt..), (p 8-15, j 8nt..), (p 0-7, j 8(nt+1)..), (p 8-15, j 8(nt+1)..)
        ldsm_x4_t(ks_a + (r8 + 8 * (mi & 1)) * PITCH + (8 * (nt + (mi >> 1))) * 2, bfr);
        mma_bf16(acc[nt], afr, bfr[0], bfr[1]);
        mma_bf16(acc[nt + 1], afr, bfr[2], bfr[3]);
      }
      // P = exp(S) straight into packed bf16 pairs (K carries log2(e), so this is one packed ex2 per
      // pair).  The S accumulator layout (row g | g+8, cols 2q4, 2q4+1) is the B-fragment layout of P^T,
      // so the second GEMM is computed transposed: O^T[pixel][i] = sum_j V^T[pixel][j] P^T[j][i], which
      // puts channel pairs in one thread (4-byte conflict-free stores into the [pixel][channel] tile),
      // and an all-ones A operand yields the softmax row sums in the same column layout for free.
      // P is formed in fp16 (the PV product runs as an f16 mma; v was converted in the staging pass): every other
      // column tile takes its exponentials on the FMA pipe (packed polynomial) instead of the XU pipe, which ncu
      // showed to be the busiest unit of this kernel (58 %).
      uint32_t pk[C / 8][2];
#pragma unroll
      for (int nt = 0; nt < C / 8; ++nt) {
        const uint32_t x0 = pack_f16x2(acc[nt][0], acc[nt][1]), x1 = pack_f16x2(acc[nt][2], acc[nt][3]);
        if ((nt & 1) && POLY) {
          pk[nt][0] = exp2_poly_f16x2(x0);
          pk[nt][1] = exp2_poly_f16x2(x1);
        } else {
          pk[nt][0] = ex2_f16x2(x0);
          pk[nt][1] = ex2_f16x2(x1);
        }
      }
      float o[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      float rsum[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      const uint32_t ones[4] = {0x3C003C00u, 0x3C003C00u, 0x3C003C00u, 0x3C003C00u};     // fp16 1.0
#pragma unroll
      for (int ks16 = 0; ks16 < C / 16; ++ks16) {
        uint32_t va[4];
        // A = V^T stored [pixel][channel] = [m][k]: (p 0-7, j 16ks..), (p 8-15, j 16ks..), (p 0-7, +8), (p 8-15, +8)
        ldsm_x4(vs_a + (r8 + 8 * (mi & 1)) * PITCH + (16 * ks16 + 8 * (mi >> 1)) * 2, va);
        mma_f16(o[0], va, pk[2 * ks16][0], pk[2 * ks16 + 1][0]);
        mma_f16(o[1], va, pk[2 * ks16][1], pk[2 * ks16 + 1][1]);
        mma_f16(rsum[0], ones, pk[2 * ks16][0], pk[2 * ks16 + 1][0]);
        mma_f16(rsum[1], ones, pk[2 * ks16][1], pk[2 * ks16 + 1][1]);
      }
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const float inv0 = __fdividef(1.f, rsum[nt][0]), inv1 = __fdividef(1.f, rsum[nt][1]);   // sums are in [C/e, C e]
        const int col = (i0 + 8 * nt + 2 * q4) * 2;
        *reinterpret_cast<uint32_t*>(os + g * PITCH + col) = pack_bf16x2(o[nt][0] * inv0, o[nt][1] * inv1);
        *reinterpret_cast<uint32_t*>(os + (g + 8) * PITCH + col) = pack_bf16x2(o[nt][2] * inv0, o[nt][3] * inv1);
      }
    }
    __syncthreads();
    {
      const int n = wi / wpi;
      const int r = wi - n * wpi;
      const int rr = r / wpr;
      const int h0 = rr * 4, w0 = (r - rr * wpr) * 4;
      for (int c = tid; c < P16 * CH; c += LT_THREADS) {
        const int p = c / CH, off = c - p * CH;
        uint4 val = *reinterpret_cast<const uint4*>(os + p * PITCH + off * 16);
        *reinterpret_cast<uint4*>(out + (((size_t)n * H + h0 + (p >> 2)) * W + w0 + (p & 3)) * C + off * 8) = val;
      }
    }
    // the next iteration's first __syncthreads orders these reads of `os` / this buffer before reuse
  }
}

template <int C, bool POLY>
int launch_impl(const __nv_bfloat16* qkv, int N, int H, int W, __nv_bfloat16* out, cudaStream_t st) {
  constexpr int PITCH = 2 * C + 16;
  const size_t smem = (size_t)(2 * 3 + 1) * P16 * PITCH;
  cudaError_t e = cudaFuncSetAttribute(local_attn_fwd_tc_kernel<C, POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  MSG_REQUIRE(e == cudaSuccess, MSG_ERR_CUDA, "local_attn_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  const long long nwin = (long long)N * (H / 4) * (W / 4);
  MSG_REQUIRE(nwin + 64LL * 1024 < 0x7fffffffLL, MSG_ERR_SHAPE, "local_attn_tc: too many windows");
  const int per_sm = C >= 256 ? 3 : (C >= 128 ? LA_CTAS_128 : 8);   // resident CTAs per SM (smem / register limits)
  long long grid = (long long)per_sm * sm_count();
  if (grid > nwin) grid = nwin;
  local_attn_fwd_tc_kernel<C, POLY><<<(unsigned)grid, LT_THREADS, smem, st>>>(qkv, N, H, W, out);
  return check_launch("local_attn_fwd_tc_kernel");
}
template <int C>
int launch(const __nv_bfloat16* qkv, int N, int H, int W, __nv_bfloat16* out, cudaStream_t st) {
  // MSG_LA_POLY=0: every exponential on the XU pipe (for A/B measurements of the polynomial split)
  static const bool poly = [] { const char* e = getenv("MSG_LA_POLY"); return !(e && e[0] == '0'); }();
  return poly ? launch_impl<C, true>(qkv, N, H, W, out, st) : launch_impl<C, false>(qkv, N, H, W, out, st);
}

}  // namespace

bool local_attn_tc_supported(int dtype, int C, const void* qkv, const void* out) {
  if (dtype != MSG_BF16) return false;
  if (C != 32 && C != 64 && C != 128 && C != 256) return false;
  return (((uintptr_t)qkv | (uintptr_t)out) & 15) == 0;
}

int local_attn_fwd_tc(const void* qkv, int N, int H, int W, int C, void* out, cudaStream_t st) {
  auto q = (const __nv_bfloat16*)qkv;
  auto o = (__nv_bfloat16*)out;
  switch (C) {
    case 32: return launch<32>(q, N, H, W, o, st);
    case 64: return launch<64>(q, N, H, W, o, st);
    case 128: return launch<128>(q, N, H, W, o, st);
    case 256: return launch<256>(q, N, H, W, o, st);
  }
  set_error("local_attn_tc: unsupported C=%d", C);
  return MSG_ERR_UNSUPPORTED;
}

}  // namespace msg
