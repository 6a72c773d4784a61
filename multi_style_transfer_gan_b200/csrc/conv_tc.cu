// tcgen05 / TMEM implicit-GEMM gather convolution for sm_100a (bf16 operands, fp32 accumulate).
//
// One CTA computes a 128 (output pixels) x BN (output channels) tile:
//   * warps 0-3 gather the A operand (128 im2col rows x 64 K) straight from the NHWC activation with
//     16-byte cp.async (zero-fill for padding / K tail): 8 consecutive lanes fetch the 8 chunks of
//     ONE row's 128-byte K slab (one full line per row, 4 rows per warp instruction) into a
//     128B-swizzled, K-major shared-memory tile, and copy the matching [BN x 64] slab of the packed
//     weights; a 3-stage mbarrier ring (full / empty) decouples them from the tensor core;
//   * warp 4 allocates TMEM and its elected lane issues tcgen05.mma.cta_group::1.kind::f16
//     (UMMA 128 x BN x 16, A and B from shared-memory descriptors, D in TMEM), releasing each
//     stage with tcgen05.commit;
//   * after their last gather the same four warps become the epilogue: tcgen05.ld (32x32b) of their
//     TMEM lane quarter, + bias, optional InstanceNorm statistics (per-column sums reduced by
//     recursive halving over the warp, then across warps through smem, one fp64 atomic per channel
//     per tile), activation, bf16 pack, staging of the warp's 32 x BN tile in (now idle) pipeline
//     smem and fully coalesced 16-byte stores (or fp32 NCHW stores for the image).
// Several CTAs are resident per SM (<= 97 KB smem, <= 128 TMEM columns each), so one CTA's epilogue
// overlaps another's main loop.
//
// Same descriptor and semantics as the SIMT engine (conv_simt.cu); api.cu picks this path when
// conv2d_tc_supported() holds.
#include "common.cuh"
#include "tcgen05.cuh"

namespace msg {
namespace {
using namespace tc;

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;
constexpr int TC_STAGES = 3;
constexpr int TC_LAG = TC_STAGES - 1;
constexpr int TC_PRODUCERS = 128;
constexpr int TC_THREADS = 160;
constexpr int A_STAGE_BYTES = TC_BM * TC_BK * 2;  // 16 KB

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }



struct TcParams {
  msg_conv_desc d;
  const __nv_bfloat16* x;
  const __nv_bfloat16* w;
  const float* bias;
  void* y;
  double* stats;
  int BN;         // N tile (multiple of 16, <= 128)
  int tmem_cols;  // power of two >= 32, >= BN
  int K, nkb;
  long long M;
};

__global__ void __launch_bounds__(TC_THREADS)
conv_tc_kernel(const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const msg_conv_desc& d = p.d;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int BN = p.BN;
  const int b_stage_bytes = BN * TC_BK * 2;
  // carve: [A stages][B stages][barriers][tmem slot]   (base rounded up to 1024 B for SW128)
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base;
  const uint32_t sB = sA + TC_STAGES * A_STAGE_BYTES;
  const uint32_t sBar = sB + TC_STAGES * b_stage_bytes;   // full[S], empty[S], accum
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  static_assert(8 * (2 * TC_STAGES + 1) <= 56, "barrier block layout");
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen_base + (sBar - base) + 56);
  int4* rowinfo = reinterpret_cast<int4*>(gen_base + (sBar - base) + 64);   // [128], 16-byte aligned
  int* rowout = reinterpret_cast<int*>(rowinfo + TC_BM);                                               // [128]
  float* red = reinterpret_cast<float*>(gen_base);        // epilogue scratch aliases stage 0 of A
  auto full_bar = [&](int s) { return sBar + 8u * s; };
  auto empty_bar = [&](int s) { return sBar + 8u * (TC_STAGES + s); };
  const uint32_t accum_bar = sBar + 8u * (2 * TC_STAGES);

  if (warp == 4) {
    if (lane == 0) {
      for (int s = 0; s < TC_STAGES; ++s) { mbar_init(full_bar(s), TC_PRODUCERS); mbar_init(empty_bar(s), 1); }
      mbar_init(accum_bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const long long m0 = (long long)blockIdx.x * TC_BM;
  const int co0 = blockIdx.y * BN;
  const int hw = d.Hg * d.Wg;

  if (warp < 4) {
    // =============================== producer: im2col gather ===============================
    {   // per-row geometry, computed once: (ih0, iw0, image pixel base, valid) and output pixel index
      const long long m = m0 + tid;
      const bool valid = m < p.M;
      const long long mm = valid ? m : 0;
      const int n = (int)(mm / hw);
      const int rem = (int)(mm - (long long)n * hw);
      const int gi = rem / d.Wg, gj = rem - gi * d.Wg;
      rowinfo[tid] = make_int4(gi * d.in_stride - d.pad_h, gj * d.in_stride - d.pad_w, n * d.Hi * d.Wi, valid ? 1 : 0);
      rowout[tid] = (n * d.Ho + gi * d.out_stride + d.out_off_h) * d.Wo + gj * d.out_stride + d.out_off_w;
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    const int j = lane & 7;                 // this lane's 16-byte chunk of the 128-byte K slab
    const int rsub = lane >> 3;             // row within a group of 4
    const __nv_bfloat16* xin = p.x + d.ci_off;
    int th = 0, tw = 0, ci = 8 * j;         // (tap, channel) of chunk j in k-block 0
    while (ci >= d.Cin) { ci -= d.Cin; if (++tw == d.KW) { tw = 0; ++th; } }
    for (int kb = 0; kb < p.nkb; ++kb) {
      const int s = kb % TC_STAGES;
      if (kb >= TC_STAGES) mbar_wait(empty_bar(s), ((kb / TC_STAGES) - 1) & 1);
      const uint32_t a_stage = sA + s * A_STAGE_BYTES;
      const bool kvalid = th < d.KH;
      const int dh = th * d.dil, dw = tw * d.dil;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int row = warp * 32 + it * 4 + rsub;
        const int4 ri = rowinfo[row];
        const int ih = ri.x + dh, iw = ri.y + dw;
        const bool ok = ri.w && kvalid && ih >= 0 && ih < d.Hi && iw >= 0 && iw < d.Wi;
        const __nv_bfloat16* src = ok ? xin + ((size_t)ri.z + (size_t)ih * d.Wi + iw) * d.Ci_total + ci : p.x;
        cp_async16(a_stage + (uint32_t)row * 128u + (((uint32_t)j ^ ((uint32_t)row & 7u)) << 4), src, ok ? 16u : 0u);
      }
      ci += TC_BK;                          // same chunk, next k-block
      while (ci >= d.Cin) { ci -= d.Cin; if (++tw == d.KW) { tw = 0; ++th; } }
      const uint32_t b_base = sB + s * b_stage_bytes;
      const int k0 = kb * TC_BK;
      for (int c = tid; c < BN * 8; c += TC_PRODUCERS) {
        const int nr = c >> 3, jb = c & 7;
        const int kk = k0 + jb * 8, co = co0 + nr;
        const bool ok = co < d.Cout && kk < p.K;
        const __nv_bfloat16* src = ok ? p.w + (size_t)co * p.K + kk : p.w;
        cp_async16(b_base + (uint32_t)nr * 128u + (((uint32_t)jb ^ ((uint32_t)nr & 7u)) << 4), src, ok ? 16u : 0u);
      }
      cp_async_commit();
      if (kb >= TC_LAG) {
        cp_async_wait<TC_LAG>();
        fence_proxy_async();
        mbar_arrive(full_bar((kb - TC_LAG) % TC_STAGES));
      }
    }
    cp_async_wait<0>();
    fence_proxy_async();
    for (int kb = (p.nkb > TC_LAG ? p.nkb - TC_LAG : 0); kb < p.nkb; ++kb) mbar_arrive(full_bar(kb % TC_STAGES));

    // =============================== epilogue ===============================
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const bool do_stats = d.flags & MSG_CONV_STATS;
    const bool nchw = d.flags & MSG_CONV_OUT_NCHW_F32;
    const bool accum = d.flags & MSG_CONV_ACCUM;
    const int row = tid;
    const bool valid = rowinfo[row].w != 0;
    const int opix = rowout[row];
    const int n_img = (int)(m0 / hw);                              // stats: whole tile in one image
    const int cmax = (d.Cout - co0) < BN ? (d.Cout - co0) : BN;   // valid columns of this tile
    const bool staged = !nchw && ((d.Co_total | d.co_off | cmax) & 7) == 0;
    // staging tile of this warp: 32 rows x (BN*2 + 16) bytes (the +16 B skews consecutive rows by one
    // 16-byte bank group: conflict-free row-per-lane writes, contiguous chunk-per-lane reads)
    const int spitch = BN * 2 + 16;
    uint8_t* stage_w = gen_base + 1024 + warp * (32 * spitch);     // after `red` (first 1 KB)
    for (int c0 = 0; c0 < cmax; c0 += 32) {
      __syncwarp();
      float v[32];
      tmem_ld32_sync(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
      for (int jj = 0; jj < 32; ++jj)
        v[jj] += (p.bias != nullptr && c0 + jj < cmax) ? __ldg(p.bias + co0 + c0 + jj) : 0.f;
      if (do_stats) {   // dispatcher guarantees: every row valid and the whole tile in one image
        float s1[32], s2[32];
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) { s1[jj] = v[jj]; s2[jj] = v[jj] * v[jj]; }
        float cs = warp_transpose_reduce32(s1, lane);
        float css = warp_transpose_reduce32(s2, lane);
        red[(warp * 2 + 0) * 32 + lane] = cs;
        red[(warp * 2 + 1) * 32 + lane] = css;
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (tid < 32 && c0 + tid < cmax) {
          float a = red[0 * 32 + tid] + red[2 * 32 + tid] + red[4 * 32 + tid] + red[6 * 32 + tid];
          float b = red[1 * 32 + tid] + red[3 * 32 + tid] + red[5 * 32 + tid] + red[7 * 32 + tid];
          double* st = p.stats + ((size_t)n_img * d.Co_total + d.co_off + co0 + c0 + tid) * 2;
          atomicAdd(st, (double)a);
          atomicAdd(st + 1, (double)b);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      if (nchw) {
        if (valid) {
          float* y = reinterpret_cast<float*>(p.y);
          const int plane = d.Ho * d.Wo;
          const int n = opix / plane, pp = opix - n * plane;
#pragma unroll
          for (int jj = 0; jj < 32; ++jj)
            if (c0 + jj < cmax)
              y[((size_t)n * d.Co_total + d.co_off + co0 + c0 + jj) * plane + pp] = apply_act(v[jj], d.act);
        }
      } else if (staged) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (c0 + g * 8 < cmax) {              // never write past this row's BN*2 bytes
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = apply_act(v[g * 8 + e], d.act);
            *reinterpret_cast<uint4*>(stage_w + lane * spitch + (c0 + g * 8) * 2) = pack8(o);
          }
        }
      } else if (valid) {
        __nv_bfloat16* y = reinterpret_cast<__nv_bfloat16*>(p.y) + (size_t)opix * d.Co_total + d.co_off + co0 + c0;
#pragma unroll
        for (int e = 0; e < 32; ++e)
          if (c0 + e < cmax) {
            float val = apply_act(v[e], d.act);
            if (accum) val += __bfloat162float(y[e]);
            y[e] = __float2bfloat16_rn(val);
          }
      }
    }
    if (staged) {
      __syncwarp();
      const int cpr = cmax >> 3;                                    // 16-byte chunks per row
      __nv_bfloat16* ybase = reinterpret_cast<__nv_bfloat16*>(p.y) + d.co_off + co0;
      for (int idx = lane; idx < 32 * cpr; idx += 32) {
        const int r = idx / cpr, q = idx - r * cpr;
        const int grow = warp * 32 + r;
        if (rowinfo[grow].w == 0) continue;
        uint4 val = *reinterpret_cast<const uint4*>(stage_w + r * spitch + q * 16);
        __nv_bfloat16* dst = ybase + (size_t)rowout[grow] * d.Co_total + q * 8;
        if (accum) {
          float a[8], b[8];
          unpack8(val, a);
          unpack8(*reinterpret_cast<const uint4*>(dst), b);
#pragma unroll
          for (int e = 0; e < 8; ++e) a[e] += b[e];
          val = pack8(a);
        }
        *reinterpret_cast<uint4*>(dst) = val;
      }
    }
    tc_fence_before();
  } else {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      // instruction descriptor: D=f32, A=B=bf16, both K-major, N=BN, M=128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      for (int kb = 0; kb < p.nkb; ++kb) {
        const int s = kb % TC_STAGES;
        mbar_wait(full_bar(s), (kb / TC_STAGES) & 1);
        tc_fence_after();
        const uint64_t da = make_sw128_desc(sA + s * A_STAGE_BYTES);
        const uint64_t db = make_sw128_desc(sB + s * b_stage_bytes);
#pragma unroll
        for (int k4 = 0; k4 < TC_BK / 16; ++k4)   // +32 B per UMMA_K step inside the 128 B swizzle row
          umma_bf16(tmem_base, da + (uint64_t)(k4 * 2), db + (uint64_t)(k4 * 2), idesc, (kb | k4) != 0);
        umma_commit(empty_bar(s));
      }
      umma_commit(accum_bar);
    }
    __syncwarp();
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

inline int pick_bn(int Cout) {
  int tiles = (Cout + 127) / 128;
  int per = (Cout + tiles - 1) / tiles;
  int bn = (per + 15) / 16 * 16;
  return bn < 16 ? 16 : bn;
}
inline size_t tc_smem_bytes(int BN) {
  return (size_t)TC_STAGES * (A_STAGE_BYTES + BN * TC_BK * 2) + 64 + TC_BM * (sizeof(int4) + sizeof(int)) + 1024;
}

}  // namespace

bool conv2d_tc_supported(const msg_conv_desc* d, const void* x, const void* w, const void* y) {
  if (d->dtype != MSG_BF16) return false;
  if (d->flags & MSG_CONV_IN_NORM) return false;
  if ((d->Cin | d->Ci_total | d->ci_off) & 7) return false;
  if (((uintptr_t)x | (uintptr_t)w) & 15) return false;
  const long long hw = (long long)d->Hg * d->Wg;
  if ((d->flags & MSG_CONV_STATS) && (hw % TC_BM) != 0) return false;
  if (!(d->flags & MSG_CONV_OUT_NCHW_F32) && ((uintptr_t)y & 15)) return false;
  if ((d->flags & MSG_CONV_OUT_NCHW_F32) && d->Cout > 128) return false;
  return true;
}

int conv2d_tc(const msg_conv_desc* d, const void* x, const void* w, const float* bias, void* y,
              double* stats, const double* in_stats, cudaStream_t st) {
  (void)in_stats;
  TcParams p;
  p.d = *d;
  p.x = (const __nv_bfloat16*)x;
  p.w = (const __nv_bfloat16*)w;
  p.bias = bias;
  p.y = y;
  p.stats = stats;
  p.BN = pick_bn(d->Cout);
  p.tmem_cols = 32;
  while (p.tmem_cols < p.BN) p.tmem_cols <<= 1;
  p.K = d->KH * d->KW * d->Cin;
  p.nkb = (p.K + TC_BK - 1) / TC_BK;
  p.M = (long long)d->N * d->Hg * d->Wg;
  const size_t smem = tc_smem_bytes(p.BN);
  static DeviceOnce attr_set;     // cudaFuncSetAttribute is per device
  if (attr_set.needed()) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    MSG_REQUIRE(e == cudaSuccess, MSG_ERR_CUDA, "conv_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set.done();
  }
  dim3 grid((unsigned)((p.M + TC_BM - 1) / TC_BM), (unsigned)((d->Cout + p.BN - 1) / p.BN));
  conv_tc_kernel<<<grid, TC_THREADS, smem, st>>>(p);
  return check_launch("conv_tc_kernel");
}

}  // namespace msg
