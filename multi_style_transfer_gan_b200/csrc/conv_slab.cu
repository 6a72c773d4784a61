// "Row-slab" tcgen05 convolution for stride-1, same-size convs with small N (sm_100a, bf16).
//
// The 7x7 output conv (N=3), the 7x7 input conv (Cin=3) and the MultiScaleBlock's 1x1 + three dilated
// 3x3 branches (N = C/4 each) are hopeless as plain implicit GEMMs: every filter tap re-fetches a full
// 128 x 64 A tile for a 128 x 16 MMA, so they run at the L2->SM fill rate, not at the tensor core's.
// Here a K block is not a tap but an INPUT ROW SLAB: [8 channel chunks][128 + 2*halo pixels][16 B],
// written by ONE 5-D TMA box in the no-swizzle "core-matrix" layout (16-byte rows, 8-row groups 128 B
// apart, K chunks LBO apart).  In that layout a horizontal tap shift sx is just a +16*sx byte offset
// of the shared-memory descriptor's start address, so all KW taps of a filter row -- and all four
// MultiScaleBlock branches, whose dilations only change the row (dy) and shift (sx) -- reuse the same
// slab: A traffic drops from (taps x tile) to (filter rows x tile), and the four branches become one
// kernel writing the concatenated [.., C] tensor (each branch accumulates into its own TMEM column
// slice).  Zero padding = TMA out-of-bounds fill.  For the 3-channel image (padded to 8 channels = one
// 16-byte chunk per pixel) the K=16 of one MMA spans two ADJACENT PIXELS (LBO = 16 B), so a 7-tap
// filter row is 4 MMAs.
//
// Pipeline = conv_tma.cu's: warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 epilogue, smem ring of
// slabs, double-buffered TMEM accumulators, persistent CTAs (one per SM).
#include <cuda.h>

#include "common.cuh"

namespace msg {
namespace {

constexpr int BM = 128;
constexpr int NTHREADS = 192;
constexpr int STAGE_PITCH = 128 + 16;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// B operand: K-major, 128-byte swizzle (rows of 64 bf16)
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// A operand: K-major, no swizzle: 8-row x 16-byte core matrices; SBO = 128 B between 8-row groups
// (rows are 16 B apart), LBO = byte distance between the two K chunks of one UMMA_K=16 step.
__device__ __forceinline__ uint64_t make_noswz_desc(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(128 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float warp_transpose_reduce32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = lane & off;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      float send = up ? v[i] : v[i + off];
      float keep = up ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

struct SlabParams {
  msg_slab_desc d;
  const float* bias;
  void* y;
  double* stats;
  int stages, tmem_cols;
  int Ws;             // slab width in pixels (128 + 2*halo, padded to a multiple of 8)
  int a_bytes;        // bytes of one A slab (chunks * Ws * 16)
  int b_bytes;        // bytes of one B stage (max taps per k-block * ncols * 128)
  int chunks;         // K chunks per slab (8, or 1 in pixel-pair mode)
  int segs;           // 128-pixel segments per image row
};

__global__ void __launch_bounds__(NTHREADS, 1)
conv_slab_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                 const SlabParams p) {
  extern __shared__ uint8_t smem_raw[];
  const msg_slab_desc& d = p.d;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = p.stages;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sB = base;                                   // B stages first: 1024-byte aligned sub-tiles
  const uint32_t sA = sB + S * p.b_bytes;
  const uint32_t sStage = (sA + S * p.a_bytes + 127u) & ~127u;
  const uint32_t sRed = sStage + 4 * 32 * STAGE_PITCH;
  const uint32_t sBar = sRed + 1024;
  uint8_t* stage_gen = gen + (sStage - base);
  float* red = reinterpret_cast<float*>(gen + (sRed - base));
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + (sBar - base) + 8 * (2 * S + 4));
  auto full_bar = [&](int s) { return sBar + 8u * s; };
  auto empty_bar = [&](int s) { return sBar + 8u * (S + s); };
  auto tfull_bar = [&](int b) { return sBar + 8u * (2 * S + b); };
  auto tempty_bar = [&](int b) { return sBar + 8u * (2 * S + 2 + b); };

  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
      for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), 4); }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapB)) : "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int total_tiles = d.N * d.H * p.segs;
  const int tap_bytes = d.ncols * 128;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      uint32_t it = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int seg = t % p.segs, ny = t / p.segs;
        const int yrow = ny % d.H, img = ny / d.H;
        const int x0 = seg * BM - d.halo;
        for (int kb = 0; kb < d.n_kblocks; ++kb, ++it) {
          const int s = it % S;
          if (it >= (uint32_t)S) mbar_wait(empty_bar(s), ((it / S) - 1) & 1);
          const int t0 = d.kb_tap_begin[kb], t1 = d.kb_tap_begin[kb + 1];
          const int nb = d.pixel_pair_k ? 1 : (t1 - t0);
          mbar_expect_tx(full_bar(s), p.a_bytes + nb * tap_bytes);
          // slab = `chunks` planes of [Ws pixels][16 B]: one box {8 ch, Ws, 1, 1} per 8-channel chunk
          for (int ch = 0; ch < p.chunks; ++ch)
            tma_load_4d(sA + s * p.a_bytes + ch * (p.Ws * 16), &mapA, full_bar(s), d.kb_cb[kb] * 64 + ch * 8, x0,
                        yrow + d.kb_dy[kb], img);
          if (d.pixel_pair_k) {
            tma_load_2d(sB + s * p.b_bytes, &mapB, full_bar(s), 0, kb * d.ncols);
          } else {
            for (int tp = t0; tp < t1; ++tp)
              tma_load_2d(sB + s * p.b_bytes + (tp - t0) * tap_bytes, &mapB, full_bar(s), 0, tp * d.ncols);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =====================================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(d.ncols >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      const uint32_t lbo = d.pixel_pair_k ? 16u : (uint32_t)(p.Ws * 16);
      const int ksteps = d.pixel_pair_k ? 1 : 4;              // UMMA_K=16 steps per tap
      uint32_t it = 0, lt = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++lt) {
        const int buf = lt & 1;
        if (lt >= 2) mbar_wait(tempty_bar(buf), ((lt >> 1) - 1) & 1);
        tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(buf * d.Ntot);
        for (int kb = 0; kb < d.n_kblocks; ++kb, ++it) {
          const int s = it % S;
          mbar_wait(full_bar(s), (it / S) & 1);
          tc_fence_after();
          const uint32_t a0 = sA + s * p.a_bytes;
          const uint32_t b0 = sB + s * p.b_bytes;
          const int t0 = d.kb_tap_begin[kb], t1 = d.kb_tap_begin[kb + 1];
          for (int tp = t0; tp < t1; ++tp) {
            const uint32_t a_tap = a0 + (uint32_t)((d.halo + d.tap_sx[tp]) * 16);
            const uint32_t b_tap = b0 + (d.pixel_pair_k ? 0u : (uint32_t)((tp - t0) * tap_bytes));
            for (int ks = 0; ks < ksteps; ++ks) {
              // 64-channel mode: K step ks = chunks 2ks, 2ks+1 (LBO apart); B advances 32 B in its 128 B row.
              // pixel-pair mode: one K step = this pixel + the next one (LBO = 16 B); B row holds 4 taps' K.
              const uint64_t da = make_noswz_desc(a_tap + (uint32_t)(ks * 2) * lbo, lbo);
              const uint64_t db = make_sw128_desc(b_tap) + (uint64_t)((d.pixel_pair_k ? d.tap_kstep[tp] : ks) * 2);
              umma_bf16(tacc + (uint32_t)d.tap_acc_col[tp], da, db, idesc, !(d.tap_first[tp] && ks == 0));
            }
          }
          umma_commit(empty_bar(s));
        }
        umma_commit(tfull_bar(buf));
      }
    }
  } else {
    // ===================================== epilogue (warps 2-5) =====================================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const bool do_stats = d.flags & MSG_CONV_STATS;
    const bool nchw = d.flags & MSG_CONV_OUT_NCHW_F32;
    uint8_t* stage_w = stage_gen + q * (32 * STAGE_PITCH);
    const int etid = tid - 64;
    const int cmax = d.n_store;                   // columns actually stored (<= Ntot)
    const bool vec = !nchw && ((d.Co_total | d.co_off | cmax) & 7) == 0;
    uint32_t lt = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++lt) {
      const int seg = t % p.segs, ny = t / p.segs;
      const int yrow = ny % d.H, img = ny / d.H;
      const int buf = lt & 1;
      const int xcol = seg * BM + row;
      const bool valid = xcol < d.W;
      const int opix = (img * d.H + yrow) * d.W + (valid ? xcol : 0);
      mbar_wait(tfull_bar(buf), (lt >> 1) & 1);
      tc_fence_after();
      const uint32_t tacc = tmem_base + (uint32_t)(buf * d.Ntot) + ((uint32_t)(q * 32) << 16);
      for (int cg = 0; cg < cmax; cg += 64) {
        const int ncol = (cmax - cg) < 64 ? (cmax - cg) : 64;
        __syncwarp();
        float v[64];
        tmem_ld32(tacc + (uint32_t)cg, *reinterpret_cast<float(*)[32]>(&v[0]));
        if (ncol > 32) tmem_ld32(tacc + (uint32_t)(cg + 32), *reinterpret_cast<float(*)[32]>(&v[32]));
        tmem_ld_wait();
        if (cg + 64 >= cmax) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(buf));
        }
        if (p.bias != nullptr) {
#pragma unroll
          for (int jj = 0; jj < 64; ++jj)
            if (jj < ncol) v[jj] += __ldg(p.bias + cg + jj);
        }
        if (do_stats) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (h * 32 < ncol) {
              float s1[32], s2[32];
#pragma unroll
              for (int jj = 0; jj < 32; ++jj) { s1[jj] = valid ? v[h * 32 + jj] : 0.f; s2[jj] = s1[jj] * s1[jj]; }
              float cs = warp_transpose_reduce32(s1, lane);
              float css = warp_transpose_reduce32(s2, lane);
              red[(q * 2 + 0) * 32 + lane] = cs;
              red[(q * 2 + 1) * 32 + lane] = css;
              asm volatile("bar.sync 1, 128;" ::: "memory");
              if (etid < 32 && h * 32 + etid < ncol) {
                float a = red[0 * 32 + etid] + red[2 * 32 + etid] + red[4 * 32 + etid] + red[6 * 32 + etid];
                float b = red[1 * 32 + etid] + red[3 * 32 + etid] + red[5 * 32 + etid] + red[7 * 32 + etid];
                double* st = p.stats + ((size_t)img * d.Co_total + d.co_off + cg + h * 32 + etid) * 2;
                atomicAdd(st, (double)a);
                atomicAdd(st + 1, (double)b);
              }
              asm volatile("bar.sync 1, 128;" ::: "memory");
            }
          }
        }
        if (nchw) {
          if (valid) {
            float* y = reinterpret_cast<float*>(p.y);
            const int plane = d.H * d.W;
            const int pp = yrow * d.W + xcol;
#pragma unroll
            for (int jj = 0; jj < 64; ++jj)
              if (jj < ncol)
                y[((size_t)img * d.Co_total + d.co_off + cg + jj) * plane + pp] = apply_act(v[jj], d.act);
          }
        } else if (vec && (ncol == 64 || ncol == 32 || ncol == 16)) {
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            if (g * 8 < ncol) {
              float o[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) o[e] = apply_act(v[g * 8 + e], d.act);
              *reinterpret_cast<uint4*>(stage_w + lane * STAGE_PITCH + g * 16) = pack8(o);
            }
          }
          __syncwarp();
          const int lg = ncol == 64 ? 3 : (ncol == 32 ? 2 : 1);
          __nv_bfloat16* ybase = reinterpret_cast<__nv_bfloat16*>(p.y) + d.co_off + cg;
#pragma unroll
          for (int itn = 0; itn < 8; ++itn) {
            const int idx = itn * 32 + lane;
            if (idx < (32 << lg)) {
              const int r = idx >> lg, ch = idx & ((1 << lg) - 1);
              uint4 val = *reinterpret_cast<const uint4*>(stage_w + r * STAGE_PITCH + ch * 16);
              const int op = __shfl_sync(0xffffffffu, opix, r);
              const int ok = __shfl_sync(0xffffffffu, (int)valid, r);
              if (ok) *reinterpret_cast<uint4*>(ybase + (size_t)op * d.Co_total + ch * 8) = val;
            }
          }
          __syncwarp();
        } else if (valid) {
          __nv_bfloat16* y = reinterpret_cast<__nv_bfloat16*>(p.y) + (size_t)opix * d.Co_total + d.co_off + cg;
#pragma unroll
          for (int e = 0; e < 64; ++e)
            if (e < ncol) y[e] = __float2bfloat16_rn(apply_act(v[e], d.act));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

}  // namespace
}  // namespace msg

using namespace msg;

extern "C" int msg_conv_slab(const msg_slab_desc* d, const void* x, const void* w_slab, const float* bias,
                             void* y, double* stats, void* stream) {
  MSG_REQUIRE(d != nullptr, MSG_ERR_SHAPE, "conv_slab: null descriptor");
  MSG_REQUIRE(d->dtype == MSG_BF16, MSG_ERR_UNSUPPORTED, "conv_slab: bf16 only");
  MSG_REQUIRE(d->N > 0 && d->H > 0 && d->W > 0 && d->W % 8 == 0, MSG_ERR_SHAPE, "conv_slab: bad plane");
  MSG_REQUIRE(d->pixel_pair_k ? (d->Cin == 8 && d->Ci_total == 8 && d->ci_off == 0)
                              : (d->Cin % 64 == 0 && (d->Ci_total & 7) == 0 && (d->ci_off & 7) == 0),
              MSG_ERR_SHAPE, "conv_slab: channel layout unsupported");
  MSG_REQUIRE(d->ncols >= 16 && d->ncols % 16 == 0 && d->ncols <= 256 && d->Ntot % 16 == 0 && d->Ntot <= 256 &&
                  d->n_store <= d->Ntot,
              MSG_ERR_SHAPE, "conv_slab: bad N configuration");
  MSG_REQUIRE(d->n_kblocks >= 1 && d->n_kblocks <= MSG_SLAB_MAX_KBLOCKS && d->n_taps >= 1 && d->n_taps <= MSG_SLAB_MAX_TAPS &&
                  d->halo >= 0 && d->halo <= 16,
              MSG_ERR_SHAPE, "conv_slab: program too large");
  MSG_REQUIRE((((uintptr_t)x | (uintptr_t)w_slab) & 15) == 0, MSG_ERR_ALIGN, "conv_slab: operands must be 16-byte aligned");
  MSG_REQUIRE(!(d->flags & MSG_CONV_STATS) || stats != nullptr, MSG_ERR_SHAPE, "conv_slab: stats buffer missing");
  EncodeTiledFn enc = get_encode();
  MSG_REQUIRE(enc != nullptr, MSG_ERR_CUDA, "conv_slab: cuTensorMapEncodeTiled unavailable");

  SlabParams p;
  p.d = *d; p.bias = bias; p.y = y; p.stats = stats;
  p.chunks = d->pixel_pair_k ? 1 : 8;
  p.Ws = (BM + 2 * d->halo + (d->pixel_pair_k ? 1 : 0) + 7) / 8 * 8;
  MSG_REQUIRE(p.Ws <= 256, MSG_ERR_SHAPE, "conv_slab: halo too large");
  p.a_bytes = p.chunks * p.Ws * 16;
  int max_taps = 0;
  for (int kb = 0; kb < d->n_kblocks; ++kb) {
    int nt = d->kb_tap_begin[kb + 1] - d->kb_tap_begin[kb];
    MSG_REQUIRE(nt >= 1, MSG_ERR_SHAPE, "conv_slab: empty k-block");
    if (nt > max_taps) max_taps = nt;
  }
  p.b_bytes = (d->pixel_pair_k ? 1 : max_taps) * d->ncols * 128;
  p.segs = (d->W + BM - 1) / BM;
  p.tmem_cols = 32;
  while (p.tmem_cols < 2 * d->Ntot) p.tmem_cols <<= 1;
  const int stage_bytes = p.a_bytes + p.b_bytes;
  const int fixed = 128 + 4 * 32 * STAGE_PITCH + 1024 + 8 * 16 + 64 + 1024;
  int stages = (220 * 1024 - fixed) / stage_bytes;
  if (stages > 8) stages = 8;
  MSG_REQUIRE(stages >= 2, MSG_ERR_UNSUPPORTED, "conv_slab: stage of %d bytes does not fit twice in shared memory", stage_bytes);
  p.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + fixed;

  CUtensorMap mapA, mapB;
  {
    cuuint64_t dims[4] = {(cuuint64_t)d->Cin, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->N};
    cuuint64_t strides[3] = {(cuuint64_t)d->Ci_total * 2, (cuuint64_t)d->W * d->Ci_total * 2,
                             (cuuint64_t)d->H * d->W * d->Ci_total * 2};
    cuuint32_t box[4] = {8, (cuuint32_t)p.Ws, 1, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    void* base = (void*)((const __nv_bfloat16*)x + d->ci_off);
    CUresult r = enc(&mapA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MSG_REQUIRE(r == CUDA_SUCCESS, MSG_ERR_CUDA, "conv_slab: cuTensorMapEncodeTiled(A) failed with %d", (int)r);
  }
  {
    cuuint64_t dims[2] = {64, (cuuint64_t)(d->pixel_pair_k ? d->n_kblocks : d->n_taps) * d->ncols};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {64, (cuuint32_t)d->ncols};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&mapB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)w_slab, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MSG_REQUIRE(r == CUDA_SUCCESS, MSG_ERR_CUDA, "conv_slab: cuTensorMapEncodeTiled(B) failed with %d", (int)r);
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_slab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    MSG_REQUIRE(e == cudaSuccess, MSG_ERR_CUDA, "conv_slab: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  int grid = sm_count();
  const long long total = (long long)d->N * d->H * p.segs;
  MSG_REQUIRE(total < 0x7fffffffLL, MSG_ERR_SHAPE, "conv_slab: too many tiles");
  if (grid > total) grid = (int)total;
  conv_slab_kernel<<<grid, NTHREADS, smem, as_stream(stream)>>>(mapA, mapB, p);
  return check_launch("conv_slab_kernel");
}
