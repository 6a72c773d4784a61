// MultiScaleBlock branches at C = 64 (enhanced_generator.py:52-71: 1x1 | 3x3 dil 1 | 3x3 dil 2 | 3x3 dil 4, 64 -> 4 x 16 channels)
// as a ROW RING of tensor-memory accumulators (sm_100a, bf16 operands, fp32 accumulate).
//
// The per-tap row-slab kernel (conv_slab.cu) issues 25 MMAs of N = 16 per K step and output row; a tcgen05.mma with M = 128
// costs ~40 cycles however small N is (its A operand is read from shared memory at 128 B/clk), so that kernel is shared-memory
// bound at 15-25 % of the tensor pipe, and it loads seven input row slabs per output row.  Here a CTA walks DOWN a 128-pixel
// column strip of one image:
//   * every input row slab [136 pixels x 64 ch] is loaded ONCE (one TMA box, zero fill = the convs' padding);
//   * for a horizontal shift sx the three vertical taps of a dilated 3x3 branch send input row r to output rows r - d, r, r + d,
//     whose accumulators sit in ADJACENT tensor-memory columns -- every branch keeps one ring of 16-column row accumulators per
//     residue class (row mod d) -- so they are ONE MMA of N = 48 over a 48-row weight stack [ky = 2 | ky = 1 | ky = 0]:
//     10 MMAs per K step and row instead of 25, each the A-read of one;
//   * the row finished by input row r (branch 1: r, branch 2: r - 1, branch 3: r - 2, branch 4: r - 4) is drained by the
//     epilogue warps (bias, IN statistics, bf16, 32-byte channel slice of its own output row) and its slot is ZEROED with
//     tcgen05.st, so every MMA accumulates and the first touch of a slot needs no special case.
// The schedule is stated (and run on tensors) in slab.py: ring_col / ring_row_mmas; this file restates it.
//
//   warp 0      TMA producer: the resident weight stacks once, then one slab per input row
//   warps 1-3   MMA issuers (warp 1 also allocates TMEM), at most two rows ahead of the epilogue (the 1x1 branch's ring has two
//               slots).  An issuing thread pays ~13 cycles per instruction (profiles/r1_mma_rate_microbench.md) and this schedule needs
//               ~6 per MMA (ring slots and runs change every row), so ONE issuer ran the MMAs at ~85 cycles each; three issue in
//               parallel: branches 1 + 2 | branch 3 | branch 4, whose accumulators are disjoint column ranges
//   warps 4-19  epilogue: four groups of four warps, group b drains the rows of branch b (thread = strip pixel = TMEM lane);
//               one group doing all four pieces of a step took ~5000 cycles per row against ~1600 of MMAs
#include <cuda.h>

#include "common.cuh"
#include "tcgen05.cuh"

namespace msg {
namespace {
using namespace tc;

constexpr int BM = 128;                 // strip width (output pixels)
constexpr int HALO = 4;
constexpr int SLAB_PX = BM + 2 * HALO;  // 136
constexpr int SLAB_BYTES = SLAB_PX * 128;          // 17408 = 17 * 1024
constexpr int W_ROWS = 16 + 9 * 48;                // 448 weight rows (msb64_ring_weights)
constexpr int W_BYTES = W_ROWS * 128;              // 57344
constexpr int NISSUE = 3;              // MMA issuing warps (branches 1+2 | 3 | 4: disjoint TMEM columns)
constexpr int EPI0 = 1 + NISSUE;       // first epilogue warp (a multiple of 4: warp & 3 = TMEM lane quarter)
constexpr int NTHREADS = 32 * (EPI0 + 16);      // TMA, 3 MMA issuers, 4 x 4 epilogue warps (group b drains branch b)

struct RingParams {
  int N, H, W, Co_total, co_off;
  int segs;
  long long total_rows;           // N * segs * H
  int stages;
  const float* bias;
  __nv_bfloat16* y;
  double* stats;
};

// TMEM column of the accumulator of output row y (>= 0) of branch b: slab.py ring_col
__device__ __forceinline__ uint32_t ring_col(int b, int y) {
  switch (b) {
    case 0: return (uint32_t)((y & 1) * 16);
    case 1: return 32u + (uint32_t)((y % 5) * 16);
    case 2: return 112u + (uint32_t)(((y & 1) * 4 + ((y >> 1) & 3)) * 16);
    default: return 240u + (uint32_t)(((y & 3) * 4 + ((y >> 2) & 3)) * 16);
  }
}

__global__ void __launch_bounds__(NTHREADS, 1)
msb64_ring_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                  const __grid_constant__ CUtensorMap mapY, const RingParams p) {
  extern __shared__ uint8_t smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = p.stages;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sB = base;                               // weight stacks [448][128 B], SW128
  const uint32_t sA = sB + W_BYTES;                       // slab ring
  const uint32_t sSum = sA + S * SLAB_BYTES;              // per-warp running column sums [16 warps][2][16] (hi, lo) pairs
  const uint32_t sTr = sSum + 16 * 32 * 8;                // per-warp transpose scratch [16][16][33] floats
  const uint32_t sBias = sTr + 16 * 528 * 4;              // bias [64]
  const uint32_t sOut = (sBias + 256 + 127u) & ~127u;     // per-warp output staging [16 warps][2][32 px][32 B]
  const uint32_t sBar = (sOut + 16 * 2 * 1024 + 7u) & ~7u;
  float* sbias = reinterpret_cast<float*>(gen + (sBias - base));
  auto full_bar = [&](int s) { return sBar + 8u * s; };
  auto empty_bar = [&](int s) { return sBar + 8u * (S + s); };
  const uint32_t wres_bar = sBar + 8u * (2 * S);
  auto rowdone_bar = [&](int k) { return sBar + 8u * (2 * S + 1 + (k & 3)); };
  auto drained_bar = [&](int k) { return sBar + 8u * (2 * S + 5 + (k & 3)); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + (sBar - base) + 8 * (2 * S + 9));

  if (tid < 64) sbias[tid] = p.bias ? p.bias[tid] : 0.f;
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), NISSUE); }
      mbar_init(wres_bar, 1);
      for (int k = 0; k < 4; ++k) { mbar_init(rowdone_bar(k), NISSUE); mbar_init(drained_bar(k), 16); }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapB)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapY)) : "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp >= EPI0 && warp < EPI0 + 4) {  // every accumulator starts at zero: each MMA of the kernel accumulates
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t z[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) z[i] = 0u;
    for (int c = 0; c < 512; c += 32) tmem_st32(tmem_base + lane_addr + (uint32_t)c, z);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  // Work = the rows of all column strips laid end to end (strip = (image, 128-pixel segment), H rows each); CTA i takes the i-th
  // equal share of that range, i.e. at most a few pieces of consecutive strips: every SM gets the same number of rows (whole-strip
  // segments left 20 of 148 SMs idle at 16 x 512 x 512) and a piece re-reads only its 8 halo rows.
  const long long g_lo = (long long)blockIdx.x * p.total_rows / gridDim.x, g_hi = (long long)(blockIdx.x + 1) * p.total_rows / gridDim.x;
  auto item = [&](long long g, int& img, int& seg, int& y0, int& y1) {
    const int strip = (int)(g / p.H);
    y0 = (int)(g - (long long)strip * p.H);
    const long long left = g_hi - g;
    y1 = (long long)(p.H - y0) < left ? p.H : y0 + (int)left;
    img = strip / p.segs;
    seg = strip - img * p.segs;
  };

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      mbar_expect_tx(wres_bar, (uint32_t)W_BYTES);
      for (int r = 0; r < W_ROWS; r += 64) tma_load_2d(sB + r * 128, &mapB, wres_bar, 0, r);
      int s = 0;
      uint32_t n = 0;                      // slabs issued
      for (long long g = g_lo; g < g_hi;) {
        int img, seg, y0, y1;
        item(g, img, seg, y0, y1);
        g += y1 - y0;
        const int r_lo = y0 - HALO < 0 ? 0 : y0 - HALO, r_hi = y1 + HALO > p.H ? p.H : y1 + HALO;
        for (int r = r_lo; r < r_hi; ++r, ++n) {
          if (n >= (uint32_t)S) mbar_wait(empty_bar(s), ((n / S) - 1) & 1);
          mbar_expect_tx(full_bar(s), (uint32_t)SLAB_BYTES);
          tma_load_4d(sA + s * SLAB_BYTES, &mapA, full_bar(s), 0, seg * BM - HALO, r, img);
          if (++s == S) s = 0;
        }
      }
    }
  } else if (warp < EPI0) {
    // ===================================== MMA issuers =====================================
    const int ib = warp - 1;               // 0: branches 1 (1x1) and 2; 1: branch 3; 2: branch 4
    const bool leader = elect_one();
    const uint32_t hi = (uint32_t)(make_sw128_desc(0) >> 32);
    const uint32_t idesc16 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    const uint32_t b_base = sB >> 4;
    int s = 0;
    uint32_t n = 0, k = 0;                 // slabs consumed, global step count
    mbar_wait(wres_bar, 0);
    for (long long g = g_lo; g < g_hi;) {
      int img, seg, y0, y1;
      item(g, img, seg, y0, y1);
      g += y1 - y0;
      for (int r = y0 - HALO; r < y1 + HALO; ++r, ++k) {
        // the epilogue must have drained (and zeroed) everything up to step k - 2: the slot this row first touches in the
        // two-slot ring of the 1x1 branch belonged to output row r - 2 (the other rings have more slack, slab.py)
#if defined(RING_EXP) && RING_EXP == 4
        if (k >= 1) mbar_wait(drained_bar((int)(k - 1)), ((k - 1) >> 2) & 1);
#elif !defined(RING_EXP) || RING_EXP != 1
        if (k >= 2) mbar_wait(drained_bar((int)(k - 2)), ((k - 2) >> 2) & 1);
#endif
        if (r >= 0 && r < p.H) {
          mbar_wait(full_bar(s), (n / S) & 1);
          tc_fence_after();
#if defined(RING_EXP) && RING_EXP == 3
          if (false) {
#else
          if (leader) {
#endif
            const uint32_t a0 = (sA + s * SLAB_BYTES) >> 4;
            if (ib == 0 && r >= y0 && r < y1) {       // branch 1: the 1x1 conv, output row r
              const uint32_t dcol = tmem_base + ring_col(0, r);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) umma_bf16_lo(dcol, a0 + (uint32_t)(HALO * 8 + 2 * ks), b_base + (uint32_t)(2 * ks), hi, idesc16, true);
            }
            {
              const int b = ib + 1;
              const int d = 1 << (b - 1);
              // entries e = 0, 1, 2 = output rows r - d, r, r + d (vertical taps ky = 2, 1, 0); runs of adjacent ring slots merge
              uint32_t col[3];
              bool ok[3];
#pragma unroll
              for (int e = 0; e < 3; ++e) {
                const int y = r + (e - 1) * d;
                ok[e] = y >= y0 && y < y1;
                col[e] = ok[e] ? ring_col(b, y) : 0u;
              }
              int e = 0;
              while (e < 3) {
                if (!ok[e]) { ++e; continue; }
                int nrun = 1;
                while (e + nrun < 3 && ok[e + nrun] && col[e + nrun] == col[e + nrun - 1] + 16u) ++nrun;
                const uint32_t idesc = (idesc16 & ~(0x3fu << 17)) | ((uint32_t)((16 * nrun) >> 3) << 17);
                const uint32_t dcol = tmem_base + col[e];
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                  const uint32_t wrow = (uint32_t)(16 + ((b - 1) * 3 + kx) * 48 + 16 * e);
                  const uint32_t av = a0 + (uint32_t)((HALO + (kx - 1) * d) * 8);
#pragma unroll
                  for (int ks = 0; ks < 4; ++ks) umma_bf16_lo(dcol, av + (uint32_t)(2 * ks), b_base + wrow * 8u + (uint32_t)(2 * ks), hi, idesc, true);
                }
                e += nrun;
              }
            }
            umma_commit(empty_bar(s));
          }
#if defined(RING_EXP) && RING_EXP == 3
          if (leader) umma_commit(empty_bar(s));
#endif
          __syncwarp();
          if (++s == S) s = 0;
          ++n;
        }
        if (leader) umma_commit(rowdone_bar((int)k));     // arrives when every MMA issued so far has completed
        __syncwarp();
      }
    }
  } else {
    // ===================================== epilogue (warps 4-19) =====================================
    const int q = warp & 3;
    const int b = (warp - EPI0) >> 2;                 // this group's branch
    const int lag = b == 0 ? 0 : (1 << (b - 1));      // input row r completes output row r - lag of branch b
    const int row = q * 32 + lane;                    // strip pixel = TMEM lane
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    float2* wsum = reinterpret_cast<float2*>(gen + (sSum - base)) + (warp - EPI0) * 32;     // [2][16]
    float* tr = reinterpret_cast<float*>(gen + (sTr - base)) + (warp - EPI0) * 528;         // [16][33]
    wsum[lane] = make_float2(0.f, 0.f);
    __syncwarp();
    const bool do_stats = p.stats != nullptr;
    int stat_img = -1;
    auto flush_stats = [&]() {
      if (stat_img >= 0 && lane < 16) {
        double* st = p.stats + ((size_t)stat_img * p.Co_total + p.co_off + 16 * b + lane) * 2;
        atomicAdd(st, f2sum_value(wsum[lane]));
        atomicAdd(st + 1, f2sum_value(wsum[16 + lane]));
        wsum[lane] = make_float2(0.f, 0.f); wsum[16 + lane] = make_float2(0.f, 0.f);
      }
      __syncwarp();
    };
    uint32_t zero16[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) zero16[i] = 0u;
    float bia[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) bia[c] = sbias[16 * b + c];
    uint32_t k = 0, nst = 0;               // steps, pieces stored by this warp
    for (long long g = g_lo; g < g_hi;) {
      int img, seg, y0, y1;
      item(g, img, seg, y0, y1);
      g += y1 - y0;
      const int xcol = seg * BM + row;
      const bool valid = xcol < p.W;
      if (do_stats && img != stat_img) { flush_stats(); stat_img = img; }
      for (int r = y0 - HALO; r < y1 + HALO; ++r, ++k) {
        const int y = r - lag;                                       // the row this input row completed for this branch
        // every warp follows every step (also those that finish nothing of its branch): a warp that ran ahead would arrive on a
        // drained barrier whose previous phase is still open
        mbar_wait(rowdone_bar((int)k), (k >> 2) & 1);
#if defined(RING_EXP) && RING_EXP == 2
        if (false) {
#else
        if (y >= y0 && y < y1) {                                     // (warp-uniform)
#endif
          tc_fence_after();
          const uint32_t taddr = tmem_base + lane_addr + ring_col(b, y);
          float v[16];
          tmem_ld16(taddr, v);
          tmem_ld_wait();
#if !defined(RING_EXP) || (RING_EXP != 5 && RING_EXP != 6 && RING_EXP != 7)
          tmem_st16(taddr, zero16);                                  // the slot is free for the row that wraps onto it
#endif
#pragma unroll
          for (int c = 0; c < 16; ++c) v[c] += bia[c];
          {
            // the warp's [32 pixels x 16 channels] piece: 32 bytes per thread into its staging buffer, then ONE TMA store (the tensor
            // map clips pixels beyond the plane).  Per-thread global stores -- 32 scattered sectors per warp instruction, each 128-byte
            // line assembled from four branches at four different times -- held the whole kernel back: 0.73 ms with two 16-byte
            // stores per thread, 0.63 with one 32-byte store, MMA-bound (0.42) without stores.
            float lo[8], hi8[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) { lo[c] = v[c]; hi8[c] = v[8 + c]; }
            uint8_t* stg = gen + (sOut - base) + ((warp - EPI0) * 2 + (int)(nst & 1)) * 1024;
            if (nst >= 2) {                                          // the store issued two pieces ago has read this buffer
              if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
              __syncwarp();
            }
            *reinterpret_cast<uint4*>(stg + lane * 32) = pack8(lo);
            *reinterpret_cast<uint4*>(stg + lane * 32 + 16) = pack8(hi8);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_4d(&mapY, smem_u32(stg), p.co_off + 16 * b, seg * BM + q * 32, y, img);
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            ++nst;
          }
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(drained_bar((int)k));
          if (do_stats) {
            // (after the slot has been handed back) transpose the warp's 32 x 16 block through smem; lanes 0-15 sum one column
            // each over the 32 rows in a fixed order on two chains
#pragma unroll
            for (int c = 0; c < 16; ++c) tr[c * 33 + lane] = valid ? v[c] : 0.f;
            __syncwarp();
            if (lane < 16) {
              float cs0 = 0.f, cs1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
              for (int rr = 0; rr < 32; rr += 2) {
                const float x0 = tr[lane * 33 + rr], x1 = tr[lane * 33 + rr + 1];
                cs0 += x0; cs1 += x1;
                q0 = fmaf(x0, x0, q0); q1 = fmaf(x1, x1, q1);
              }
              f2sum_add(wsum[lane], cs0 + cs1);
              f2sum_add(wsum[16 + lane], q0 + q1);
            }
            __syncwarp();
          }
        } else {
          if (lane == 0) mbar_arrive(drained_bar((int)k));           // nothing of this branch finishes at this step
        }
      }
    }
    if (do_stats) flush_stats();
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace
}  // namespace msg

using namespace msg;

extern "C" int msg_msb64_ring(const msg_msb_ring_desc* d, const void* x, const void* w_stacks, const float* bias, void* y,
                              double* stats, void* stream) {
  MSG_REQUIRE(d != nullptr && x && w_stacks && y, MSG_ERR_SHAPE, "msb64_ring: null argument");
  MSG_REQUIRE(d->dtype == MSG_BF16, MSG_ERR_UNSUPPORTED, "msb64_ring: bf16 only");
  MSG_REQUIRE(d->N > 0 && d->H > 0 && d->W > 0, MSG_ERR_SHAPE, "msb64_ring: bad plane");
  MSG_REQUIRE((d->Ci_total & 7) == 0 && (d->ci_off & 7) == 0 && d->ci_off + 64 <= d->Ci_total, MSG_ERR_SHAPE, "msb64_ring: input channel layout");
  MSG_REQUIRE((d->Co_total & 7) == 0 && (d->co_off & 7) == 0 && d->co_off + 64 <= d->Co_total, MSG_ERR_SHAPE, "msb64_ring: output channel layout");
  MSG_REQUIRE((((uintptr_t)x | (uintptr_t)w_stacks | (uintptr_t)y) & 15) == 0, MSG_ERR_ALIGN, "msb64_ring: operands must be 16-byte aligned");
  EncodeTiledFn enc = get_encode();
  MSG_REQUIRE(enc != nullptr, MSG_ERR_CUDA, "msb64_ring: cuTensorMapEncodeTiled unavailable");

  RingParams p;
  p.N = d->N; p.H = d->H; p.W = d->W; p.Co_total = d->Co_total; p.co_off = d->co_off;
  p.bias = bias; p.y = reinterpret_cast<__nv_bfloat16*>(y); p.stats = (d->flags & MSG_CONV_STATS) ? stats : nullptr;
  MSG_REQUIRE(!(d->flags & MSG_CONV_STATS) || stats != nullptr, MSG_ERR_SHAPE, "msb64_ring: stats buffer missing");
  p.segs = (d->W + BM - 1) / BM;
  const int sms = sm_count();
  p.total_rows = (long long)d->N * p.segs * d->H;
  MSG_REQUIRE(p.total_rows < (1LL << 40), MSG_ERR_SHAPE, "msb64_ring: too many rows");
  const int fixed = W_BYTES + 16 * 32 * 8 + 16 * 528 * 4 + 256 + 128 + 16 * 2 * 1024 + 8 + 512 + 1024;
  int stages = (220 * 1024 - fixed) / SLAB_BYTES;
  if (stages > 8) stages = 8;
  p.stages = stages;
  const size_t smem = (size_t)stages * SLAB_BYTES + fixed;

  CUtensorMap mapA, mapB, mapY;
  {
    cuuint64_t dims[4] = {(cuuint64_t)d->Co_total, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->N};
    cuuint64_t strides[3] = {(cuuint64_t)d->Co_total * 2, (cuuint64_t)d->W * d->Co_total * 2, (cuuint64_t)d->H * d->W * d->Co_total * 2};
    cuuint32_t box[4] = {16, 32, 1, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&mapY, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, y, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MSG_REQUIRE(r == CUDA_SUCCESS, MSG_ERR_CUDA, "msb64_ring: cuTensorMapEncodeTiled(y) failed with %d", (int)r);
  }
  {
    cuuint64_t dims[4] = {64, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->N};
    cuuint64_t strides[3] = {(cuuint64_t)d->Ci_total * 2, (cuuint64_t)d->W * d->Ci_total * 2, (cuuint64_t)d->H * d->W * d->Ci_total * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)SLAB_PX, 1, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    void* b0 = (void*)((const __nv_bfloat16*)x + d->ci_off);
    CUresult r = enc(&mapA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, b0, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MSG_REQUIRE(r == CUDA_SUCCESS, MSG_ERR_CUDA, "msb64_ring: cuTensorMapEncodeTiled(x) failed with %d", (int)r);
  }
  {
    cuuint64_t dims[2] = {64, (cuuint64_t)W_ROWS};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {64, 64};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&mapB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)w_stacks, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MSG_REQUIRE(r == CUDA_SUCCESS, MSG_ERR_CUDA, "msb64_ring: cuTensorMapEncodeTiled(w) failed with %d", (int)r);
  }
  static DeviceOnce attr_set;
  if (attr_set.needed()) {
    cudaError_t e = cudaFuncSetAttribute(msb64_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    MSG_REQUIRE(e == cudaSuccess, MSG_ERR_CUDA, "msb64_ring: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set.done();
  }
  long long grid_ll = p.total_rows / 8;       // at least 8 rows per CTA (each piece re-reads 8 halo rows)
  int grid = grid_ll < 1 ? 1 : (grid_ll > sms ? sms : (int)grid_ll);
  msb64_ring_kernel<<<grid, NTHREADS, smem, as_stream(stream)>>>(mapA, mapB, mapY, p);
  return check_launch("msb64_ring_kernel");
}
