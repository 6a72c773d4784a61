// "Taps-as-N" row-slab convolution (sm_100a, bf16): stride-1 same-size convs whose per-tap N is tiny.
//
// A tcgen05.mma with M=128, K=16 costs ~39-45 cycles on B200 whatever N is when N <= 64 (tools/mma_rate.cu: the 4 KB
// A-operand read from shared memory dominates; ~90 with the generic issue loop around it), so issuing one N=16 MMA
// per filter tap (conv_slab.cu) leaves the tensor core ~5-10x under-used.  Here ALL taps that share an input-row slab are ONE MMA:
// the slab is the UNSHIFTED A operand [128 slab pixels x 64 ch], the taps' filters are stacked along N
// (7 taps x 4 padded filters = 32 columns for the 7x7 output conv; 1 + 3x3 (branch, sx) column groups
// of 16 = 160 columns for the MultiScaleBlock at C=64), giving partial products
//      D'[p][tap cols] = sum_c slab[p][c] * W_tap[c]        for every slab pixel p,
// and the horizontal tap shift is applied in the EPILOGUE:  out[x][c] = sum_taps D'[x + shift_tap][col_tap + c]
// through a shared-memory exchange tile (rows = slab pixels).  A tile yields 128 - 2*halo outputs.
// MMAs per tile: (filter rows) x (Cin/16) instead of (taps) x (Cin/16): 28 instead of 196 for the 7x7.
//
// Pipeline: warp 0 TMA producer (weights resident in smem, loaded once; one 128B-swizzled slab box per
// k-block), warp 1 MMA issuer (warp-uniform loop, elected lane), warps 2-5 epilogue, double-buffered TMEM.
// Multi-row tiles (msg_shift_desc::tile_rows, kb_same_slab, grp_row): the kernel moves ~24 B/clk/SM of slabs from L2, its
// ceiling, so a tile of two output rows loads each input row ONCE and feeds it to both rows' filter rows (7x7 conv: 8
// slab loads per 2 rows instead of 14); the output conv's tanh is MUFU.TANH (the epilogue sets the pace after that).
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "tcgen05.cuh"

namespace msg {
namespace {
using namespace tc;

constexpr int BM = 128;
constexpr int NTHREADS = 320;     // TMA, MMA, 2 x 4 epilogue warps (group g drains TMEM accumulator buffer g: every other tile)
constexpr int A_BYTES = BM * 128;


struct ShiftParams {
  msg_shift_desc d;
  const float* bias;
  void* y;
  double* stats;
  int stages, tmem_cols, b_bytes, b_rows, Wv, segs, ex_pitch;
  int groups;          // epilogue groups (1 or 2)
};

__global__ void __launch_bounds__(NTHREADS, 1)
conv_shift_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                  const ShiftParams p) {
  extern __shared__ uint8_t smem_raw[];
  const msg_shift_desc& d = p.d;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = p.stages;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sB = base;                                   // resident weights [b_rows][128 B], SW128
  const uint32_t sA = sB + p.b_bytes;                         // slab ring
  const uint32_t sEx = sA + S * A_BYTES;                      // exchange tiles [2 groups][128][Ntot + 4] floats
  const uint32_t sSum = (sEx + p.groups * BM * p.ex_pitch * 4 + 7u) & ~7u;  // per-warp column sums [8][2][64] (hi, lo) pairs
  const uint32_t sTr = sSum + 8 * 128 * 8;                    // per-warp transpose scratch [8][32][33]
  const uint32_t sBias = sTr + 8 * 1056 * 4;                  // bias [<= 256]
  const uint32_t sBar = (sBias + 1024 + 7u) & ~7u;
  float* ex_all = reinterpret_cast<float*>(gen + (sEx - base));
  float* sbias = reinterpret_cast<float*>(gen + (sBias - base));
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + (sBar - base) + 8 * (2 * S + 5));
  auto full_bar = [&](int s) { return sBar + 8u * s; };
  auto empty_bar = [&](int s) { return sBar + 8u * (S + s); };
  auto tfull_bar = [&](int b) { return sBar + 8u * (2 * S + b); };
  auto tempty_bar = [&](int b) { return sBar + 8u * (2 * S + 2 + b); };
  const uint32_t wres_bar = sBar + 8u * (2 * S + 4);

  if (p.bias != nullptr)
    for (int i = tid; i < d.n_out; i += NTHREADS) sbias[i] = p.bias[i];
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
      for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), 4); }
      mbar_init(wres_bar, 1);
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
  const int R = d.tile_rows > 1 ? d.tile_rows : 1;      // output rows per tile
  const int Hy = (d.H + R - 1) / R;
  const int total_tiles = d.N * Hy * p.segs;
  const int t_begin = (int)((long long)blockIdx.x * total_tiles / gridDim.x);
  const int t_end = (int)((long long)(blockIdx.x + 1) * total_tiles / gridDim.x);

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      mbar_expect_tx(wres_bar, (uint32_t)p.b_bytes);          // weights: once, boxes of 32 rows
      for (int r = 0; r < p.b_rows; r += 32) tma_load_2d(sB + r * 128, &mapB, wres_bar, 0, r);
      int s = 0;
      uint32_t ph = 0;
      bool wrapped = false;
      for (int t = t_begin; t < t_end; ++t) {
        const int seg = t % p.segs, ny = t / p.segs;
        const int yrow = (ny % Hy) * R, img = ny / Hy;
        const int x0 = seg * p.Wv - d.halo;
        for (int kb = 0; kb < d.n_kblocks; ++kb) {
          if (d.kb_same_slab[kb]) continue;                    // re-uses the previous k-block's slab
          if (wrapped) mbar_wait(empty_bar(s), ph ^ 1u);
          mbar_expect_tx(full_bar(s), A_BYTES);
          tma_load_4d(sA + s * A_BYTES, &mapA, full_bar(s), d.kb_cb[kb] * 64, x0, yrow + d.kb_dy[kb], img);
          if (++s == S) { s = 0; ph ^= 1u; wrapped = true; }   // no div/mod by the runtime stage count
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer (warp-uniform, elected lane issues) =============
    const uint32_t idesc0 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BM >> 4) << 24);
    const uint64_t sw128_hi = make_sw128_desc(0);
    // Single-warp serial issue: every instruction per K block is on the critical path.  Lane kb keeps the
    // tile-invariant operands of K block kb (<= 32 K blocks) in registers; the loop broadcasts them with
    // independent shuffles; stage / phase are counted incrementally (no div/mod by the runtime stage count).
    uint32_t kb_b = 0, kb_idesc = 0, kb_col = 0, kb_acc = 1, kb_same = 0, kb_last = 1;
    if (lane < d.n_kblocks) {
      kb_same = d.kb_same_slab[lane] != 0;                     // shares the slab (pipeline stage) of k-block lane-1
      kb_last = lane + 1 >= d.n_kblocks || d.kb_same_slab[lane + 1] == 0;   // last user of its slab: releases the stage
      kb_b = (sB + (uint32_t)d.kb_wrow[lane] * 128u) >> 4;
      kb_idesc = idesc0 | ((uint32_t)(d.kb_ncols[lane] >> 3) << 17);
      kb_col = (uint32_t)d.kb_col0[lane];
      kb_acc = !d.kb_first[lane];
    }
    const uint32_t a_base = sA >> 4;
    const uint32_t hi = (uint32_t)(sw128_hi >> 32);
    const bool leader = elect_one();          // one election for the whole kernel
    int s = 0;
    uint32_t ph = 0, lt = 0;
    mbar_wait(wres_bar, 0);
    for (int t = t_begin; t < t_end; ++t, ++lt) {
      const int buf = lt & 1;
      if (lt >= 2) mbar_wait(tempty_bar(buf), ((lt >> 1) - 1) & 1);
      tc_fence_after();
      const uint32_t tacc = tmem_base + (uint32_t)(buf * d.Ntot);
      for (int kb = 0; kb < d.n_kblocks; ++kb) {
        const uint32_t b_lo = __shfl_sync(0xffffffffu, kb_b, kb);
        const uint32_t idesc = __shfl_sync(0xffffffffu, kb_idesc, kb);
        const uint32_t dcol = tacc + __shfl_sync(0xffffffffu, kb_col, kb);
        const uint32_t acc = __shfl_sync(0xffffffffu, kb_acc, kb);
        const uint32_t same = __shfl_sync(0xffffffffu, kb_same, kb);
        const uint32_t last = __shfl_sync(0xffffffffu, kb_last, kb);
        const uint32_t a_lo = a_base + (uint32_t)s * (A_BYTES >> 4);
        if (!same) {
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
        }
        if (leader) {            // descriptors as {32-bit start-address word, constant high word}: 32-bit K-step adds
          umma_bf16_lo(dcol, a_lo, b_lo, hi, idesc, acc != 0);
          umma_bf16_lo(dcol, a_lo + 2, b_lo + 2, hi, idesc, true);
          umma_bf16_lo(dcol, a_lo + 4, b_lo + 4, hi, idesc, true);
          umma_bf16_lo(dcol, a_lo + 6, b_lo + 6, hi, idesc, true);
          if (last) umma_commit(empty_bar(s));
        }
        __syncwarp();
        if (last && ++s == S) { s = 0; ph ^= 1u; }
      }
      if (leader) umma_commit(tfull_bar(buf));
      __syncwarp();
    }
  } else {
    // ===================================== epilogue (warps 2-5 and 6-9) =====================================
    // Two groups of four warps: group g owns accumulator buffer g and its own exchange tile, i.e. every other tile -- the
    // TMEM -> exchange -> shifted sums -> store chain of one tile runs under that of the next (one group was latency-bound:
    // issue slots 28 % busy, tensor pipe 16 %, ncu).
    const int q = warp & 3;
    const int grp = (warp - 2) >> 2;
    float* ex = ex_all + grp * (BM * p.ex_pitch);
    const int row = q * 32 + lane;                // slab pixel held by this thread (TMEM lane)
    const bool do_stats = d.flags & MSG_CONV_STATS;
    const bool nchw = d.flags & MSG_CONV_OUT_NCHW_F32;
    float2* wsum = reinterpret_cast<float2*>(gen + (sSum - base)) + (grp * 4 + q) * 128;  // [2][64], fp64 above the 32-row partials
    float* tr = reinterpret_cast<float*>(gen + (sTr - base)) + (grp * 4 + q) * 1056;      // [32][33]
    for (int i = lane; i < 128; i += 32) wsum[i] = make_float2(0.f, 0.f);
    __syncwarp();
    int stat_img = -1;
    auto flush_stats = [&]() {
      if (stat_img >= 0) {
        for (int c = lane; c < d.n_out; c += 32) {
          double* st = p.stats + ((size_t)stat_img * d.Co_total + d.co_off + c) * 2;
          atomicAdd(st, f2sum_value(wsum[c]));
          atomicAdd(st + 1, f2sum_value(wsum[64 + c]));
          wsum[c] = make_float2(0.f, 0.f); wsum[64 + c] = make_float2(0.f, 0.f);
        }
      }
      __syncwarp();
    };
    uint32_t lt = 0;
    const int G = p.groups;
    for (int t = t_begin + (G == 2 ? grp : 0); t < t_end && grp < G; t += G, ++lt) {
      const int seg = t % p.segs, ny = t / p.segs;
      const int yrow = (ny % Hy) * R, img = ny / Hy;
      const int buf = G == 2 ? grp : (int)(lt & 1);
      const int xcol = seg * p.Wv + row;                      // output column of this thread (if row < Wv)
      const bool col_ok = row < p.Wv && xcol < d.W;
      if (do_stats && img != stat_img) { flush_stats(); stat_img = img; }
      mbar_wait(tfull_bar(buf), G == 2 ? (lt & 1) : ((lt >> 1) & 1));
      tc_fence_after();
      const uint32_t tacc = tmem_base + (uint32_t)(buf * d.Ntot) + ((uint32_t)(q * 32) << 16);
      // ---- phase 1: every accumulator column of this thread's slab pixel -> exchange tile (float4 stores;
      //      pitch = Ntot + 4 floats keeps row-per-lane 128-bit accesses conflict-free)
      float* exrow = ex + row * p.ex_pitch;
      for (int c0 = 0; c0 < d.Ntot; c0 += 32) {
        __syncwarp();
        float v[32];
        tmem_ld32(tacc + (uint32_t)c0, v);
        tmem_ld_wait();
        if (c0 + 32 >= d.Ntot) {                              // accumulator fully read: hand it back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(buf));
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (c0 + 4 * i < d.Ntot)
            *reinterpret_cast<float4*>(exrow + c0 + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      }
      asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
      // ---- phase 2: out[x][c] = sum_terms ex[x + shift][col + c], then bias / stats / activation / store
      for (int g = 0; g < d.n_groups; ++g) {
        const int oc = d.grp_out_cols[g];                     // <= 16
        const int oc0 = d.grp_out_col0[g];
        const int orow = yrow + d.grp_row[g];                 // output row of this group (multi-row tiles)
        const bool valid = col_ok && orow < d.H;
        const size_t opix = ((size_t)img * d.H + (orow < d.H ? orow : 0)) * d.W + (col_ok ? xcol : 0);
        float o[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) o[c] = 0.f;
        if (row < p.Wv) {
          for (int tm = d.grp_term_begin[g]; tm < d.grp_term_begin[g + 1]; ++tm) {
            const float* src = ex + (row + d.term_shift[tm]) * p.ex_pitch + d.grp_col0[g] + d.term_col[tm];
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
              if (c4 * 4 < oc) {
                const float4 f = *reinterpret_cast<const float4*>(src + 4 * c4);
                o[4 * c4] += f.x; o[4 * c4 + 1] += f.y; o[4 * c4 + 2] += f.z; o[4 * c4 + 3] += f.w;
              }
            }
          }
        }
        if (p.bias != nullptr) {
#pragma unroll
          for (int c = 0; c < 16; ++c)
            if (c < oc) o[c] += sbias[oc0 + c];
        }
        if (do_stats) {
#pragma unroll
          for (int c = 0; c < 16; ++c) tr[c * 33 + lane] = (valid && c < oc) ? o[c] : 0.f;
          __syncwarp();
          if (lane < 16) {
            float cs = 0.f, css = 0.f;
#pragma unroll
            for (int r = 0; r < 32; ++r) { const float xv = tr[lane * 33 + r]; cs += xv; css = fmaf(xv, xv, css); }
            if (lane < oc) { f2sum_add(wsum[oc0 + lane], cs); f2sum_add(wsum[64 + oc0 + lane], css); }
          }
          __syncwarp();
        }
        if (d.act == MSG_ACT_RELU) {
#pragma unroll
          for (int c = 0; c < 16; ++c) o[c] = fmaxf(o[c], 0.f);
        } else if (d.act == MSG_ACT_LRELU) {
#pragma unroll
          for (int c = 0; c < 16; ++c) o[c] = o[c] > 0.f ? o[c] : 0.2f * o[c];
        } else if (d.act == MSG_ACT_TANH) {
#pragma unroll
          for (int c = 0; c < 16; ++c)
            if (c < oc) asm("tanh.approx.f32 %0, %1;" : "=f"(o[c]) : "f"(o[c]));   // MUFU.TANH: 2^-11 rel. error, far below
                                                                                    // the bf16 operands' noise; tanhf was ~25
                                                                                    // instructions x 3 on an epilogue-bound kernel
        }
        if (valid) {
          if (nchw) {
            float* y = reinterpret_cast<float*>(p.y);
            const size_t plane = (size_t)d.H * d.W;
            const size_t pp = (size_t)orow * d.W + xcol;
#pragma unroll
            for (int c = 0; c < 16; ++c)
              if (c < oc) y[((size_t)img * d.Co_total + d.co_off + oc0 + c) * plane + pp] = o[c];
          } else {
            __nv_bfloat16* y = reinterpret_cast<__nv_bfloat16*>(p.y) + opix * d.Co_total + d.co_off + oc0;
            if (oc == 16 && ((d.Co_total | d.co_off | oc0) & 7) == 0) {
              float lo[8], hi[8];
#pragma unroll
              for (int c = 0; c < 8; ++c) { lo[c] = o[c]; hi[c] = o[8 + c]; }
              *reinterpret_cast<uint4*>(y) = pack8(lo);
              *reinterpret_cast<uint4*>(y + 8) = pack8(hi);
            } else {
#pragma unroll
              for (int c = 0; c < 16; ++c)
                if (c < oc) y[c] = __float2bfloat16_rn(o[c]);
            }
          }
        }
      }
      asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");          // exchange tile free for the next tile
    }
    if (do_stats && grp < G) flush_stats();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}


}  // namespace
}  // namespace msg

using namespace msg;

extern "C" int msg_conv_shift(const msg_shift_desc* d, const void* x, const void* w_rows, const float* bias,
                              void* y, double* stats, void* stream) {
  MSG_REQUIRE(d != nullptr, MSG_ERR_SHAPE, "conv_shift: null descriptor");
  MSG_REQUIRE(d->dtype == MSG_BF16, MSG_ERR_UNSUPPORTED, "conv_shift: bf16 only");
  MSG_REQUIRE(d->N > 0 && d->H > 0 && d->W > 0, MSG_ERR_SHAPE, "conv_shift: bad plane");
  MSG_REQUIRE(d->Cin % 64 == 0 && (d->Ci_total & 7) == 0 && (d->ci_off & 7) == 0, MSG_ERR_SHAPE, "conv_shift: channel layout unsupported");
  MSG_REQUIRE(d->Ntot % 16 == 0 && d->Ntot >= 16 && d->Ntot <= 256 && d->n_out >= 1 && d->n_out <= 64, MSG_ERR_SHAPE, "conv_shift: bad N configuration");
  MSG_REQUIRE(d->halo >= 0 && d->halo <= 16 && d->n_kblocks >= 1 && d->n_kblocks <= MSG_SHIFT_MAX_KBLOCKS &&
                  d->n_groups >= 1 && d->n_groups <= MSG_SHIFT_MAX_GROUPS && d->n_terms <= MSG_SHIFT_MAX_TERMS,
              MSG_ERR_SHAPE, "conv_shift: program too large");
  MSG_REQUIRE((((uintptr_t)x | (uintptr_t)w_rows) & 15) == 0, MSG_ERR_ALIGN, "conv_shift: operands must be 16-byte aligned");
  MSG_REQUIRE(!(d->flags & MSG_CONV_STATS) || stats != nullptr, MSG_ERR_SHAPE, "conv_shift: stats buffer missing");
  int rows = 0;
  for (int kb = 0; kb < d->n_kblocks; ++kb) {
    MSG_REQUIRE(d->kb_ncols[kb] >= 16 && d->kb_ncols[kb] % 16 == 0 && d->kb_col0[kb] >= 0 &&
                    d->kb_col0[kb] + d->kb_ncols[kb] <= d->Ntot && d->kb_wrow[kb] % 8 == 0,
                MSG_ERR_SHAPE, "conv_shift: bad k-block %d", kb);
    if (d->kb_wrow[kb] + d->kb_ncols[kb] > rows) rows = d->kb_wrow[kb] + d->kb_ncols[kb];
  }
  for (int g = 0; g < d->n_groups; ++g)
    MSG_REQUIRE(d->grp_span[g] >= 1 && d->grp_span[g] <= 64 && d->grp_out_cols[g] >= 1 && d->grp_out_cols[g] <= 16 &&
                    d->grp_col0[g] >= 0 && d->grp_col0[g] + d->grp_span[g] <= d->Ntot,
                MSG_ERR_SHAPE, "conv_shift: bad group %d", g);
  EncodeTiledFn enc = get_encode();
  MSG_REQUIRE(enc != nullptr, MSG_ERR_CUDA, "conv_shift: cuTensorMapEncodeTiled unavailable");

  ShiftParams p;
  p.d = *d; p.bias = bias; p.y = y; p.stats = stats;
  p.b_rows = (rows + 31) / 32 * 32;
  p.b_bytes = p.b_rows * 128;
  MSG_REQUIRE(p.b_bytes <= 100 * 1024, MSG_ERR_UNSUPPORTED, "conv_shift: %d bytes of weights do not fit in shared memory", p.b_bytes);
  p.Wv = BM - 2 * d->halo;
  p.segs = (d->W + p.Wv - 1) / p.Wv;
  p.tmem_cols = 32;
  while (p.tmem_cols < 2 * ((d->Ntot + 31) / 32 * 32)) p.tmem_cols <<= 1;
  MSG_REQUIRE(p.tmem_cols <= 512, MSG_ERR_SHAPE, "conv_shift: accumulator too wide");
  for (int g = 0; g < d->n_groups; ++g)   // the epilogue reads whole 32-column chunks: keep them inside the allocation
    MSG_REQUIRE(d->Ntot + d->grp_col0[g] + ((d->grp_span[g] + 31) / 32) * 32 <= p.tmem_cols, MSG_ERR_SHAPE,
                "conv_shift: group %d reads past the TMEM allocation", g);
  p.ex_pitch = ((d->Ntot + 3) / 4) * 4 + 4;      // 16-byte aligned rows, pitch % 32 floats == 4: conflict-free float4 access
  if (p.ex_pitch % 32 != 4) p.ex_pitch += (4 - p.ex_pitch % 32 + 32) % 32;
  // two epilogue groups (each with its own exchange tile) when that still leaves a slab ring of >= 4 stages
  auto fixed_for = [&](int groups) { return p.b_bytes + groups * BM * p.ex_pitch * 4 + 8 + 8 * 128 * 8 + 8 * 1056 * 4 + 1024 + 256 + 1024; };
  p.groups = (220 * 1024 - fixed_for(2)) / A_BYTES >= 4 ? 2 : 1;
  const int fixed = fixed_for(p.groups);
  int stages = (220 * 1024 - fixed) / A_BYTES;
  if (stages > 8) stages = 8;
  MSG_REQUIRE(stages >= 2, MSG_ERR_UNSUPPORTED, "conv_shift: not enough shared memory for the slab ring");
  p.stages = stages;
  const size_t smem = (size_t)stages * A_BYTES + fixed;

  CUtensorMap mapA, mapB;
  {
    cuuint64_t dims[4] = {(cuuint64_t)d->Cin, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->N};
    cuuint64_t strides[3] = {(cuuint64_t)d->Ci_total * 2, (cuuint64_t)d->W * d->Ci_total * 2,
                             (cuuint64_t)d->H * d->W * d->Ci_total * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)BM, 1, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    void* base = (void*)((const __nv_bfloat16*)x + d->ci_off);
    CUresult r = enc(&mapA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MSG_REQUIRE(r == CUDA_SUCCESS, MSG_ERR_CUDA, "conv_shift: cuTensorMapEncodeTiled(A) failed with %d", (int)r);
  }
  {
    cuuint64_t dims[2] = {64, (cuuint64_t)rows};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {64, 32};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&mapB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)w_rows, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MSG_REQUIRE(r == CUDA_SUCCESS, MSG_ERR_CUDA, "conv_shift: cuTensorMapEncodeTiled(B) failed with %d", (int)r);
  }
  static DeviceOnce attr_set;     // cudaFuncSetAttribute is per device
  if (attr_set.needed()) {
    cudaError_t e = cudaFuncSetAttribute(conv_shift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    MSG_REQUIRE(e == cudaSuccess, MSG_ERR_CUDA, "conv_shift: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set.done();
  }
  int grid = sm_count();
  const long long total = (long long)d->N * d->H * p.segs;
  MSG_REQUIRE(total < 0x7fffffffLL, MSG_ERR_SHAPE, "conv_shift: too many tiles");
  if (grid > total) grid = (int)total;
  conv_shift_kernel<<<grid, NTHREADS, smem, as_stream(stream)>>>(mapA, mapB, p);
  return check_launch("conv_shift_kernel");
}
