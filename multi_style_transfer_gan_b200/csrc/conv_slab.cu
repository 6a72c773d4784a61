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
//
// What bounds these kernels was measured (profiles/r1_mma_rate_microbench.md, r1_conv_slab_msb_ncu.md) and shaped the
// variants below:
//   * an M=128, K=16 tcgen05.mma costs ~39-45 cycles however small N is (its 4 KB A operand is re-read from shared memory),
//     and one issuing thread must stay under ~4 instructions per MMA: the programs known at compile time (MultiScaleBlock
//     at C = 64 / 128, 7x7 input conv) are template instantiations whose issue loop is straight-line code with immediate
//     operands and {32-bit address word, constant high word} descriptors; the generic loop serves everything else
//     (optionally split between two issuing warps by accumulator column slice);
//   * per-CTA-distinct TMA traffic saturates at ~24 B/clk/SM: programs with streamed weights take TWO output rows per tile
//     (template parameter TR), so a weight fetch serves 256 pixels; the four sub-pixel phases of a 4x4 stride-2 transposed
//     conv run as 2x2 programs with a strided output (out_stride = 2), each input row slab serving both horizontal taps;
//   * a second epilogue group (every other tile) where the epilogue's instruction stream is the bottleneck (7x7 input conv,
//     transposed-conv phases); bias adds and column statistics use packed f32x2 arithmetic.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "tcgen05.cuh"

namespace msg {
namespace {
using namespace tc;

constexpr int BM = 128;
constexpr int NTHREADS_MAX = 352;       // producer + MMA issuer A + up to two epilogue groups of four warps + MMA issuer B
constexpr int STAGE_PITCH = 128 + 16;
constexpr int EPI_STAGE = 4 * 32 * STAGE_PITCH;          // per epilogue group: four per-warp skewed staging tiles
constexpr int EPI_TR = 4 * 1088 * 4;                     // per epilogue group: transpose scratch [4][32][34] floats
                                                         // (+ fp64 column sums [4][2][Ntot] = 64 * Ntot bytes: SlabParams::epi_red)


struct SlabParams {
  msg_slab_desc d;
  const float* bias;
  void* y;
  double* stats;
  int stages, tmem_cols;
  int Ws;             // slab width in pixels (128 + 2*halo, padded to a multiple of 8)
  int a_bytes;        // bytes of one A slab (chunks * Ws * 16)
  int b_bytes;        // bytes of one B stage (0 when the weights are resident)
  int b_resident;     // all weight tiles stay in smem for the whole kernel (loaded once): total bytes, or 0
  int b_box_taps;     // streaming mode: weight tiles fetched per TMA box
  int chunks;         // K chunks per slab (8, or 1 in pixel-pair mode)
  int segs;           // 128-pixel segments per image row
  int a_rows;         // pixel-pair mode: input rows per slab (consecutive filter rows fused into ONE K block)
  int b_tiles;        // weight tiles in w_slab (= taps, or filter rows in pixel-pair mode)
  int epi_red;        // bytes of one epilogue group's reduction scratch: fp64 sums [4][2][Ntot] + transposes
  int trows;          // output rows per tile (1, or 2 for the streamed-weight MSB program: the weight tiles of a K block then
                      // serve 256 pixels, halving the weight bytes per pixel -- that kernel is bound by L2 -> SM bytes)
  int interleave;     // tile schedule: 0 = contiguous range per CTA, 1 = round-robin over the grid
  int epi_groups;     // 1: warps 2-5 drain both accumulator buffers; 2: warps 2-5 own buffer 0, warps 6-9 buffer 1
  int issuers;        // 1: warp 1 issues every MMA; 2: the taps are split between warp 1 and the last warp by ACCUMULATOR COLUMN
                      //    SLICE (tap_owner), so the two instruction streams never touch the same TMEM columns and need no
                      //    ordering between them -- the issue loop is ONE thread's serial instruction stream and, for the
                      //    small-N programs, slower than the tensor pipe (ncu: issuer 85 % busy, pipe 12-18 %)
  unsigned char tap_owner[MSG_SLAB_MAX_TAPS];   // issuer of each tap
  int a_mode;         // 0: chunk planes [chunk][pixel][16 B], no swizzle (8 TMA boxes per slab)
                      // 1: pixel rows [pixel][128 B], 128B swizzle, ONE TMA box; tap shift = +128 B/pixel on the
                      //    descriptor start, swizzle phase carried by the descriptor's base_offset field
                      // 2: as 1 with base_offset = 0
};

// ---- compile-time description of the fused MultiScaleBlock forward program (slab.py: msb_program) ----------------
// K blocks: dy in (0,-4,-2,-1,1,2,4) x channel block.  The centre row leads with its four UN-SHIFTED taps (b1 and the centres of
// b2 / b3 / b4), then (b2: -1, +1), (b3: -2, +2), (b4: -4, +4); the other rows hold the three horizontal taps of the 3x3 branch
// whose dilation is |dy|.  Knowing this at compile time turns the issue loop into straight-line code with immediate operands
// (~5 instructions per MMA instead of ~14): the generic loop is ONE thread's serial instruction stream and was slower than
// the tensor pipe (ncu: issuer 85 % busy).  The four un-shifted taps read the same slab view and write the four adjacent
// accumulator slices: ONE MMA of N = C instead of four of N = C/4 (an MMA costs ~40 cycles whatever N < 128), and as the
// first touch of every slice its first K step clears the accumulators.
struct MsbTap { int sx, branch; };
__host__ __device__ constexpr int msb_dy(int dyi) { return dyi == 0 ? 0 : dyi == 1 ? -4 : dyi == 2 ? -2 : dyi == 3 ? -1 : dyi == 4 ? 1 : dyi == 5 ? 2 : 4; }
__host__ __device__ constexpr int msb_ntaps(int dyi) { return dyi == 0 ? 10 : 3; }
__host__ __device__ constexpr MsbTap msb_tap(int dyi, int j) {
  if (dyi == 0) {
    if (j < 4) return MsbTap{0, j};
    const int b = 1 + (j - 4) / 2, d = b == 1 ? 1 : b == 2 ? 2 : 4;
    return MsbTap{((j - 4) % 2) ? d : -d, b};
  }
  const int dy = msb_dy(dyi), ad = dy < 0 ? -dy : dy;
  const int b = ad == 1 ? 1 : ad == 2 ? 2 : 3;
  return MsbTap{(j - 1) * ad, b};
}
__host__ __device__ constexpr bool msb_first(int dyi, int cb, int j) { return dyi == 0 && cb == 0 && j < 4; }
// taps per K block (slab.py: msb_max_taps).  Measured on B200 for C = 128 (streamed weights): splitting the 10-tap centre row
// into 4 + 4 + 2 (4 pipeline stages of 33 KB instead of 2 of 57 KB, but two more slab loads per tile) is SLOWER, 0.585 ->
// 0.626 ms per 16 images: the kernel moves ~24 B/clk/SM from L2 either way, i.e. it is bound by L2 -> SM bytes, not latency.
__host__ __device__ constexpr int msb_maxt(int) { return 32; }   // = no split for C = 64 and C = 128
__host__ __device__ constexpr int msb_taps_before(int dyi) {   // taps of one channel block in the K blocks before row dyi
  int n = 0;
  for (int i = 0; i < dyi; ++i) n += msb_ntaps(i);
  return n;
}

template <int MSBC, int TR>   // TR: output rows per tile (compile-time for the straight-line issue code)
                         // MSBC 0: generic program from the descriptor; 64 / 128: the MultiScaleBlock forward program of that width;
                         // 7: the 7x7 input conv on the 8-channel image (pixel-pair K, all seven rows in one K block)
__global__ void __launch_bounds__(NTHREADS_MAX, 1)
conv_slab_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                 const SlabParams p) {
  extern __shared__ uint8_t smem_raw[];
  const msg_slab_desc& d = p.d;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = p.stages;
  const int G = p.epi_groups;
  const int mma_b_warp = 2 + 4 * G;                          // second MMA issuer (issuers == 2)
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sB = base;                                   // weight tiles first: 1024-byte aligned
  const uint32_t sA = sB + (p.b_resident ? p.b_resident : S * p.b_bytes);
  const uint32_t sStage = (sA + S * p.a_bytes + 127u) & ~127u;
  const uint32_t sRed = sStage + G * EPI_STAGE;               // per group: per-warp column sums [4][2][256] + transpose scratch
  const uint32_t sBias = sRed + G * p.epi_red;                  // bias staged once per CTA (Ntot <= 256 floats)
  const uint32_t sBar = sBias + 1024;
  uint8_t* stage_gen = gen + (sStage - base);
  float* red = reinterpret_cast<float*>(gen + (sRed - base));
  float* sbias = reinterpret_cast<float*>(gen + (sBias - base));
  const bool bias_in_smem = p.bias != nullptr;
  if (bias_in_smem)
    for (int i = tid; i < d.Ntot; i += (int)blockDim.x) sbias[i] = p.bias[i];
  const uint32_t wres_bar = sBar + 8u * (2 * S + 4);          // "resident weights have landed"
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + (sBar - base) + 8 * (2 * S + 5));
  auto full_bar = [&](int s) { return sBar + 8u * s; };
  auto empty_bar = [&](int s) { return sBar + 8u * (S + s); };
  auto tfull_bar = [&](int b) { return sBar + 8u * (2 * S + b); };
  auto tempty_bar = [&](int b) { return sBar + 8u * (2 * S + 2 + b); };

  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), (uint32_t)p.issuers); }
      for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), (uint32_t)p.issuers); mbar_init(tempty_bar(b), 4); }
      mbar_init(wres_bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapB)) : "memory");
  } else if (warp == 2 && p.a_mode != 0) {
    // swizzled-slab mode: tile-invariant tap operands {A offset, B offset, TMEM column, overwrite} in 16-byte
    // descriptor units, one 16-byte smem entry per tap, so the elected lane's issue loop is one LDS.128 + three
    // adds per tap instead of a chain of indexed constant loads
    uint4* tt = reinterpret_cast<uint4*>(gen + (sBar + 256u - base));
    for (int tp = lane; tp < d.n_taps; tp += 32)
      tt[tp] = make_uint4((uint32_t)d.tap_sx[tp] >> 4, (uint32_t)d.tap_kstep[tp] >> 4, (uint32_t)d.tap_acc_col[tp],
                          (uint32_t)d.tap_first[tp] | ((uint32_t)p.tap_owner[tp] << 1));
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int Hy = d.H / TR;                        // tile rows per image
  const int total_tiles = d.N * Hy * p.segs;
  // tile schedule: CTA-contiguous ranges (tstep = 1), or round-robin (tstep = grid): then all CTAs work on neighbouring
  // rows of the same image at any time, so every input row comes from DRAM once and its 7 re-reads (one per dy) hit L2
  const int tstep = p.interleave ? (int)gridDim.x : 1;
  const int t_begin = p.interleave ? (int)blockIdx.x : (int)((long long)blockIdx.x * total_tiles / gridDim.x);
  const int t_end = p.interleave ? total_tiles : (int)((long long)(blockIdx.x + 1) * total_tiles / gridDim.x);
  const int tap_bytes = d.ncols * 128;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      if (p.b_resident) {       // all weight tiles, once: they stay in smem for every tile of this CTA
        const int tiles_total = p.b_tiles;
        mbar_expect_tx(wres_bar, (uint32_t)(((tiles_total + p.b_box_taps - 1) / p.b_box_taps) * p.b_box_taps * tap_bytes));
        for (int bx = 0; bx * p.b_box_taps < tiles_total; ++bx)
          tma_load_2d(sB + bx * p.b_box_taps * tap_bytes, &mapB, wres_bar, 0, bx * p.b_box_taps * d.ncols);
      }
      int s = 0;
      uint32_t ph = 0;
      bool wrapped = false;
      for (int t = t_begin; t < t_end; t += tstep) {
        const int seg = t % p.segs, ny = t / p.segs;
        const int yrow = (ny % Hy) * TR, img = ny / Hy;
        const int x0 = seg * BM - d.halo;
        for (int kb = 0; kb < d.n_kblocks; ++kb) {
          if (wrapped) mbar_wait(empty_bar(s), ph ^ 1u);
          const int t0 = d.kb_tap_begin[kb], t1 = d.kb_tap_begin[kb + 1];
          const int nbox = p.b_resident ? 0 : (d.pixel_pair_k ? 1 : (t1 - t0 + p.b_box_taps - 1) / p.b_box_taps);
          mbar_expect_tx(full_bar(s), p.a_bytes + nbox * p.b_box_taps * tap_bytes);
          if (p.a_mode != 0) {
            // slab = [Ws pixels][64 channels], 128B-swizzled: one box {64 ch, Ws, 1, 1}
            tma_load_4d(sA + s * p.a_bytes, &mapA, full_bar(s), d.kb_cb[kb] * 64, x0, yrow + d.kb_dy[kb], img);
          } else {
            // slab = `chunks` planes of [Ws pixels][16 B]: one box {8 ch, Ws, 1, 1} per 8-channel chunk
            for (int ch = 0; ch < p.chunks; ++ch)
              tma_load_4d(sA + s * p.a_bytes + ch * (p.Ws * 16), &mapA, full_bar(s), d.kb_cb[kb] * 64 + ch * 8, x0,
                          yrow + d.kb_dy[kb], img);
          }
          // streamed weights: boxes of b_box_taps tiles (small TMA requests are latency-bound, so few big ones)
          const int row0 = (d.pixel_pair_k ? kb : t0) * d.ncols;
          for (int bx = 0; bx < nbox; ++bx)
            tma_load_2d(sB + s * p.b_bytes + bx * p.b_box_taps * tap_bytes, &mapB, full_bar(s), 0,
                        row0 + bx * p.b_box_taps * d.ncols);
          if (++s == S) { s = 0; ph ^= 1u; wrapped = true; }   // no div/mod by the runtime stage count
        }
      }
    }
  } else if (warp == 1 || (p.issuers == 2 && warp == mma_b_warp)) {
    // ===================================== MMA issuer(s) =====================================
    const int issuer = warp == 1 ? 0 : 1;
    // The WHOLE warp runs this loop (warp-uniform control flow, per-tap constants read from the kernel
    // parameters = constant bank) so that descriptors stay in uniform registers; only the tcgen05
    // instructions themselves are issued by one elected lane.
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(d.ncols >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    const uint32_t lbo = (uint32_t)(p.Ws * 16);
    const uint64_t sw128_hi = make_sw128_desc(0);             // descriptor with a zero start address
    const int nch = d.n_chains;
    const uint32_t ch1 = nch > 1 ? (uint32_t)d.Ntot : 0u;
    const uint32_t ch2 = nch > 2 ? (uint32_t)(2 * d.Ntot) : 0u;
    const uint32_t ch3 = nch > 2 ? (uint32_t)(3 * d.Ntot) : ch1;
    // pixel-pair mode: lane l <-> tap l (n_taps <= 32, checked on the host); all in 16-byte descriptor units
    const uint64_t noswz_hi = make_noswz_desc(0, 16u);
    uint32_t pp_a = 0, pp_b = 0, pp_col = 0, pp_acc = 1;
    if (d.pixel_pair_k && lane < d.n_taps) {
      pp_a = (uint32_t)d.tap_sx[lane] >> 4;
      pp_b = (uint32_t)d.tap_kstep[lane] >> 4;
      const int chain = nch > 1 ? lane % nch : 0;           // K block 0 starts at tap 0
      pp_col = (uint32_t)d.tap_acc_col[lane] + (uint32_t)(chain * d.Ntot);
      pp_acc = !(lane < nch);                                // the first tap of each chain overwrites
    }
    // swizzled-slab mode: tile-invariant tap operands {A offset, B offset, TMEM column, overwrite} in 16-byte
    // descriptor units, one 16-byte smem entry per tap, so the elected lane's issue loop is one LDS.128 + three
    // adds per tap instead of a chain of indexed constant loads
    const uint4* tap_tab = reinterpret_cast<const uint4*>(gen + (sBar + 256u - base));   // built before the block barrier
    int s = 0;
    uint32_t ph = 0, lt = 0;
    if (p.b_resident) mbar_wait(wres_bar, 0);
    if constexpr (MSBC == 7) {
      // ---- 7x7 input conv: 7 rows x 4 pixel-pair taps, one K block per tile, 28 MMAs of straight-line code --------
      // (the generic loop broadcast per-tap operands with shuffles: ~195 cycles per MMA against 48 on the pipe)
      const uint32_t a_hi = (uint32_t)(noswz_hi >> 32), a_lo0 = (uint32_t)noswz_hi;      // LBO lives in the low word
      const uint32_t b_hi = (uint32_t)(sw128_hi >> 32);
      const bool leader = elect_one();
      const uint32_t a_base = (sA >> 4) + a_lo0, a_step = (uint32_t)p.a_bytes >> 4;
      const uint32_t row_u = (uint32_t)(p.a_bytes / p.a_rows) >> 4;      // one input row of the slab, 16-byte units
      const uint32_t tile_u = (uint32_t)tap_bytes >> 4;                 // one filter row's weight tile
      const uint32_t b_base = sB >> 4;
      for (int t = t_begin; t < t_end; t += tstep, ++lt) {
        const int buf = lt & 1;
        if (lt >= 2) mbar_wait(tempty_bar(buf), ((lt >> 1) - 1) & 1);
        tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(buf * d.Ntot);
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint32_t a0 = a_base + (uint32_t)s * a_step;
        if (leader) {
#pragma unroll
          for (int kh = 0; kh < 7; ++kh)
#pragma unroll
            for (int j = 0; j < 4; ++j)
              umma_bf16_lo2(tacc, a0 + (uint32_t)kh * row_u + (uint32_t)(2 * j), a_hi,
                            b_base + (uint32_t)kh * tile_u + (uint32_t)(2 * j), b_hi, idesc, !(kh == 0 && j == 0));
          umma_commit(empty_bar(s));
          umma_commit(tfull_bar(buf));
        }
        __syncwarp();
        if (++s == S) { s = 0; ph ^= 1u; }
      }
    } else if constexpr (MSBC != 0) {
      // ---- specialised straight-line issue: every tap operand is an immediate --------------------------------
      constexpr int CB = MSBC / 64, Q = MSBC / 4, TAPU = Q * 8;      // weight tile of one tap in 16-byte units
      constexpr int MAXT = msb_maxt(MSBC);
      static_assert(MAXT >= 4, "the merged centre taps must sit in one K block");
      const uint32_t idesc_c = (idesc & ~(0x3fu << 17)) | ((uint32_t)(MSBC >> 3) << 17);      // N = MSBC
      const uint32_t hi = (uint32_t)(sw128_hi >> 32);
      const bool leader = elect_one();
      const uint32_t a_base = sA >> 4, a_step = (uint32_t)p.a_bytes >> 4;
      const uint32_t b_base = sB >> 4, b_step = (uint32_t)p.b_bytes >> 4;
      const bool bres = p.b_resident != 0;
      for (int t = t_begin; t < t_end; t += tstep, ++lt) {
        const int buf = lt & 1;
        if (lt >= 2) mbar_wait(tempty_bar(buf), ((lt >> 1) - 1) & 1);
        tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(buf * MSBC * TR);
        const uint32_t row_u = (uint32_t)(p.Ws * 128) >> 4;                  // second output row: next slab row, next MSBC columns
#pragma unroll
        for (int dyi = 0; dyi < 7; ++dyi) {
#pragma unroll
          for (int cb = 0; cb < CB; ++cb) {
#pragma unroll
            for (int j0 = 0; j0 < msb_ntaps(dyi); j0 += MAXT) {      // K blocks of at most MAXT taps (slab.py: msb_program)
              mbar_wait(full_bar(s), ph);
              tc_fence_after();
              const int tp0 = msb_taps_before(dyi) * CB + cb * msb_ntaps(dyi) + j0;   // program-order index of this K block's first tap
              const uint32_t a0 = a_base + (uint32_t)s * a_step;
              const uint32_t b0 = bres ? b_base + (uint32_t)(tp0 * TAPU) : b_base + (uint32_t)s * b_step;
              if (leader) {
#pragma unroll
                for (int r = 0; r < TR; ++r) {
                  const uint32_t ar = a0 + (uint32_t)r * row_u, dr = tacc + (uint32_t)(r * MSBC);
#pragma unroll
                  for (int j = j0; j < (j0 + MAXT < msb_ntaps(dyi) ? j0 + MAXT : msb_ntaps(dyi)); ++j) {
                    if (dyi == 0 && j < 4) {
                      if (j == 0) {          // the four un-shifted taps: one N = MSBC MMA per K step over the four adjacent weight tiles
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)
                          umma_bf16_lo(dr, ar + (uint32_t)(4 * 8 + 2 * ks), b0 + (uint32_t)(2 * ks), hi, idesc_c, !(cb == 0 && ks == 0));
                      }
                      continue;
                    }
                    const MsbTap tap = msb_tap(dyi, j);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                      umma_bf16_lo(dr + (uint32_t)(tap.branch * Q), ar + (uint32_t)((4 + tap.sx) * 8 + 2 * ks),
                                   b0 + (uint32_t)((j - j0) * TAPU + 2 * ks), hi, idesc, true);
                  }
                }
                umma_commit(empty_bar(s));
              }
              __syncwarp();
              if (++s == S) { s = 0; ph ^= 1u; }
            }
          }
        }
        if (leader) umma_commit(tfull_bar(buf));
        __syncwarp();
      }
    } else
    for (int t = t_begin; t < t_end; t += tstep, ++lt) {
      const int buf = lt & 1;
      if (lt >= 2) mbar_wait(tempty_bar(buf), ((lt >> 1) - 1) & 1);
      tc_fence_after();
      const uint32_t tacc = tmem_base + (uint32_t)(buf * TR * d.Ntot * d.n_chains);
      for (int kb = 0; kb < d.n_kblocks; ++kb) {
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint32_t a0 = sA + s * p.a_bytes;
        const uint32_t b0 = p.b_resident ? sB : sB + s * p.b_bytes;
        const int t0 = d.kb_tap_begin[kb], t1 = d.kb_tap_begin[kb + 1];
        if (p.a_mode != 0) {
          // [pixel][128 B] slab, 128B swizzle: tap shift = +128 B per pixel on the start address; the
          // swizzle is a function of the absolute smem address, so base_offset stays 0 (verified on B200)
          const uint32_t a0s = a0 >> 4, b0s = b0 >> 4;
          if (elect_one()) {
#pragma unroll
            for (int r = 0; r < TR; ++r) {           // TR == 2: second output row = next slab row, next Ntot accumulator columns
              const uint32_t ar = a0s + (uint32_t)r * ((uint32_t)(p.Ws * 128) >> 4), dr = tacc + (uint32_t)(r * d.Ntot);
              uint4 e = tap_tab[t0];
              for (int tp = t0; tp < t1; ++tp) {
                const uint4 nx = tap_tab[tp + 1 < t1 ? tp + 1 : tp];      // prefetch the next entry
                if ((int)(e.w >> 1) != issuer) { e = nx; continue; }     // the other issuer's accumulator slice
                e.w &= 1u;
                const uint64_t da = sw128_hi | (uint64_t)(ar + e.x);
                const uint64_t db = sw128_hi | (uint64_t)(b0s + e.y);
                const uint32_t dcol = dr + e.z;
                if (nch == 1) {
                  umma_bf16(dcol, da, db, idesc, !e.w);
                  umma_bf16(dcol, da + 2, db + 2, idesc, 1);
                  umma_bf16(dcol, da + 4, db + 4, idesc, 1);
                  umma_bf16(dcol, da + 6, db + 6, idesc, 1);
                } else {
                  umma_bf16(dcol, da, db, idesc, !e.w);
                  umma_bf16(dcol + ch1, da + 2, db + 2, idesc, !e.w);
                  umma_bf16(dcol + ch2, da + 4, db + 4, idesc, !(e.w && nch > 2));
                  umma_bf16(dcol + ch3, da + 6, db + 6, idesc, !(e.w && nch > 2));
                }
                e = nx;
              }
            }
          }
        } else if (d.pixel_pair_k) {
          // One MMA per tap.  A serial per-tap chain (indexed constant loads -> descriptor -> uniform registers ->
          // tcgen05.mma) cost ~370 cycles per MMA on this single warp (ncu: the issuer never waits on a barrier,
          // the tensor pipe idles 65 %), so lane l holds the tile-invariant operands of tap l (built once per
          // kernel, above) and the issue loop only broadcasts them with independent shuffles and adds the stage /
          // TMEM-buffer bases.  One elected lane issues every MMA (a tcgen05.commit only tracks the MMAs of the
          // thread that executes it).
          const uint32_t a0s = a0 >> 4, b0s = b0 >> 4;
          for (int tp = t0; tp < t1; ++tp) {
            const uint32_t ta = __shfl_sync(0xffffffffu, pp_a, tp) + a0s;
            const uint32_t tb = __shfl_sync(0xffffffffu, pp_b, tp) + b0s;
            const uint32_t dc = __shfl_sync(0xffffffffu, pp_col, tp) + tacc;
            const uint32_t ac = __shfl_sync(0xffffffffu, pp_acc, tp);
            if (elect_one()) umma_bf16(dc, noswz_hi | ta, sw128_hi | tb, idesc, ac);
          }
        } else {
          for (int tp = t0; tp < t1; ++tp) {
            const uint64_t db = sw128_hi | (uint64_t)((b0 + (uint32_t)d.tap_kstep[tp]) >> 4);
            const uint32_t dcol = tacc + (uint32_t)d.tap_acc_col[tp];
            const uint32_t first = (uint32_t)d.tap_first[tp];
            if (elect_one()) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                umma_bf16(dcol, make_noswz_desc(a0 + (uint32_t)d.tap_sx[tp] + (uint32_t)(ks * 2) * lbo, lbo), db + (uint64_t)(ks * 2),
                          idesc, !(first && ks == 0));
            }
          }
        }
        __syncwarp();
        if (elect_one()) umma_commit(empty_bar(s));
        if (++s == S) { s = 0; ph ^= 1u; }
      }
      __syncwarp();
      if (elect_one()) umma_commit(tfull_bar(buf));
    }
  } else if (warp >= 2 && warp < 2 + 4 * G) {
    // ===================================== epilogue (warps 2-5, and 6-9 with two groups) ===========
    // With two groups, group g drains accumulator buffer g, i.e. every other tile: two epilogues in flight per scheduler.
    const int grp = (warp - 2) >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const bool do_stats = d.flags & MSG_CONV_STATS;
    const bool nchw = d.flags & MSG_CONV_OUT_NCHW_F32;
    uint8_t* stage_w = stage_gen + grp * EPI_STAGE + q * (32 * STAGE_PITCH);
    const int cmax = d.n_store;                   // columns actually stored (<= Ntot)
    const int WS = d.Ntot;                        // sums in [0, WS), sums of squares in [WS, 2 WS)
    float2* wsum = reinterpret_cast<float2*>(red + grp * (p.epi_red / 4)) + q * 2 * WS;   // [2][256] running column sums of this warp (fp64 above the
                                                               // fixed-order 32-row fp32 partials: grouping-independent)
    float* tr = red + grp * (p.epi_red / 4) + 4 * 4 * WS + q * 1088;   // [32][34] transpose scratch of this warp (after the 4 warps' fp64 sums)
    for (int i = lane; i < 2 * WS; i += 32) wsum[i] = make_float2(0.f, 0.f);
    __syncwarp();
    int stat_img = -1;
    auto flush_stats = [&]() {
      if (stat_img >= 0) {
        for (int c = lane; c < cmax; c += 32) {
          double* st = p.stats + ((size_t)stat_img * d.Co_total + d.co_off + c) * 2;
          atomicAdd(st, f2sum_value(wsum[c]));
          atomicAdd(st + 1, f2sum_value(wsum[WS + c]));
          wsum[c] = make_float2(0.f, 0.f); wsum[WS + c] = make_float2(0.f, 0.f);
        }
      }
      __syncwarp();
    };
    const bool vec = !nchw && ((d.Co_total | d.co_off | cmax) & 7) == 0;
    uint32_t lt = 0;
    for (int t = t_begin + (G == 2 ? grp * tstep : 0); t < t_end; t += G * tstep, ++lt) {
      const int seg = t % p.segs, ny = t / p.segs;
      const int yrow0 = (ny % Hy) * TR, img = ny / Hy;
      const int buf = G == 2 ? grp : (int)(lt & 1);                      // accumulator buffer of this tile
      const uint32_t tf_parity = G == 2 ? (lt & 1) : ((lt >> 1) & 1);    // its (lt or lt/2)-th use
      const int xcol = seg * BM + row;
      const bool valid = xcol < d.W;
      if (do_stats && img != stat_img) { flush_stats(); stat_img = img; }
      mbar_wait(tfull_bar(buf), tf_parity);
      tc_fence_after();
      for (int rr = 0; rr < TR; ++rr) {
      const int yrow = yrow0 + rr;
      const int os = d.out_stride > 1 ? d.out_stride : 1;      // sub-pixel phase of a transposed conv: strided output
      const int opix = ((img * d.H + yrow) * os + d.out_off_h) * (d.W * os) + (valid ? xcol : 0) * os + d.out_off_w;
      const uint32_t tacc = tmem_base + (uint32_t)((buf * TR + rr) * d.Ntot * d.n_chains) + ((uint32_t)(q * 32) << 16);
      for (int cg = 0; cg < cmax; cg += 64) {
        const int ncol = (cmax - cg) < 64 ? (cmax - cg) : 64;
        __syncwarp();
        float v[64];
        tmem_ld32(tacc + (uint32_t)cg, *reinterpret_cast<float(*)[32]>(&v[0]));
        if (ncol > 32) tmem_ld32(tacc + (uint32_t)(cg + 32), *reinterpret_cast<float(*)[32]>(&v[32]));
        tmem_ld_wait();
        for (int c = 1; c < d.n_chains; ++c) {      // add the other accumulation chains
          float u[64];
          tmem_ld32(tacc + (uint32_t)(c * d.Ntot + cg), *reinterpret_cast<float(*)[32]>(&u[0]));
          if (ncol > 32) tmem_ld32(tacc + (uint32_t)(c * d.Ntot + cg + 32), *reinterpret_cast<float(*)[32]>(&u[32]));
          tmem_ld_wait();
#pragma unroll
          for (int jj = 0; jj < 64; ++jj)
            if (jj < 32 || ncol > 32) v[jj] += u[jj];
        }
        if (cg + 64 >= cmax && rr == TR - 1) {      // last read of this accumulator buffer
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(buf));
        }
        if (p.bias != nullptr) {
          if (bias_in_smem) {
            // packed f32x2 adds on 128-bit shared loads (columns >= ncol are never stored; sbias holds 256 floats)
            const float4* b4 = reinterpret_cast<const float4*>(sbias + cg);
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
              const float4 b = b4[jj];
              const float2 lo = __fadd2_rn(make_float2(v[4 * jj], v[4 * jj + 1]), make_float2(b.x, b.y));
              const float2 hi2 = __fadd2_rn(make_float2(v[4 * jj + 2], v[4 * jj + 3]), make_float2(b.z, b.w));
              v[4 * jj] = lo.x; v[4 * jj + 1] = lo.y; v[4 * jj + 2] = hi2.x; v[4 * jj + 3] = hi2.y;
            }
          } else {
#pragma unroll
            for (int jj = 0; jj < 64; ++jj)
              if (jj < ncol) v[jj] += __ldg(p.bias + cg + jj);
          }
        }
        if (do_stats) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (h * 32 < ncol) {
              // transpose this warp's 32 x 32 block through smem ([col][34]: conflict-free both ways); lane l
              // then owns column l and sums its 32 rows (x and x^2) -- ~2x fewer instructions than shuffles
#pragma unroll
              for (int jj = 0; jj < 32; ++jj) tr[jj * 34 + lane] = valid ? v[h * 32 + jj] : 0.f;
              __syncwarp();
              // packed f32x2: rows (2r, 2r+1) of this lane's column per LDS.64 (pitch 34: conflict-free both ways)
              float2 s2 = make_float2(0.f, 0.f), q2 = make_float2(0.f, 0.f);
#pragma unroll
              for (int r = 0; r < 16; ++r) {
                const float2 xv = *reinterpret_cast<const float2*>(&tr[lane * 34 + 2 * r]);
                s2 = __fadd2_rn(s2, xv);
                q2 = __ffma2_rn(xv, xv, q2);
              }
              const float cs = s2.x + s2.y, css = q2.x + q2.y;
              __syncwarp();
              if (cg + h * 32 + lane < WS) {
                f2sum_add(wsum[cg + h * 32 + lane], cs);
                f2sum_add(wsum[WS + cg + h * 32 + lane], css);
              }
            }
          }
        }
        if (d.act != MSG_ACT_NONE) {              // activation, switch hoisted out of the element loop
          if (d.act == MSG_ACT_RELU) {
#pragma unroll
            for (int jj = 0; jj < 64; ++jj) v[jj] = fmaxf(v[jj], 0.f);
          } else if (d.act == MSG_ACT_LRELU) {
#pragma unroll
            for (int jj = 0; jj < 64; ++jj) v[jj] = v[jj] > 0.f ? v[jj] : 0.2f * v[jj];
          } else {
#pragma unroll
            for (int jj = 0; jj < 64; ++jj)
              if (jj < ncol) v[jj] = tanhf(v[jj]);
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
                y[((size_t)img * d.Co_total + d.co_off + cg + jj) * plane + pp] = v[jj];
          }
        } else if (vec && (ncol == 64 || ncol == 32 || ncol == 16)) {
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            if (g * 8 < ncol) {
              float o[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) o[e] = v[g * 8 + e];
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
            if (e < ncol) y[e] = __float2bfloat16_rn(v[e]);
        }
      }
      }   // rr
    }
    if (do_stats) flush_stats();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}


// does the descriptor hold exactly the compile-time MultiScaleBlock program of width C?
template <int C>
bool is_msb_program(const msg_slab_desc* d) {
  constexpr int CB = C / 64, Q = C / 4, MAXT = msb_maxt(C);
  if (d->Cin != C || d->Ntot != C || d->ncols != Q || d->halo != 4 || d->pixel_pair_k || d->n_chains != 1 || d->n_taps != 28 * CB)
    return false;
  int tp = 0, kb = 0;
  for (int dyi = 0; dyi < 7; ++dyi)
    for (int cb = 0; cb < CB; ++cb)
      for (int j0 = 0; j0 < msb_ntaps(dyi); j0 += MAXT, ++kb) {
        if (kb >= d->n_kblocks || d->kb_dy[kb] != msb_dy(dyi) || d->kb_cb[kb] != cb || d->kb_tap_begin[kb] != tp) return false;
        for (int j = j0; j < msb_ntaps(dyi) && j < j0 + MAXT; ++j, ++tp) {
          const MsbTap t = msb_tap(dyi, j);
          if (d->tap_sx[tp] != t.sx || d->tap_acc_col[tp] != t.branch * Q) return false;
          if (d->tap_first[tp] != (msb_first(dyi, cb, j) ? 1 : 0)) return false;
        }
      }
  return kb == d->n_kblocks && d->kb_tap_begin[kb] == tp;
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
  MSG_REQUIRE((d->n_chains == 1 || d->n_chains == 2 || d->n_chains == 4) && d->n_chains * d->Ntot <= 256, MSG_ERR_SHAPE,
              "conv_slab: bad n_chains");
  MSG_REQUIRE(d->n_kblocks >= 1 && d->n_kblocks <= MSG_SLAB_MAX_KBLOCKS && d->n_taps >= 1 && d->n_taps <= MSG_SLAB_MAX_TAPS &&
                  d->halo >= 0 && d->halo <= 16,
              MSG_ERR_SHAPE, "conv_slab: program too large");
  MSG_REQUIRE(!d->pixel_pair_k || d->n_taps <= 32, MSG_ERR_SHAPE, "conv_slab: pixel-pair programs hold at most 32 taps");
  MSG_REQUIRE((((uintptr_t)x | (uintptr_t)w_slab) & 15) == 0, MSG_ERR_ALIGN, "conv_slab: operands must be 16-byte aligned");
  MSG_REQUIRE(!(d->flags & MSG_CONV_STATS) || stats != nullptr, MSG_ERR_SHAPE, "conv_slab: stats buffer missing");
  MSG_REQUIRE(d->out_stride == 0 || d->out_stride == 1 ||
                  (d->out_stride == 2 && !(d->flags & MSG_CONV_OUT_NCHW_F32) && (unsigned)d->out_off_h < 2u && (unsigned)d->out_off_w < 2u),
              MSG_ERR_SHAPE, "conv_slab: out_stride must be 1, or 2 with offsets in {0,1} and an NHWC bf16 output");
  EncodeTiledFn enc = get_encode();
  MSG_REQUIRE(enc != nullptr, MSG_ERR_CUDA, "conv_slab: cuTensorMapEncodeTiled unavailable");

  SlabParams p;
  p.d = *d; p.bias = bias; p.y = y; p.stats = stats;
  p.chunks = d->pixel_pair_k ? 1 : 8;
  p.Ws = (BM + 2 * d->halo + (d->pixel_pair_k ? 1 : 0) + 7) / 8 * 8;
  MSG_REQUIRE(p.Ws <= 256, MSG_ERR_SHAPE, "conv_slab: halo too large");
  static const int env_mode = [] { const char* e = getenv("MSG_SLAB_AMODE"); return e ? atoi(e) : 2; }();
  p.a_mode = d->pixel_pair_k ? 0 : env_mode;
  MSG_REQUIRE(p.a_mode != 0 || d->pixel_pair_k || d->n_chains == 1, MSG_ERR_UNSUPPORTED, "conv_slab: chunk-plane mode is single-chain");
  p.a_bytes = p.a_mode != 0 ? p.Ws * 128 : p.chunks * p.Ws * 16;
  int max_taps = 0;
  for (int kb = 0; kb < d->n_kblocks; ++kb) {
    int nt = d->kb_tap_begin[kb + 1] - d->kb_tap_begin[kb];
    MSG_REQUIRE(nt >= 1, MSG_ERR_SHAPE, "conv_slab: empty k-block");
    if (nt > max_taps) max_taps = nt;
  }
  const int tap_bytes = d->ncols * 128;
  int min_taps = max_taps;
  for (int kb = 0; kb < d->n_kblocks; ++kb) {
    int nt = d->kb_tap_begin[kb + 1] - d->kb_tap_begin[kb];
    if (nt < min_taps) min_taps = nt;
  }
  const int n_tiles_total = d->pixel_pair_k ? d->n_kblocks : d->n_taps;
  p.b_box_taps = d->pixel_pair_k ? 1 : min_taps;
  while (p.b_box_taps * d->ncols > 256) --p.b_box_taps;                       // TMA box rows <= 256
  static const bool env_res = [] { const char* e = getenv("MSG_SLAB_RESIDENT"); return !(e && e[0] == '0'); }();
  const int res_bytes = ((n_tiles_total + p.b_box_taps - 1) / p.b_box_taps) * p.b_box_taps * tap_bytes;
  p.b_resident = (env_res && res_bytes <= 104 * 1024) ? res_bytes : 0;
  p.b_bytes = p.b_resident ? 0 : (d->pixel_pair_k ? 1 : (max_taps + p.b_box_taps - 1) / p.b_box_taps * p.b_box_taps) * tap_bytes;
  p.segs = (d->W + BM - 1) / BM;
  // per-tap constants of the MMA issue loop, precomputed here and read from the constant bank:
  //   tap_sx    <- byte offset of the tap's shifted view inside the A slab
  //   tap_kstep <- byte offset of the tap's weight tile (or 16-wide K slice) inside the B stage / resident block
  for (int kb = 0; kb < d->n_kblocks; ++kb)
    for (int tp = d->kb_tap_begin[kb]; tp < d->kb_tap_begin[kb + 1]; ++tp) {
      p.d.tap_sx[tp] = (d->halo + d->tap_sx[tp]) * (p.a_mode != 0 ? 128 : 16);
      if (d->pixel_pair_k) p.d.tap_kstep[tp] = d->tap_kstep[tp] * 32 + (p.b_resident ? kb * tap_bytes : 0);
      else p.d.tap_kstep[tp] = (p.b_resident ? tp : tp - d->kb_tap_begin[kb]) * tap_bytes;
    }
  p.a_rows = 1;
  p.b_tiles = n_tiles_total;
  // straight-line issue code for the programs known at compile time (MSG_SLAB_SPECIALISE=0: generic loop)
  static const bool env_spec = [] { const char* e = getenv("MSG_SLAB_SPECIALISE"); return !(e && e[0] == '0'); }();
  int spec = 0;
  if (d->pixel_pair_k && p.b_resident && d->n_kblocks > 1) {
    // Consecutive filter rows -> ONE K block: a single TMA box {8 ch, Ws, rows} brings every input row of the tile
    // (out-of-image rows zero-filled), the issuer runs all taps back to back and commits once, instead of one
    // barrier round trip per filter row (the issuing warp's serial instruction stream is the bottleneck here).
    bool consecutive = true;
    for (int kb = 1; kb < d->n_kblocks; ++kb)
      consecutive = consecutive && d->kb_dy[kb] == d->kb_dy[0] + kb && d->kb_cb[kb] == d->kb_cb[0];
    if (consecutive && d->n_kblocks * p.a_bytes <= 32 * 1024) {
      for (int kb = 0; kb < d->n_kblocks; ++kb)
        for (int tp = d->kb_tap_begin[kb]; tp < d->kb_tap_begin[kb + 1]; ++tp) p.d.tap_sx[tp] += kb * p.a_bytes;
      p.a_rows = d->n_kblocks;
      p.a_bytes *= d->n_kblocks;
      p.d.n_kblocks = 1;
      p.d.kb_tap_begin[1] = d->n_taps;
    }
  }
  if (env_spec && d->pixel_pair_k && p.a_rows == 7 && p.b_resident && d->n_taps == 28 && d->n_chains == 1) {
    bool ok = true;                                   // conv7_in_program: row kh, taps sx = 2j - 3 on K slice j
    for (int tp = 0; tp < 28; ++tp)
      ok = ok && d->tap_sx[tp] == 2 * (tp & 3) - 3 && d->tap_kstep[tp] == (tp & 3) && d->tap_acc_col[tp] == 0 &&
           d->kb_tap_begin[tp >> 2] == (tp & ~3) && d->kb_dy[tp >> 2] == (tp >> 2) - 3;
    if (ok && d->halo == 3) spec = 7;
  }
  // straight-line issue code for the MultiScaleBlock forward programs (C = 64, 128)
  if (spec == 0 && env_spec && p.a_mode != 0) spec = is_msb_program<64>(d) ? 64 : is_msb_program<128>(d) ? 128 : 0;
  p.epi_red = 64 * d->Ntot + EPI_TR;
  auto fixed_for = [&](int groups) {
    return 128 + groups * (EPI_STAGE + p.epi_red) + 1024 + 256 + MSG_SLAB_MAX_TAPS * 16 + 1024 + p.b_resident;
  };
  // two output rows per tile where the weights are streamed (C = 128): one weight fetch serves 256 pixels
  static const int env_trows = [] { const char* e = getenv("MSG_SLAB_TROWS"); return e ? atoi(e) : 0; }();
  p.trows = ((spec == 128 || spec == 0) && p.a_mode != 0 && !d->pixel_pair_k && d->n_chains == 1 && !p.b_resident &&
             4 * d->Ntot <= 512 && d->H % 2 == 0 && env_trows != 1 &&
             2 * (2 * p.a_bytes + p.b_bytes) + fixed_for(1) <= 220 * 1024) ? 2 : 1;   // two double-row stages must fit
  if (p.trows == 2) { p.a_rows = 2; p.a_bytes *= 2; }
  p.tmem_cols = 32;
  while (p.tmem_cols < 2 * p.trows * d->Ntot * d->n_chains) p.tmem_cols <<= 1;
  MSG_REQUIRE(p.tmem_cols <= 512, MSG_ERR_SHAPE, "conv_slab: accumulators do not fit in TMEM");
  const int stage_bytes = p.a_bytes + p.b_bytes;
  static const int env_groups = [] { const char* e = getenv("MSG_SLAB_EPI_GROUPS"); return e ? atoi(e) : 0; }();
  static const int env_issuers = [] { const char* e = getenv("MSG_SLAB_ISSUERS"); return e ? atoi(e) : 0; }();
  // A second epilogue group (every other tile) is available (MSG_SLAB_EPI_GROUPS=2) but off by default: measured on
  // B200 the programs here are bound by shared-memory bandwidth (A re-read by every small-N MMA), not by the
  // epilogue's instruction stream, and the group's staging costs pipeline stages (MSB C=64: 0.78 -> 0.90 ms).
  // The 7x7 input conv (28 MMAs per tile once its issue loop is straight-line code) IS bound by the epilogue's
  // instruction stream (ncu: issuer waits on tempty 42 %, epilogue warps 86 % busy): second group, 0.54 -> 0.43 ms.
  // Same for any light program (<= 64 MMAs per tile, e.g. a transposed-conv phase at N = 64).
  const int mmas_per_tile = d->pixel_pair_k ? d->n_taps : 4 * d->n_taps * p.trows;
  p.epi_groups = (mmas_per_tile <= 64 && (220 * 1024 - fixed_for(2)) / stage_bytes >= 3) ? 2 : 1;
  if (env_groups == 1 || env_groups == 2) p.epi_groups = env_groups;
  // Two issuers split the taps by accumulator column slice (disjoint TMEM columns => no ordering between the two
  // instruction streams): slices sorted by tap count, each given to the less loaded issuer.  Programs with a single
  // slice (dgrad, 7x7 convs) and the pixel-pair / chunk-plane modes keep one issuer.
  p.issuers = 1;
  for (int tp = 0; tp < MSG_SLAB_MAX_TAPS; ++tp) p.tap_owner[tp] = 0;
  if (spec == 0 && p.a_mode != 0 && d->n_chains == 1 && env_issuers != 1) {
    int cols[MSG_SLAB_MAX_TAPS], cnt[MSG_SLAB_MAX_TAPS], ns = 0;
    for (int tp = 0; tp < d->n_taps; ++tp) {
      int i = 0;
      while (i < ns && cols[i] != d->tap_acc_col[tp]) ++i;
      if (i == ns) { cols[ns] = d->tap_acc_col[tp]; cnt[ns] = 0; ++ns; }
      ++cnt[i];
    }
    if (ns >= 2) {
      int load[2] = {0, 0}, owner[MSG_SLAB_MAX_TAPS];
      bool used[MSG_SLAB_MAX_TAPS] = {false};
      for (int k = 0; k < ns; ++k) {                  // largest slice first
        int best = -1;
        for (int i = 0; i < ns; ++i)
          if (!used[i] && (best < 0 || cnt[i] > cnt[best])) best = i;
        used[best] = true;
        const int o = load[1] < load[0] ? 1 : 0;
        owner[best] = o; load[o] += cnt[best];
      }
      for (int tp = 0; tp < d->n_taps; ++tp) {
        int i = 0;
        while (cols[i] != d->tap_acc_col[tp]) ++i;
        p.tap_owner[tp] = (unsigned char)owner[i];
      }
      p.issuers = 2;
    }
  }
  static const int env_il = [] { const char* e = getenv("MSG_SLAB_INTERLEAVE"); return e ? atoi(e) : -1; }();
  // Contiguous ranges by default.  Measured on B200 (16 images): round-robin (MSG_SLAB_INTERLEAVE=1) removes the DRAM
  // re-reads of the C = 64 MSB program (1.8 -> 0.6 GB) but not its time (shared-memory bound: 0.79 -> 0.80 ms), is neutral for
  // C = 128 (0.50 -> 0.49 ms) and costs where an image has few tiles (statistics flushed with fp64 atomics every few tiles:
  // transposed-conv phases 0.49 -> 0.82 ms, 7x7 input conv 0.40 -> 0.45 ms).
  // Round 2: on for the two specialised MSB programs -- same time, but 2-3x less DRAM traffic for the other streams' kernels
  // that run beside them in the stylise step (ncu: 1.80 -> ~0.6 GB read for a 0.54 GB input at C = 64).
  p.interleave = env_il >= 0 ? env_il : (spec == 64 || spec == 128 ? 1 : 0);
  const int fixed = fixed_for(p.epi_groups);
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
    cuuint32_t box[4] = {(cuuint32_t)(p.a_mode != 0 ? 64 : 8), (cuuint32_t)p.Ws, (cuuint32_t)p.a_rows, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    void* base = (void*)((const __nv_bfloat16*)x + d->ci_off);
    CUresult r = enc(&mapA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, p.a_mode != 0 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MSG_REQUIRE(r == CUDA_SUCCESS, MSG_ERR_CUDA, "conv_slab: cuTensorMapEncodeTiled(A) failed with %d", (int)r);
  }
  {
    cuuint64_t dims[2] = {64, (cuuint64_t)(d->pixel_pair_k ? d->n_kblocks : d->n_taps) * d->ncols};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {64, (cuuint32_t)(d->ncols * p.b_box_taps)};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&mapB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)w_slab, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MSG_REQUIRE(r == CUDA_SUCCESS, MSG_ERR_CUDA, "conv_slab: cuTensorMapEncodeTiled(B) failed with %d", (int)r);
  }
  static DeviceOnce attr_set;     // cudaFuncSetAttribute is per device
  if (attr_set.needed()) {
    cudaError_t e = cudaFuncSetAttribute(conv_slab_kernel<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_slab_kernel<64, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_slab_kernel<128, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_slab_kernel<128, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_slab_kernel<7, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_slab_kernel<0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    MSG_REQUIRE(e == cudaSuccess, MSG_ERR_CUDA, "conv_slab: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set.done();
  }
  int grid = sm_count();
  const long long total = (long long)d->N * (d->H / p.trows) * p.segs;
  MSG_REQUIRE(total < 0x7fffffffLL, MSG_ERR_SHAPE, "conv_slab: too many tiles");
  if (grid > total) grid = (int)total;
  const int nthreads = 64 + 128 * p.epi_groups + (p.issuers == 2 ? 32 : 0);
  if (spec == 64) conv_slab_kernel<64, 1><<<grid, nthreads, smem, as_stream(stream)>>>(mapA, mapB, p);
  else if (spec == 128 && p.trows == 2) conv_slab_kernel<128, 2><<<grid, nthreads, smem, as_stream(stream)>>>(mapA, mapB, p);
  else if (spec == 128) conv_slab_kernel<128, 1><<<grid, nthreads, smem, as_stream(stream)>>>(mapA, mapB, p);
  else if (spec == 7) conv_slab_kernel<7, 1><<<grid, nthreads, smem, as_stream(stream)>>>(mapA, mapB, p);
  else if (p.trows == 2) conv_slab_kernel<0, 2><<<grid, nthreads, smem, as_stream(stream)>>>(mapA, mapB, p);
  else conv_slab_kernel<0, 1><<<grid, nthreads, smem, as_stream(stream)>>>(mapA, mapB, p);
  return check_launch("conv_slab_kernel");
}
