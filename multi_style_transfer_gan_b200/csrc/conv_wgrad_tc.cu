// tcgen05 weight-gradient kernel (sm_100a, bf16 operands, fp32 accumulate).
//
//   dw[co][tap][ci] += sum over output pixels p of  dy[p][co] * x[p @ tap][ci]
// is a GEMM whose CONTRACTION dimension is the pixel index: both operands are "MN-major" (channels
// contiguous, pixels along K).  That is exactly how a TMA box of the NHWC tensors lands in shared
// memory -- rows = pixels, 128 B = 64 channels, 128B swizzle -- so no transposition is needed:
//   A = dy tile  [128 pixels x 128 co]  (two 64-channel boxes, LBO = 16 KB apart)
//   B = x tiles  [128 pixels x 64 ci] for up to FOUR (tap, ci-block) pairs, 16 KB apart = N up to 256
// and one k-block (128 pixels) is 8 x tcgen05.mma (M=128, N=256, K=16) with a_major = b_major = MN.
// The x tile of a tap is the same shifted / strided TMA box the forward kernel uses (zero fill =
// padding).  A CTA owns one (co tile, group of <=4 pairs) and a contiguous range of pixel tiles; the
// fp32 accumulator stays in TMEM for the whole range and is added to dw (packed [Cout][tap][Cin] fp32)
// with atomics at the end.  Channels beyond Cout are TMA out-of-bounds = zero rows.
// Row-slab mode (stride-1 convs with KW > 1: the 3x3 / 7x7 convs and the 2x2 transposed-conv phases).  Per-tap x tiles re-fetch
// every input pixel KW times (ncu: 540 MB of L2 -> SM traffic for a 33 MB tensor in a 3x3 weight gradient).  Here a CTA owns
// (kh, ci-block) GROUPS: per group ONE box of [R rows x (Wt + (KW-1) dil) pixels x 64 ch] lands per k-block, and the KW
// horizontal taps are KW overlapping N blocks of a single MN-major operand: LBO (the byte distance between 64-element N
// blocks) = dil * 128 bytes = `dil` pixels further along the slab -- the 128B swizzle is a function of the absolute shared-memory
// address, so shifted views need no re-layout (as in conv_slab.cu).  One tcgen05.mma of N = KW * 64 per K step and group.
#include <cuda.h>

#include "common.cuh"
#include "tcgen05.cuh"

namespace msg {
namespace {
using namespace tc;

constexpr int BM = 128;           // pixels per k-block
constexpr int TILE = BM * 128;    // one [128 px x 64 ch] box
constexpr int NTHREADS = 192;
constexpr int MAXP = 4;           // (tap, ci-block) pairs per CTA


struct WgParams {
  msg_conv_desc d;
  float* dw;
  int Wt, R;             // pixel tile = R rows x Wt pixels of the GEMM grid
  int m_tiles;           // pixel tiles in total
  int cblocks;           // Cin / 64
  int n_pairs;           // KH*KW*cblocks
  int groups;            // ceil(n_pairs / MAXP)
  int co_tiles;          // ceil(Cout / 128)
  int splits;            // pixel-range splits
  int stages;
  int per_image;         // 1: every image accumulates into its own [Cout][K] block (batched Gram matrices); the
                         //    pixel-range splits then never straddle an image (splits = N * splits_per_image)
  int splits_per_image;
  int slab;              // row-slab mode: 1
  int n_g, gmax;         // (kh, ci-block) groups in total / per CTA
  int slab_rows, slab_bytes, h2;     // pixel rows of one slab box, its 1024-byte-rounded size, (KW - 1) * dil
};

__global__ void __launch_bounds__(NTHREADS, 1)
conv_wgrad_tc_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapDY,
                     const WgParams p) {
  extern __shared__ uint8_t smem_raw[];
  const msg_conv_desc& d = p.d;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = p.stages;
  // work item of this CTA
  int w = blockIdx.x;
  const int split = w % p.splits; w /= p.splits;
  const int grp = w % p.groups;
  const int cot = w / p.groups;
  // pairs mode: up to MAXP (tap, ci-block) pairs; slab mode: up to gmax (kh, ci-block) groups of KW taps each
  const int g0 = grp * p.gmax;
  const int ng = p.slab ? ((p.n_g - g0) < p.gmax ? (p.n_g - g0) : p.gmax) : 0;
  const int pair0 = p.slab ? 0 : grp * MAXP;
  const int npair = p.slab ? ng * d.KW : ((p.n_pairs - pair0) < MAXP ? (p.n_pairs - pair0) : MAXP);
  int mt_begin, mt_end, out_img = 0;
  if (p.per_image) {
    const int tpi = p.m_tiles / d.N;                       // M tiles per image
    out_img = split / p.splits_per_image;
    const int sub = split - out_img * p.splits_per_image;
    mt_begin = out_img * tpi + (int)((long long)sub * tpi / p.splits_per_image);
    mt_end = out_img * tpi + (int)((long long)(sub + 1) * tpi / p.splits_per_image);
  } else {
    mt_begin = (int)((long long)split * p.m_tiles / p.splits);
    mt_end = (int)((long long)(split + 1) * p.m_tiles / p.splits);
  }
  // co tiles with <= 64 channels load ONE 64-channel dy box; the second M block of the A operand aliases it (LBO = 0): rows
  // 64-127 of the accumulator repeat rows 0-63 and are never written -- 16 KB less L2 -> SM traffic per k-block (the MSB branch,
  // 7x7 and 128 -> 64 transposed-conv gradients were bound by it)
  const int ndy = (d.Cout - cot * 128) <= 64 ? 1 : 2;
  const int stage_bytes = p.slab ? 2 * TILE + p.gmax * p.slab_bytes : (2 + MAXP) * TILE;

  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sBar = base + S * stage_bytes;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + S * stage_bytes + 8 * (2 * S + 1));
  auto full_bar = [&](int s) { return sBar + 8u * s; };
  auto empty_bar = [&](int s) { return sBar + 8u * (S + s); };
  const uint32_t done_bar = sBar + 8u * (2 * S);
  const uint32_t ncols = (uint32_t)(npair * 64);
  uint32_t tmem_cols = 32;
  while (tmem_cols < ncols) tmem_cols <<= 1;

  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
      mbar_init(done_bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapX)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapDY)) : "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int tiles_per_row = d.Wg / p.Wt;
  const int tile_rows = d.Hg / p.R;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      uint32_t it = 0;
      for (int mt = mt_begin; mt < mt_end; ++mt, ++it) {
        const int img = mt / (tile_rows * tiles_per_row);
        const int rem = mt - img * (tile_rows * tiles_per_row);
        const int rg = rem / tiles_per_row, cg = rem - rg * tiles_per_row;
        const int i0 = rg * p.R, j0 = cg * p.Wt;
        const int s = it % S;
        if (it >= (uint32_t)S) mbar_wait(empty_bar(s), ((it / S) - 1) & 1);
        const uint32_t st = base + s * stage_bytes;
        mbar_expect_tx(full_bar(s), p.slab ? (uint32_t)(ndy * TILE + ng * p.slab_rows * 128) : (uint32_t)((ndy + npair) * TILE));
        // dy: two 64-channel boxes of this co tile (channels >= Cout are out of bounds -> zeros)
        const int oy = i0 * d.out_stride + d.out_off_h, ox = j0 * d.out_stride + d.out_off_w;
        tma_load_4d(st, &mapDY, full_bar(s), cot * 128, ox, oy, img);
        if (ndy == 2) tma_load_4d(st + TILE, &mapDY, full_bar(s), cot * 128 + 64, ox, oy, img);
        // x: one shifted / strided box per (tap, ci-block) pair
        const int w_base = j0 * d.in_stride - d.pad_w, h_base = i0 * d.in_stride - d.pad_h;
        if (p.slab) {
          for (int gi = 0; gi < ng; ++gi) {           // one slab per (kh, ci-block) group: all KW taps read it
            const int gg = g0 + gi;
            const int th = gg / p.cblocks, cb = gg - th * p.cblocks;
            tma_load_4d(st + 2 * TILE + gi * p.slab_bytes, &mapX, full_bar(s), cb * 64, w_base, h_base + th * d.dil, img);
          }
        } else
        for (int pi = 0; pi < npair; ++pi) {
          const int pr = pair0 + pi;
          const int tap = pr / p.cblocks, cb = pr - tap * p.cblocks;
          const int th = tap / d.KW, tw = tap - th * d.KW;
          tma_load_4d(st + (2 + pi) * TILE, &mapX, full_bar(s), cb * 64, w_base + tw * d.dil, h_base + th * d.dil, img);
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =====================================
    // D[co][pair*64 + ci] += A^T B ; both operands MN-major (bits 15, 16), M = 128, N = npair*64
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                           ((uint32_t)(ncols >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const bool leader = elect_one();           // one election for the whole kernel (a commit only tracks its own thread's MMAs)
    // slab mode: tile-invariant operand words
    const uint32_t mn_hi = (uint32_t)(make_mn_sw128_desc(0, 0) >> 32);
    const uint32_t a_lbo = ndy == 2 ? (uint32_t)(TILE >> 4) << 16 : 0u, b_lbo = (uint32_t)((d.dil * 128) >> 4) << 16;
    const uint32_t slab_u = (uint32_t)p.slab_bytes >> 4, tap4_u = (uint32_t)(4 * d.dil * 128) >> 4;
    const int kw_first = d.KW < 4 ? d.KW : 4, kw_rest = d.KW - kw_first;
    const uint32_t idesc_n0 = idesc & ~(0x3fu << 17);
    const uint32_t idesc_a = idesc_n0 | ((uint32_t)((kw_first * 64) >> 3) << 17), idesc_b = idesc_n0 | ((uint32_t)((kw_rest * 64) >> 3) << 17);
    uint32_t koff[BM / 16];
#pragma unroll
    for (int ks = 0; ks < BM / 16; ++ks) {       // a K step (16 pixels) never straddles tile rows (Wt % 16 == 0)
      const int p0 = ks * 16, r = p0 / p.Wt, x0 = p0 - r * p.Wt;
      koff[ks] = (uint32_t)((r * (p.Wt + p.h2) + x0) * 128) >> 4;
    }
    uint32_t it = 0;
    for (int mt = mt_begin; mt < mt_end; ++mt, ++it) {
      const int s = it % S;
      mbar_wait(full_bar(s), (it / S) & 1);
      tc_fence_after();
      const uint32_t st = base + s * stage_bytes;
      const uint64_t da = make_mn_sw128_desc(st, ndy == 2 ? TILE : 0);
      const uint64_t db = make_mn_sw128_desc(st + 2 * TILE, TILE);
      if (p.slab) {
        if (leader) {
          // everything tile-invariant was computed before the loop: an MMA costs ~8 issue instructions (the first version built
          // 64-bit descriptors with a division per K step and was issue-bound: 18 vs 14 ms per train step)
          const uint32_t a_lo0 = (st >> 4) | a_lbo;
          const uint32_t b_st = (st + 2 * TILE) >> 4;
#pragma unroll
          for (int ks = 0; ks < BM / 16; ++ks) {
            for (int gi = 0; gi < ng; ++gi) {
              const uint32_t b_lo = (b_st + (uint32_t)gi * slab_u + koff[ks]) | b_lbo;
              const uint32_t dcol = tmem_base + (uint32_t)(gi * d.KW * 64);
              // taps kw = 0 .. KW-1 are the N blocks of ONE operand, `dil` pixels apart (at most 4 per MMA: N <= 256)
              umma_bf16_lo2(dcol, a_lo0 + (uint32_t)(ks * 128), mn_hi, b_lo, mn_hi, idesc_a, (it | ks) != 0);
              if (kw_rest) umma_bf16_lo2(dcol + 256, a_lo0 + (uint32_t)(ks * 128), mn_hi, b_lo + tap4_u, mn_hi, idesc_b, (it | ks) != 0);
            }
          }
        }
      } else if (leader) {
#pragma unroll
        for (int ks = 0; ks < BM / 16; ++ks)       // 16 pixels per MMA = 2 swizzle atoms = 2048 B further down
          umma_bf16(tmem_base, da + (uint64_t)(ks * 128), db + (uint64_t)(ks * 128), idesc, (it | ks) != 0);
      }
      __syncwarp();
      if (leader) umma_commit(empty_bar(s));
    }
    __syncwarp();
    if (leader) umma_commit(done_bar);
  } else {
    // ===================================== epilogue: TMEM -> shared memory -> bulk reduce-add on dw =====================
    // A thread owns one co row.  The row's npair x 64 fp32 values are staged in the (now idle) pipeline memory and added to dw
    // with ONE cp.reduce.async.bulk (256 contiguous bytes) per (row, pair): the scalar atomicAdd version issued 64 atomics per
    // thread and pair, each warp instruction touching 32 different rows -- the fixed ~50 us that made every weight-gradient
    // launch of the train step cost the same whatever its size.
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int co = cot * 128 + row;
    mbar_wait(done_bar, 0);
    tc_fence_after();
    const int Ktot = d.KH * d.KW * d.Cin;
    if (mt_end > mt_begin) {
      // staged in chunks of `pc` pairs: 128 rows x (pc * 256 + 16) bytes must fit the pipeline memory (up to 7 pairs = 448
      // columns in slab mode).  A thread only ever touches its own row, so a chunk is reused as soon as the thread's own
      // bulk reads of the previous one are done.
      int pc = (S * stage_bytes / 128 - 16) / 256;
      if (pc > npair) pc = npair;
      const int pitch = pc * 256 + 16;                           // bytes; + 16: conflict-free 16-byte stores down a column of rows
      uint8_t* srow = gen + (size_t)row * pitch;
      for (int pc0 = 0; pc0 < npair; pc0 += pc) {
        const int pn = npair - pc0 < pc ? npair - pc0 : pc;
        for (int pj = 0; pj < pn; ++pj) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            __syncwarp();
            float v[32];
            tmem_ld32_sync(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((pc0 + pj) * 64 + h * 32), v);
            float4* o = reinterpret_cast<float4*>(srow + pj * 256 + h * 128);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
        }
        fence_proxy_async();
        if (co < d.Cout) {
          for (int pj = 0; pj < pn; ++pj) {
            const int pi = pc0 + pj;
            int tap, cb;
            if (p.slab) {                              // accumulator block pi = (group gi, horizontal tap kw)
              const int gi = pi / d.KW, kw = pi - gi * d.KW, gg = g0 + gi;
              const int th = gg / p.cblocks;
              cb = gg - th * p.cblocks;
              tap = th * d.KW + kw;
            } else {
              const int pr = pair0 + pi;
              tap = pr / p.cblocks;
              cb = pr - tap * p.cblocks;
            }
            float* dst = p.dw + (size_t)out_img * d.Cout * Ktot + (size_t)co * Ktot + (size_t)tap * d.Cin + cb * 64;
            asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], 256;"
                         ::"l"(dst), "r"(base + (uint32_t)(row * pitch + pj * 256)) : "memory");
          }
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");     // the staging must outlive the reads
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}


bool pick_tiling(const msg_conv_desc* d, int* Wt, int* R) {
  if (d->Wg >= 128) {
    if (d->Wg % 128) return false;
    *Wt = 128; *R = 1;
    return true;
  }
  if (128 % d->Wg) return false;
  *Wt = d->Wg; *R = 128 / d->Wg;
  return d->Hg % *R == 0;
}

}  // namespace

bool conv2d_wgrad_tc_supported(const msg_conv_desc* d, const void* x, const void* dy) {
  if (d->dtype != MSG_BF16) return false;
  if (d->flags & (MSG_CONV_IN_NORM | MSG_CONV_OUT_NCHW_F32)) return false;
  if (d->Cin % 64 || (d->Ci_total & 7) || (d->ci_off & 7)) return false;
  if ((d->Co_total & 7) || (d->co_off & 7)) return false;
  if (((uintptr_t)x | (uintptr_t)dy) & 15) return false;
  if ((d->in_stride != 1 && d->in_stride != 2) || (d->out_stride != 1 && d->out_stride != 2)) return false;
  int Wt, R;
  if (!pick_tiling(d, &Wt, &R)) return false;
  if (Wt * d->in_stride > 256 || R * d->in_stride > 256 || Wt * d->out_stride > 256 || R * d->out_stride > 256) return false;
  return get_encode() != nullptr;
}

int conv2d_wgrad_tc_impl(const msg_conv_desc* d, const void* x, const void* dy, float* dw, bool per_image, cudaStream_t st);
int conv2d_wgrad_tc(const msg_conv_desc* d, const void* x, const void* dy, float* dw, cudaStream_t st) {
  return conv2d_wgrad_tc_impl(d, x, dy, dw, false, st);
}
// dw[n] += dY[n]^T X[n] for every image n separately: dw is [N][Cout][KH*KW*Cin] fp32 (zeroed by the caller).
// With x == dy and a 1x1 geometry this is the batch of Gram matrices F^T F (gram.cu).
int conv2d_wgrad_tc_per_image(const msg_conv_desc* d, const void* x, const void* dy, float* dw, cudaStream_t st) {
  return conv2d_wgrad_tc_impl(d, x, dy, dw, true, st);
}

int conv2d_wgrad_tc_impl(const msg_conv_desc* d, const void* x, const void* dy, float* dw, bool per_image, cudaStream_t st) {
  EncodeTiledFn enc = get_encode();
  MSG_REQUIRE(enc != nullptr, MSG_ERR_CUDA, "wgrad_tc: cuTensorMapEncodeTiled unavailable");
  WgParams p;
  p.d = *d; p.dw = dw;
  MSG_REQUIRE(pick_tiling(d, &p.Wt, &p.R), MSG_ERR_UNSUPPORTED, "wgrad_tc: unsupported plane geometry");
  p.m_tiles = (int)((long long)d->N * d->Hg * d->Wg / BM);
  p.cblocks = d->Cin / 64;
  p.n_pairs = d->KH * d->KW * p.cblocks;
  p.groups = (p.n_pairs + MAXP - 1) / MAXP;
  static const bool slab_ok = [] { const char* e = getenv("MSG_WGRAD_SLAB"); return !e || atoi(e) != 0; }();
  p.slab = 0; p.n_g = 0; p.gmax = 1; p.slab_rows = 0; p.slab_bytes = 0;
  p.h2 = (d->KW - 1) * d->dil;
  if (slab_ok && d->in_stride == 1 && d->KW >= 2 && d->KW * 64 <= 512 && p.Wt % 16 == 0 && p.Wt + p.h2 <= 256 &&
      p.R * (p.Wt + p.h2) * 128 <= 40 * 1024) {
    p.slab = 1;
    p.n_g = d->KH * p.cblocks;
    p.gmax = 512 / (d->KW * 64);
    if (p.gmax > p.n_g) p.gmax = p.n_g;
    p.slab_rows = p.R * (p.Wt + p.h2);
    p.slab_bytes = (p.slab_rows * 128 + 1023) / 1024 * 1024;
    while (p.gmax > 1 && 2 * (2 * TILE + p.gmax * p.slab_bytes) > 200 * 1024) --p.gmax;     // two pipeline stages must fit
    p.groups = (p.n_g + p.gmax - 1) / p.gmax;
  }
  p.co_tiles = (d->Cout + 127) / 128;
  const int items = p.groups * p.co_tiles;
  int splits = sm_count() / items;                       // ONE wave of CTAs (one CTA per SM: 192 KB of pipeline stages each)
  if (splits > p.m_tiles) splits = p.m_tiles;
  if (splits < 1) splits = 1;
  p.per_image = per_image ? 1 : 0;
  p.splits_per_image = 1;
  if (per_image) {
    MSG_REQUIRE(p.m_tiles % d->N == 0, MSG_ERR_SHAPE, "wgrad_tc: per-image mode needs whole tiles per image");
    const int tpi = p.m_tiles / d->N;
    int spi = sm_count() / (items * d->N);
    if (spi > tpi) spi = tpi;
    if (spi < 1) spi = 1;
    p.splits_per_image = spi;
    splits = spi * d->N;
  }
  p.splits = splits;
  p.stages = 2;
  const size_t stage_b = p.slab ? (size_t)(2 * TILE + p.gmax * p.slab_bytes) : (size_t)(2 + MAXP) * TILE;
  if (p.slab && 3 * stage_b + 4096 <= 220 * 1024) p.stages = 3;
  const size_t smem = (size_t)p.stages * stage_b + 8 * (2 * p.stages + 1) + 16 + 1024;

  CUtensorMap mapX, mapDY;
  {
    cuuint64_t dims[4] = {(cuuint64_t)d->Cin, (cuuint64_t)d->Wi, (cuuint64_t)d->Hi, (cuuint64_t)d->N};
    cuuint64_t strides[3] = {(cuuint64_t)d->Ci_total * 2, (cuuint64_t)d->Wi * d->Ci_total * 2,
                             (cuuint64_t)d->Hi * d->Wi * d->Ci_total * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)(p.slab ? p.Wt + p.h2 : p.Wt * d->in_stride), (cuuint32_t)(p.R * d->in_stride), 1};
    cuuint32_t es[4] = {1, (cuuint32_t)d->in_stride, (cuuint32_t)d->in_stride, 1};
    void* base = (void*)((const __nv_bfloat16*)x + d->ci_off);
    CUresult r = enc(&mapX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MSG_REQUIRE(r == CUDA_SUCCESS, MSG_ERR_CUDA, "wgrad_tc: cuTensorMapEncodeTiled(x) failed with %d", (int)r);
  }
  {
    cuuint64_t dims[4] = {(cuuint64_t)d->Cout, (cuuint64_t)d->Wo, (cuuint64_t)d->Ho, (cuuint64_t)d->N};
    cuuint64_t strides[3] = {(cuuint64_t)d->Co_total * 2, (cuuint64_t)d->Wo * d->Co_total * 2,
                             (cuuint64_t)d->Ho * d->Wo * d->Co_total * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)(p.Wt * d->out_stride), (cuuint32_t)(p.R * d->out_stride), 1};
    cuuint32_t es[4] = {1, (cuuint32_t)d->out_stride, (cuuint32_t)d->out_stride, 1};
    void* base = (void*)((const __nv_bfloat16*)dy + d->co_off);
    CUresult r = enc(&mapDY, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MSG_REQUIRE(r == CUDA_SUCCESS, MSG_ERR_CUDA, "wgrad_tc: cuTensorMapEncodeTiled(dy) failed with %d", (int)r);
  }
  static DeviceOnce attr_set;     // cudaFuncSetAttribute is per device
  if (attr_set.needed()) {
    cudaError_t e = cudaFuncSetAttribute(conv_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    MSG_REQUIRE(e == cudaSuccess, MSG_ERR_CUDA, "wgrad_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set.done();
  }
  const int grid = p.co_tiles * p.groups * p.splits;
  conv_wgrad_tc_kernel<<<grid, NTHREADS, smem, st>>>(mapX, mapDY, p);
  return check_launch("conv_wgrad_tc_kernel");
}

}  // namespace msg
