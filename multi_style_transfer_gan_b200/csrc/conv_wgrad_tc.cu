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
  const int pair0 = grp * MAXP;
  const int npair = (p.n_pairs - pair0) < MAXP ? (p.n_pairs - pair0) : MAXP;
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
  const int stage_bytes = (2 + MAXP) * TILE;

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
        mbar_expect_tx(full_bar(s), (uint32_t)((2 + npair) * TILE));
        // dy: two 64-channel boxes of this co tile (channels >= Cout are out of bounds -> zeros)
        const int oy = i0 * d.out_stride + d.out_off_h, ox = j0 * d.out_stride + d.out_off_w;
        tma_load_4d(st, &mapDY, full_bar(s), cot * 128, ox, oy, img);
        tma_load_4d(st + TILE, &mapDY, full_bar(s), cot * 128 + 64, ox, oy, img);
        // x: one shifted / strided box per (tap, ci-block) pair
        const int w_base = j0 * d.in_stride - d.pad_w, h_base = i0 * d.in_stride - d.pad_h;
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
    uint32_t it = 0;
    for (int mt = mt_begin; mt < mt_end; ++mt, ++it) {
      const int s = it % S;
      mbar_wait(full_bar(s), (it / S) & 1);
      tc_fence_after();
      const uint32_t st = base + s * stage_bytes;
      const uint64_t da = make_mn_sw128_desc(st, TILE);
      const uint64_t db = make_mn_sw128_desc(st + 2 * TILE, TILE);
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < BM / 16; ++ks)       // 16 pixels per MMA = 2 swizzle atoms = 2048 B further down
          umma_bf16(tmem_base, da + (uint64_t)(ks * 128), db + (uint64_t)(ks * 128), idesc, (it | ks) != 0);
      }
      __syncwarp();
      if (elect_one()) umma_commit(empty_bar(s));
    }
    __syncwarp();
    if (elect_one()) umma_commit(done_bar);
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
      const int pitch = npair * 256 + 16;                        // bytes; + 16: conflict-free 16-byte stores down a column of rows
      uint8_t* srow = gen + (size_t)row * pitch;
      for (int pi = 0; pi < npair; ++pi) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          __syncwarp();
          float v[32];
          tmem_ld32_sync(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(pi * 64 + h * 32), v);
          float4* o = reinterpret_cast<float4*>(srow + pi * 256 + h * 128);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
      }
      fence_proxy_async();
      if (co < d.Cout) {
        for (int pi = 0; pi < npair; ++pi) {
          const int pr = pair0 + pi;
          const int tap = pr / p.cblocks, cb = pr - tap * p.cblocks;
          float* dst = p.dw + (size_t)out_img * d.Cout * Ktot + (size_t)co * Ktot + (size_t)tap * d.Cin + cb * 64;
          asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], 256;"
                       ::"l"(dst), "r"(base + (uint32_t)(row * pitch + pi * 256)) : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");     // the staging must outlive the reads
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
  const size_t smem = (size_t)p.stages * (2 + MAXP) * TILE + 8 * (2 * p.stages + 1) + 16 + 1024;

  CUtensorMap mapX, mapDY;
  {
    cuuint64_t dims[4] = {(cuuint64_t)d->Cin, (cuuint64_t)d->Wi, (cuuint64_t)d->Hi, (cuuint64_t)d->N};
    cuuint64_t strides[3] = {(cuuint64_t)d->Ci_total * 2, (cuuint64_t)d->Wi * d->Ci_total * 2,
                             (cuuint64_t)d->Hi * d->Wi * d->Ci_total * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)(p.Wt * d->in_stride), (cuuint32_t)(p.R * d->in_stride), 1};
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
