// SIMT (CUDA-core, fp32-accumulate) implicit-GEMM gather convolution: forward and wgrad.
//
// This is the fp32 parity engine (BASELINE.json: <=1e-4 vs the reference with TF32 off) and the
// path for shapes the tcgen05 kernel does not take (Cin % 8 != 0, tiny Cout).  The bf16 hot path
// is conv_tc.cu.  Same descriptor, same semantics (include/msg_b200.h).
#include "common.cuh"

namespace msg {

namespace {

constexpr int BM = 64, BN = 64, BK = 16, LDS = 68;

struct RowCoord {
  int n, ih0, iw0;   // image index, top-left input coordinate of tap (0,0)
  bool valid;
};

__device__ __forceinline__ RowCoord decode_row(const msg_conv_desc& d, long long m, long long M) {
  RowCoord r;
  r.valid = m < M;
  long long mm = r.valid ? m : 0;
  int hw = d.Hg * d.Wg;
  r.n = (int)(mm / hw);
  int rem = (int)(mm - (long long)r.n * hw);
  int i = rem / d.Wg, j = rem - i * d.Wg;
  r.ih0 = i * d.in_stride - d.pad_h;
  r.iw0 = j * d.in_stride - d.pad_w;
  return r;
}

__device__ __forceinline__ size_t out_pixel_offset(const msg_conv_desc& d, long long m) {
  int hw = d.Hg * d.Wg;
  int n = (int)(m / hw);
  int rem = (int)(m - (long long)n * hw);
  int i = rem / d.Wg, j = rem - i * d.Wg;
  int oh = i * d.out_stride + d.out_off_h, ow = j * d.out_stride + d.out_off_w;
  return ((size_t)n * d.Ho + oh) * d.Wo + ow;
}

// Gather 4 consecutive K elements (k0..k0+3) of the im2col row `rc` into v[].
template <typename T, bool VEC>
__device__ __forceinline__ void gather4(const msg_conv_desc& d, const T* __restrict__ x,
                                        const double* __restrict__ in_stats, double inv_hw_in,
                                        const RowCoord& rc, int k0, int K, float (&v)[4]) {
  v[0] = v[1] = v[2] = v[3] = 0.f;
  if (!rc.valid || k0 >= K) return;
  if (VEC) {
    int tap = k0 / d.Cin, ci = k0 - tap * d.Cin;
    int th = tap / d.KW, tw = tap - th * d.KW;
    int ih = rc.ih0 + th * d.dil, iw = rc.iw0 + tw * d.dil;
    if (ih < 0 || ih >= d.Hi || iw < 0 || iw >= d.Wi) return;
    size_t off = (((size_t)rc.n * d.Hi + ih) * d.Wi + iw) * d.Ci_total + d.ci_off + ci;
    load4(x + off, v);
    if (d.flags & MSG_CONV_IN_NORM) {
      const double* st = in_stats + ((size_t)rc.n * d.Ci_total + d.ci_off + ci) * 2;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float mean, rstd;
        finalize_stats(st[2 * e], st[2 * e + 1], inv_hw_in, mean, rstd);
        v[e] = apply_act((v[e] - mean) * rstd, d.in_act);
      }
    }
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      int k = k0 + e;
      if (k >= K) break;
      int tap = k / d.Cin, ci = k - tap * d.Cin;
      int th = tap / d.KW, tw = tap - th * d.KW;
      int ih = rc.ih0 + th * d.dil, iw = rc.iw0 + tw * d.dil;
      if (ih < 0 || ih >= d.Hi || iw < 0 || iw >= d.Wi) continue;
      size_t off = (((size_t)rc.n * d.Hi + ih) * d.Wi + iw) * d.Ci_total + d.ci_off + ci;
      float val = to_f<T>(x[off]);
      if (d.flags & MSG_CONV_IN_NORM) {
        const double* st = in_stats + ((size_t)rc.n * d.Ci_total + d.ci_off + ci) * 2;
        float mean, rstd;
        finalize_stats(st[0], st[1], inv_hw_in, mean, rstd);
        val = apply_act((val - mean) * rstd, d.in_act);
      }
      v[e] = val;
    }
  }
}

template <typename T, bool VEC>
__global__ void __launch_bounds__(256)
conv_simt_kernel(const msg_conv_desc d, const T* __restrict__ x, const T* __restrict__ w,
                 const float* __restrict__ bias, void* __restrict__ yv, double* __restrict__ stats,
                 const double* __restrict__ in_stats) {
  __shared__ __align__(16) float As[BK][LDS];
  __shared__ __align__(16) float Bs[BK][LDS];
  const int tid = threadIdx.x;
  const int K = d.KH * d.KW * d.Cin;
  const long long M = (long long)d.N * d.Hg * d.Wg;
  const long long m0 = (long long)blockIdx.x * BM;
  const int co0 = blockIdx.y * BN;
  const double inv_hw_in = 1.0 / ((double)d.Hi * (double)d.Wi);

  const int lrow = tid >> 2, kc = (tid & 3) * 4;
  const RowCoord rc = decode_row(d, m0 + lrow, M);
  const int bco = co0 + lrow;
  const bool bvalid = bco < d.Cout;
  const T* wrow = w + (size_t)(bvalid ? bco : 0) * K;

  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;

  for (int kt = 0; kt < K; kt += BK) {
    float av[4], bv[4];
    gather4<T, VEC>(d, x, in_stats, inv_hw_in, rc, kt + kc, K, av);
    bv[0] = bv[1] = bv[2] = bv[3] = 0.f;
    if (bvalid) {
      int k0 = kt + kc;
      if (VEC) {
        if (k0 < K) load4(wrow + k0, bv);
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (k0 + e < K) bv[e] = to_f<T>(wrow[k0 + e]);
      }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      As[kc + e][lrow] = av[e];
      Bs[kc + e][lrow] = bv[e];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---- epilogue: bias, (stats), activation, store
  float bcol[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int co = co0 + tx * 4 + j;
    bcol[j] = (bias != nullptr && co < d.Cout) ? bias[co] : 0.f;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] += bcol[j];

  if (d.flags & MSG_CONV_STATS) {
    const int hw = d.Hg * d.Wg;
    long long mlast = m0 + BM - 1 < M ? m0 + BM - 1 : M - 1;
    int n_first = (int)(m0 / hw), n_last = (int)(mlast / hw);
    // fp32 tile partials only when tiles never straddle images (plane a multiple of the tile): otherwise the grouping of an
    // image's rows would depend on its position in the batch, and with it the last bits of its statistics
    if (n_first == n_last && hw % BM == 0) {
      float s[4] = {0, 0, 0, 0}, ss[4] = {0, 0, 0, 0};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (m0 + ty * 4 + i < M) {
#pragma unroll
          for (int j = 0; j < 4; ++j) { s[j] += acc[i][j]; ss[j] = fmaf(acc[i][j], acc[i][j], ss[j]); }
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) { As[ty][tx * 4 + j] = s[j]; Bs[ty][tx * 4 + j] = ss[j]; }
      __syncthreads();
      if (tid < BN && co0 + tid < d.Cout) {
        float ts = 0.f, tss = 0.f;
#pragma unroll
        for (int r = 0; r < 16; ++r) { ts += As[r][tid]; tss += Bs[r][tid]; }
        double* st = stats + ((size_t)n_first * d.Co_total + d.co_off + co0 + tid) * 2;
        atomicAdd(st, (double)ts);
        atomicAdd(st + 1, (double)tss);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        long long m = m0 + ty * 4 + i;
        if (m >= M) continue;
        int n = (int)(m / hw);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int co = co0 + tx * 4 + j;
          if (co >= d.Cout) continue;
          double* st = stats + ((size_t)n * d.Co_total + d.co_off + co) * 2;
          atomicAdd(st, (double)acc[i][j]);
          atomicAdd(st + 1, (double)acc[i][j] * (double)acc[i][j]);
        }
      }
    }
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    long long m = m0 + ty * 4 + i;
    if (m >= M) continue;
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = apply_act(acc[i][j], d.act);
    int co = co0 + tx * 4;
    if (d.flags & MSG_CONV_OUT_NCHW_F32) {
      int hw = d.Hg * d.Wg;
      int n = (int)(m / hw);
      int rem = (int)(m - (long long)n * hw);
      int ii = rem / d.Wg, jj = rem - ii * d.Wg;
      int oh = ii * d.out_stride + d.out_off_h, ow = jj * d.out_stride + d.out_off_w;
      float* y = reinterpret_cast<float*>(yv);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (co + j < d.Cout)
          y[(((size_t)n * d.Co_total + d.co_off + co + j) * d.Ho + oh) * d.Wo + ow] = o[j];
    } else {
      T* y = reinterpret_cast<T*>(yv);
      size_t off = out_pixel_offset(d, m) * d.Co_total + d.co_off + co;
      const bool accum = d.flags & MSG_CONV_ACCUM;
      if (VEC && co + 3 < d.Cout && ((d.Co_total | d.co_off) & 3) == 0) {
        if (accum) {
          float old[4];
          load4(y + off, old);
#pragma unroll
          for (int j = 0; j < 4; ++j) o[j] += old[j];
        }
        store4(y + off, o);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (co + j < d.Cout) y[off + j] = from_f<T>(accum ? o[j] + to_f<T>(y[off + j]) : o[j]);
      }
    }
  }
}

// wgrad: dw[co][k] += sum_m dy[m][co] * A[m][k]   (A = im2col gather of x)
constexpr int WP = 16;  // pixels per reduction chunk
template <typename T, bool VEC>
__global__ void __launch_bounds__(256)
conv_wgrad_simt_kernel(const msg_conv_desc d, const T* __restrict__ x, const T* __restrict__ dy,
                       float* __restrict__ dw, const double* __restrict__ in_stats,
                       long long pixels_per_split) {
  __shared__ __align__(16) float Ds[WP][LDS];  // dy tile  [pixel][co]
  __shared__ __align__(16) float As[WP][LDS];  // im2col   [pixel][k]
  const int tid = threadIdx.x;
  const int K = d.KH * d.KW * d.Cin;
  const long long M = (long long)d.N * d.Hg * d.Wg;
  const int k0b = blockIdx.x * 64, co0 = blockIdx.y * 64;
  long long m_begin = (long long)blockIdx.z * pixels_per_split;
  long long m_end = m_begin + pixels_per_split < M ? m_begin + pixels_per_split : M;
  const double inv_hw_in = 1.0 / ((double)d.Hi * (double)d.Wi);
  const int p = tid >> 4, c4 = (tid & 15) * 4;
  const int ty = tid >> 4, tx = tid & 15;
  const bool dy_vec = VEC && ((d.Co_total | d.co_off) & 3) == 0;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;

  for (long long mc = m_begin; mc < m_end; mc += WP) {
    long long m = mc + p;
    bool valid = m < m_end;
    float dv[4] = {0.f, 0.f, 0.f, 0.f}, av[4];
    if (valid) {
      size_t off = out_pixel_offset(d, m) * d.Co_total + d.co_off + co0 + c4;
      if (dy_vec && co0 + c4 + 3 < d.Cout) {
        load4(dy + off, dv);
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (co0 + c4 + e < d.Cout) dv[e] = to_f<T>(dy[off + e]);
      }
    }
    RowCoord rc = decode_row(d, m, valid ? M : 0);
    gather4<T, VEC>(d, x, in_stats, inv_hw_in, rc, k0b + c4, K, av);
    *reinterpret_cast<float4*>(&Ds[p][c4]) = make_float4(dv[0], dv[1], dv[2], dv[3]);
    *reinterpret_cast<float4*>(&As[p][c4]) = make_float4(av[0], av[1], av[2], av[3]);
    __syncthreads();
#pragma unroll
    for (int q = 0; q < WP; ++q) {
      float4 a4 = *reinterpret_cast<const float4*>(&Ds[q][ty * 4]);
      float4 b4 = *reinterpret_cast<const float4*>(&As[q][tx * 4]);
      float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int co = co0 + ty * 4 + i;
    if (co >= d.Cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int k = k0b + tx * 4 + j;
      if (k < K) atomicAdd(dw + (size_t)co * K + k, acc[i][j]);
    }
  }
}

// ---- weight packing -------------------------------------------------------------------------
__device__ __forceinline__ int convt_src_k(int phase_bit, int kprime) {
  // 4x4 s2 p1 transposed conv: output parity 0 uses kernel rows {3,1}, parity 1 uses {2,0},
  // for input offsets {-1,0} and {0,+1} respectively (descriptor: pad = 1 - parity, dil = 1).
  return phase_bit == 0 ? (kprime == 0 ? 3 : 1) : (kprime == 0 ? 2 : 0);
}

template <typename T>
__global__ void pack_weight_kernel(const float* __restrict__ w, int D0, int D1, int KH, int KW,
                                   int mode, const float* __restrict__ denom, T* __restrict__ out,
                                   long long total) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  float scale = denom ? 1.f / *denom : 1.f;
  float v;
  if (mode == MSG_PACK_FWD) {  // out[o][kh][kw][i] = w[o][i][kh][kw]
    int i = (int)(idx % D1); long long t = idx / D1;
    int kw = (int)(t % KW); t /= KW;
    int kh = (int)(t % KH); int o = (int)(t / KH);
    v = w[(((size_t)o * D1 + i) * KH + kh) * KW + kw];
  } else if (mode == MSG_PACK_DGRAD_S1) {  // out[i][th][tw][o] = w[o][i][KH-1-th][KW-1-tw]
    int o = (int)(idx % D0); long long t = idx / D0;
    int tw = (int)(t % KW); t /= KW;
    int th = (int)(t % KH); int i = (int)(t / KH);
    v = w[(((size_t)o * D1 + i) * KH + (KH - 1 - th)) * KW + (KW - 1 - tw)];
  } else {  // CONVT_PHASES: w[I=D0][O=D1][4][4] -> out[ph*2+pw][o][kh'][kw'][i]
    int i = (int)(idx % D0); long long t = idx / D0;
    int kwp = (int)(t % 2); t /= 2;
    int khp = (int)(t % 2); t /= 2;
    int o = (int)(t % D1); int phase = (int)(t / D1);
    int kh = convt_src_k(phase >> 1, khp), kw = convt_src_k(phase & 1, kwp);
    v = w[(((size_t)i * D1 + o) * 4 + kh) * 4 + kw];
  }
  out[idx] = from_f<T>(v * scale);
}

__global__ void unpack_wgrad_kernel(const float* __restrict__ src, int D0, int D1, int KH, int KW,
                                    int mode, float* __restrict__ dw, long long total) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  size_t dst;
  if (mode == MSG_PACK_FWD) {
    int i = (int)(idx % D1); long long t = idx / D1;
    int kw = (int)(t % KW); t /= KW;
    int kh = (int)(t % KH); int o = (int)(t / KH);
    dst = (((size_t)o * D1 + i) * KH + kh) * KW + kw;
  } else if (mode == MSG_PACK_DGRAD_S1) {
    int o = (int)(idx % D0); long long t = idx / D0;
    int tw = (int)(t % KW); t /= KW;
    int th = (int)(t % KH); int i = (int)(t / KH);
    dst = (((size_t)o * D1 + i) * KH + (KH - 1 - th)) * KW + (KW - 1 - tw);
  } else {
    int i = (int)(idx % D0); long long t = idx / D0;
    int kwp = (int)(t % 2); t /= 2;
    int khp = (int)(t % 2); t /= 2;
    int o = (int)(t % D1); int phase = (int)(t / D1);
    int kh = convt_src_k(phase >> 1, khp), kw = convt_src_k(phase & 1, kwp);
    dst = (((size_t)i * D1 + o) * 4 + kh) * 4 + kw;
  }
  dw[dst] += src[idx];  // each destination element has exactly one source: no race
}

template <typename T>
__global__ void __launch_bounds__(256)
bias_grad_kernel(const T* __restrict__ dy, long long rows, int C_total, int c_off, int C,
                 float* __restrict__ db) {
  // block: 32 channels x 8 row-lanes; grid.x over channel groups, grid.y over row slabs
  __shared__ float red[8][33];
  int c = blockIdx.x * 32 + (threadIdx.x & 31);
  int rl = threadIdx.x >> 5;
  float s = 0.f;
  if (c < C) {
    for (long long r = (long long)blockIdx.y * 8 + rl; r < rows; r += (long long)gridDim.y * 8)
      s += to_f<T>(dy[(size_t)r * C_total + c_off + c]);
  }
  red[rl][threadIdx.x & 31] = s;
  __syncthreads();
  if (rl == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x & 31];
    atomicAdd(db + c, t);
  }
}

// bf16 fast path (C, c_off, C_total multiples of 8; C/8 a power of two <= 256): 16-byte loads, 4 of them in flight per
// thread; a thread owns one 8-channel chunk for all its rows, the block's row-lanes are folded through shared memory and
// each block adds C partial sums with atomics.  (The scalar kernel above read 2 bytes per thread per row.)
__global__ void __launch_bounds__(256)
bias_grad_bf16_vec_kernel(const __nv_bfloat16* __restrict__ dy, long long rows, int C_total, int c_off, int C,
                          float* __restrict__ db) {
  __shared__ float red[256][9];
  const int tpr = C >> 3;                         // threads per row
  const int chunk = threadIdx.x & (tpr - 1), rl = threadIdx.x / tpr, rpb = 256 / tpr;   // row lane, rows per block step
  float s[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) s[e] = 0.f;
  const __nv_bfloat16* src = dy + c_off + chunk * 8;
  const long long step = (long long)gridDim.x * rpb;
  for (long long r = (long long)blockIdx.x * rpb + rl; r < rows; r += 4 * step) {
    uint4 raw[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (r + u * step < rows) raw[u] = *reinterpret_cast<const uint4*>(src + (size_t)(r + u * step) * C_total);
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (r + u * step < rows) {
        float v[8];
        unpack8(raw[u], v);
#pragma unroll
        for (int e = 0; e < 8; ++e) s[e] += v[e];
      }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) red[threadIdx.x][e] = s[e];
  __syncthreads();
  if (threadIdx.x < C) {                          // thread c folds the row lanes of channel c
    const int ch = threadIdx.x >> 3, e = threadIdx.x & 7;
    float t = 0.f;
    for (int l = 0; l < rpb; ++l) t += red[l * tpr + ch][e];
    atomicAdd(db + threadIdx.x, t);
  }
}

int validate_desc(const msg_conv_desc* d) {
  MSG_REQUIRE(d != nullptr, MSG_ERR_SHAPE, "conv: null descriptor");
  MSG_REQUIRE(d->dtype == MSG_F32 || d->dtype == MSG_BF16, MSG_ERR_UNSUPPORTED, "conv: bad dtype %d", d->dtype);
  MSG_REQUIRE(d->N > 0 && d->Hi > 0 && d->Wi > 0 && d->Cin > 0 && d->Cout > 0 && d->Hg > 0 && d->Wg > 0,
              MSG_ERR_SHAPE, "conv: non-positive dimension");
  MSG_REQUIRE(d->ci_off >= 0 && d->ci_off + d->Cin <= d->Ci_total, MSG_ERR_SHAPE, "conv: input channel slice out of range");
  MSG_REQUIRE(d->co_off >= 0 && d->co_off + d->Cout <= d->Co_total, MSG_ERR_SHAPE, "conv: output channel slice out of range");
  MSG_REQUIRE(d->KH > 0 && d->KW > 0 && d->in_stride > 0 && d->dil > 0 && d->out_stride > 0, MSG_ERR_SHAPE, "conv: bad tap geometry");
  MSG_REQUIRE((d->Hg - 1) * d->out_stride + d->out_off_h < d->Ho && (d->Wg - 1) * d->out_stride + d->out_off_w < d->Wo &&
                  d->out_off_h >= 0 && d->out_off_w >= 0,
              MSG_ERR_SHAPE, "conv: output grid exceeds the output tensor");
  MSG_REQUIRE(!((d->flags & MSG_CONV_STATS) && d->act != MSG_ACT_NONE), MSG_ERR_UNSUPPORTED,
              "conv: stats epilogue requires act == NONE");
  MSG_REQUIRE(!((d->flags & MSG_CONV_ACCUM) && (d->flags & (MSG_CONV_OUT_NCHW_F32 | MSG_CONV_STATS))), MSG_ERR_UNSUPPORTED,
              "conv: ACCUM cannot be combined with NCHW output or stats");
  return MSG_OK;
}

}  // namespace

// ---- Cout <= 2, long K (the discriminator heads: 4x4 conv 512 -> 1 on a 15x15 map, enhanced_generator.py:256,265): a GEMM
// with one useful column.  On the tensor-core gather kernel it is 15 CTAs each walking 128 K blocks in sequence (0.13 ms for
// 29 MFLOP, 20 launches per train step); here one WARP owns an output pixel, its lanes stride over the 16-byte chunks of the
// (tap, channel) axis, and a shuffle tree finishes the dot product.
template <int CO>
__global__ void __launch_bounds__(256)
conv_dot_kernel(const msg_conv_desc d, const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                const float* __restrict__ bias, __nv_bfloat16* __restrict__ y) {
  const long long m = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const long long M = (long long)d.N * d.Hg * d.Wg;
  if (m >= M) return;
  const int lane = threadIdx.x & 31;
  const int n = (int)(m / (d.Hg * d.Wg));
  const int rem = (int)(m - (long long)n * d.Hg * d.Wg);
  const int i = rem / d.Wg, j = rem - i * d.Wg;
  const int c8 = d.Cin >> 3, chunks = d.KH * d.KW * c8;
  const int K = d.KH * d.KW * d.Cin;
  float acc[CO];
#pragma unroll
  for (int c = 0; c < CO; ++c) acc[c] = 0.f;
  for (int ch = lane; ch < chunks; ch += 32) {
    const int tap = ch / c8, cc = ch - tap * c8;
    const int th = tap / d.KW, tw = tap - th * d.KW;
    const int ih = i * d.in_stride - d.pad_h + th * d.dil, iw = j * d.in_stride - d.pad_w + tw * d.dil;
    if (ih < 0 || ih >= d.Hi || iw < 0 || iw >= d.Wi) continue;
    const uint4 xv = *reinterpret_cast<const uint4*>(x + (((size_t)n * d.Hi + ih) * d.Wi + iw) * d.Ci_total + d.ci_off + cc * 8);
    const uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
    for (int c = 0; c < CO; ++c) {
      const uint4 wv = *reinterpret_cast<const uint4*>(w + (size_t)c * K + (size_t)tap * d.Cin + cc * 8);
      const uint32_t ww[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        acc[c] = fmaf(__uint_as_float(xw[e] << 16), __uint_as_float(ww[e] << 16), acc[c]);
        acc[c] = fmaf(__uint_as_float(xw[e] & 0xffff0000u), __uint_as_float(ww[e] & 0xffff0000u), acc[c]);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < CO; ++c) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], off);
  }
  if (lane == 0) {
    const int oh = i * d.out_stride + d.out_off_h, ow = j * d.out_stride + d.out_off_w;
    __nv_bfloat16* dst = y + (((size_t)n * d.Ho + oh) * d.Wo + ow) * d.Co_total + d.co_off;
#pragma unroll
    for (int c = 0; c < CO; ++c)
      if (c < d.Cout) dst[c] = __float2bfloat16_rn(apply_act(acc[c] + (bias ? bias[c] : 0.f), d.act));
  }
}

bool conv2d_dot_supported(const msg_conv_desc* d, const void* x, const void* w) {
  if (d->dtype != MSG_BF16 || d->Cout > 2) return false;
  if (d->flags & (MSG_CONV_STATS | MSG_CONV_OUT_NCHW_F32 | MSG_CONV_ACCUM | MSG_CONV_PER_IMAGE_W | MSG_CONV_IN_NORM)) return false;
  if ((d->Cin | d->Ci_total | d->ci_off) & 7) return false;
  if (((uintptr_t)x | (uintptr_t)w) & 15) return false;
  return d->KH * d->KW * d->Cin >= 2048;
}

int conv2d_dot(const msg_conv_desc* d, const void* x, const void* w, const float* bias, void* y, cudaStream_t st) {
  const long long M = (long long)d->N * d->Hg * d->Wg;
  const unsigned grid = (unsigned)((M + 7) / 8);
  if (d->Cout == 1)
    conv_dot_kernel<1><<<grid, 256, 0, st>>>(*d, (const __nv_bfloat16*)x, (const __nv_bfloat16*)w, bias, (__nv_bfloat16*)y);
  else
    conv_dot_kernel<2><<<grid, 256, 0, st>>>(*d, (const __nv_bfloat16*)x, (const __nv_bfloat16*)w, bias, (__nv_bfloat16*)y);
  return check_launch("conv_dot_kernel");
}

int conv2d_simt(const msg_conv_desc* d, const void* x, const void* w, const float* bias, void* y,
                double* stats, const double* in_stats, cudaStream_t st) {
  const long long M = (long long)d->N * d->Hg * d->Wg;
  dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((d->Cout + BN - 1) / BN));
  const bool vec = (d->Cin % 4 == 0) && (d->Ci_total % 4 == 0) && (d->ci_off % 4 == 0);
  if (d->dtype == MSG_F32) {
    auto X = (const float*)x; auto W = (const float*)w;
    if (vec) conv_simt_kernel<float, true><<<grid, 256, 0, st>>>(*d, X, W, bias, y, stats, in_stats);
    else conv_simt_kernel<float, false><<<grid, 256, 0, st>>>(*d, X, W, bias, y, stats, in_stats);
  } else {
    auto X = (const __nv_bfloat16*)x; auto W = (const __nv_bfloat16*)w;
    if (vec) conv_simt_kernel<__nv_bfloat16, true><<<grid, 256, 0, st>>>(*d, X, W, bias, y, stats, in_stats);
    else conv_simt_kernel<__nv_bfloat16, false><<<grid, 256, 0, st>>>(*d, X, W, bias, y, stats, in_stats);
  }
  return check_launch("conv_simt_kernel");
}

int conv_validate(const msg_conv_desc* d) { return validate_desc(d); }

}  // namespace msg

namespace msg {
bool conv2d_wgrad_tc_supported(const msg_conv_desc* d, const void* x, const void* dy);
int conv2d_wgrad_tc(const msg_conv_desc* d, const void* x, const void* dy, float* dw, cudaStream_t st);
}  // namespace msg

using namespace msg;

extern "C" int msg_conv2d_wgrad(const msg_conv_desc* d, const void* x, const void* dy,
                                float* dw_packed, void* stream) {
  int rc = validate_desc(d);
  if (rc) return rc;
  MSG_REQUIRE(!(d->flags & MSG_CONV_OUT_NCHW_F32), MSG_ERR_UNSUPPORTED, "wgrad: NCHW dy unsupported");
  if (!(d->flags & MSG_CONV_FORCE_SIMT) && !((uintptr_t)dw_packed & 15) && conv2d_wgrad_tc_supported(d, x, dy))   // (bulk reduce: 16-byte rows)
    return conv2d_wgrad_tc(d, x, dy, dw_packed, as_stream(stream));      // tcgen05 kernel (conv_wgrad_tc.cu)
  const int K = d->KH * d->KW * d->Cin;
  const long long M = (long long)d->N * d->Hg * d->Wg;
  unsigned gx = (K + 63) / 64, gy = (d->Cout + 63) / 64;
  // enough splits to fill the machine ~4x, each a multiple of WP pixels
  long long want = (4LL * sm_count() + gx * gy - 1) / (gx * gy);
  long long max_splits = (M + WP - 1) / WP;
  if (want > max_splits) want = max_splits;
  if (want < 1) want = 1;
  if (want > 65535) want = 65535;
  long long per = (M + want - 1) / want;
  per = (per + WP - 1) / WP * WP;
  unsigned gz = (unsigned)((M + per - 1) / per);
  dim3 grid(gx, gy, gz);
  cudaStream_t st = as_stream(stream);
  const bool vec = (d->Cin % 4 == 0) && (d->Ci_total % 4 == 0) && (d->ci_off % 4 == 0);
  const double* in_stats = nullptr;
  MSG_REQUIRE(!(d->flags & MSG_CONV_IN_NORM), MSG_ERR_UNSUPPORTED, "wgrad: fused input norm unsupported");
  if (d->dtype == MSG_F32) {
    if (vec) conv_wgrad_simt_kernel<float, true><<<grid, 256, 0, st>>>(*d, (const float*)x, (const float*)dy, dw_packed, in_stats, per);
    else conv_wgrad_simt_kernel<float, false><<<grid, 256, 0, st>>>(*d, (const float*)x, (const float*)dy, dw_packed, in_stats, per);
  } else {
    if (vec) conv_wgrad_simt_kernel<__nv_bfloat16, true><<<grid, 256, 0, st>>>(*d, (const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, dw_packed, in_stats, per);
    else conv_wgrad_simt_kernel<__nv_bfloat16, false><<<grid, 256, 0, st>>>(*d, (const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, dw_packed, in_stats, per);
  }
  return check_launch("conv_wgrad_simt_kernel");
}

extern "C" int msg_pack_conv_weight(const float* w, int dim0, int dim1, int KH, int KW, int mode,
                                    int dtype, const float* denom, void* out, void* stream) {
  MSG_REQUIRE(mode >= 0 && mode <= 2, MSG_ERR_UNSUPPORTED, "pack: bad mode %d", mode);
  MSG_REQUIRE(mode != MSG_PACK_CONVT_PHASES || (KH == 4 && KW == 4), MSG_ERR_SHAPE, "pack: convT phases need a 4x4 kernel");
  long long total = (long long)dim0 * dim1 * KH * KW;
  unsigned blocks = (unsigned)((total + 255) / 256);
  if (dtype == MSG_F32) pack_weight_kernel<float><<<blocks, 256, 0, as_stream(stream)>>>(w, dim0, dim1, KH, KW, mode, denom, (float*)out, total);
  else if (dtype == MSG_BF16) pack_weight_kernel<__nv_bfloat16><<<blocks, 256, 0, as_stream(stream)>>>(w, dim0, dim1, KH, KW, mode, denom, (__nv_bfloat16*)out, total);
  else MSG_REQUIRE(false, MSG_ERR_UNSUPPORTED, "pack: bad dtype");
  return check_launch("pack_weight_kernel");
}

extern "C" int msg_unpack_conv_wgrad(const float* dw_packed, int dim0, int dim1, int KH, int KW,
                                     int mode, float* dw, void* stream) {
  MSG_REQUIRE(mode >= 0 && mode <= 2, MSG_ERR_UNSUPPORTED, "unpack: bad mode %d", mode);
  long long total = (long long)dim0 * dim1 * KH * KW;
  unsigned blocks = (unsigned)((total + 255) / 256);
  unpack_wgrad_kernel<<<blocks, 256, 0, as_stream(stream)>>>(dw_packed, dim0, dim1, KH, KW, mode, dw, total);
  return check_launch("unpack_wgrad_kernel");
}

extern "C" int msg_bias_grad(int dtype, const void* dy, long long rows, int C_total, int c_off,
                             int C, float* db, void* stream) {
  MSG_REQUIRE(rows > 0 && C > 0 && c_off >= 0 && c_off + C <= C_total, MSG_ERR_SHAPE, "bias_grad: bad shape");
  if (dtype == MSG_BF16 && ((C | c_off | C_total) & 7) == 0 && C <= 256 && ((C >> 3) & ((C >> 3) - 1)) == 0 &&
      (((uintptr_t)dy) & 15) == 0) {
    const int rpb = 256 / (C >> 3);
    long long g = (rows + (long long)rpb * 16 - 1) / ((long long)rpb * 16);    // ~16 rows per thread
    const long long cap2 = 8LL * sm_count();
    if (g > cap2) g = cap2;
    if (g < 1) g = 1;
    bias_grad_bf16_vec_kernel<<<(unsigned)g, 256, 0, as_stream(stream)>>>((const __nv_bfloat16*)dy, rows, C_total, c_off, C, db);
    return check_launch("bias_grad_bf16_vec_kernel");
  }
  unsigned gx = (C + 31) / 32;
  long long gy = (rows + 8 * 64 - 1) / (8 * 64);
  long long cap = (8LL * sm_count() + gx - 1) / gx;
  if (gy > cap) gy = cap;
  if (gy < 1) gy = 1;
  dim3 grid(gx, (unsigned)gy);
  if (dtype == MSG_F32) bias_grad_kernel<float><<<grid, 256, 0, as_stream(stream)>>>((const float*)dy, rows, C_total, c_off, C, db);
  else if (dtype == MSG_BF16) bias_grad_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>((const __nv_bfloat16*)dy, rows, C_total, c_off, C, db);
  else MSG_REQUIRE(false, MSG_ERR_UNSUPPORTED, "bias_grad: bad dtype");
  return check_launch("bias_grad_kernel");
}
