// Bandwidth-bound helpers around the conv path: layout conversion at the module boundary,
// the multi-style output blend, losses, fused Adam, spectral-norm power iteration, pooling.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace msg {
namespace {

constexpr int EW_TPB = 256;
inline unsigned ew_blocks(long long n, int per_thread = 1) {
  long long b = (n + (long long)EW_TPB * per_thread - 1) / ((long long)EW_TPB * per_thread);
  long long cap = 32LL * sm_count();
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

// ---- layout ---------------------------------------------------------------------------------
template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, int N, int C, int H, int W, int Cp,
                                    T* __restrict__ y) {
  long long total = (long long)N * H * W * Cp;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int c = (int)(idx % Cp);
    long long p = idx / Cp;
    int w = (int)(p % W); p /= W;
    int h = (int)(p % H); int n = (int)(p / H);
    float v = c < C ? x[(((size_t)n * C + c) * H + h) * W + w] : 0.f;
    y[idx] = from_f<T>(v);
  }
}
// fast path: one padded pixel = ONE 16-byte vector (bf16: Cp = 8, fp32: Cp = 4).  One thread per pixel: C coalesced
// plane reads, one coalesced 16-byte store (the element-per-thread kernel above moved 117 MB in 0.18 ms = 0.64 TB/s).
template <typename T>
__global__ void nchw_to_nhwc_vec_kernel(const float* __restrict__ x, int N, int C, long long HW, T* __restrict__ y) {
  constexpr int Cp = 16 / (int)sizeof(T);
  const long long total = (long long)N * HW;
  for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < total;
       pix += (long long)gridDim.x * blockDim.x) {
    const long long n = pix / HW, p = pix - n * HW;
    const float* src = x + (size_t)n * C * HW + p;
    float v[Cp];
#pragma unroll
    for (int c = 0; c < Cp; ++c) v[c] = c < C ? src[(size_t)c * HW] : 0.f;
    if constexpr (sizeof(T) == 2) {
      float v8[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) v8[c] = v[c];
      *reinterpret_cast<uint4*>(y + pix * Cp) = pack8(v8);
    } else {
      *reinterpret_cast<float4*>(y + pix * Cp) = make_float4(v[0], v[1], v[2], v[3]);
    }
  }
}
template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ x, int N, int C, int H, int W, int Cp,
                                    float* __restrict__ y) {
  long long total = (long long)N * C * H * W;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    long long p = idx;
    int w = (int)(p % W); p /= W;
    int h = (int)(p % H); p /= H;
    int c = (int)(p % C); int n = (int)(p / C);
    y[idx] = to_f<T>(x[(((size_t)n * H + h) * W + w) * Cp + c]);
  }
}

// ---- blend ----------------------------------------------------------------------------------
constexpr int MAX_STYLES = 8;
struct BlendArgs {
  const float* ys[MAX_STYLES];
  float w[MAX_STYLES];
  int S;
};
__global__ void blend_kernel(BlendArgs a, const float* __restrict__ x, float w_x, float gain,
                             int do_clip, float lo, float hi, long long n,
                             float* __restrict__ out, uint8_t* __restrict__ out_u8) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float v = 0.f;
#pragma unroll
    for (int s = 0; s < MAX_STYLES; ++s)
      if (s < a.S) v = fmaf(a.w[s], a.ys[s][i], v);
    if (x) v = fmaf(w_x, x[i], v);
    v *= gain;
    if (do_clip) v = fminf(fmaxf(v, lo), hi);
    if (out) out[i] = v;
    if (out_u8) {
      float u = fminf(fmaxf((v + 1.f) * 0.5f, 0.f), 1.f) * 255.f;
      out_u8[i] = (uint8_t)u;  // truncation, as numpy .astype(uint8) in direct_transform.py:71
    }
  }
}

// ---- losses ---------------------------------------------------------------------------------
__device__ __forceinline__ void block_atomic_add(float v, float* dst) {
  __shared__ float red[EW_TPB / 32];
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < EW_TPB / 32; ++i) t += red[i];
    atomicAdd(dst, t);
  }
}
__global__ void __launch_bounds__(EW_TPB)
mse_kernel(const float* __restrict__ a, const float* __restrict__ b, float bc, long long n,
           float scale, float* __restrict__ loss, float* __restrict__ ga) {
  float acc = 0.f;
  const float inv_n = 1.f / (float)n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float d = a[i] - (b ? b[i] : bc);
    acc = fmaf(d, d, acc);
    if (ga) ga[i] = 2.f * d * inv_n * scale;
  }
  if (loss) block_atomic_add(acc * inv_n * scale, loss);
}
__global__ void __launch_bounds__(EW_TPB)
l1_kernel(const float* __restrict__ a, const float* __restrict__ b, float bc, long long n,
          float scale, float* __restrict__ loss, float* __restrict__ ga, float* __restrict__ gb) {
  float acc = 0.f;
  const float inv_n = 1.f / (float)n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float d = a[i] - (b ? b[i] : bc);
    acc += fabsf(d);
    float s = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);  // torch sign(): 0 at 0
    if (ga) ga[i] = s * inv_n * scale;
    if (gb) gb[i] = -s * inv_n * scale;
  }
  if (loss) block_atomic_add(acc * inv_n * scale, loss);
}

// ---- Adam -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(EW_TPB)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
            float* __restrict__ v, long long n, float lr, float b1, float b2, float eps,
            float bc1, float sqrt_bc2, float gscale) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float gi = g[i] * gscale;
    float mi = fmaf(b1, m[i], (1.f - b1) * gi);   // m.mul_(b1).add_(g, alpha=1-b1)
    float vi = fmaf(b2, v[i], (1.f - b2) * gi * gi);
    m[i] = mi; v[i] = vi;
    float denom = sqrtf(vi) / sqrt_bc2 + eps;
    p[i] -= (lr / bc1) * (mi / denom);
  }
}

// ---- spectral norm --------------------------------------------------------------------------
// One power iteration (old-style torch.nn.utils.spectral_norm, enhanced_generator.py:269-271):
//   v = normalize(W^T u); u = normalize(W v); sigma = u^T W v.
// The two passes over W (up to 512 x 4608 fp32 = 9 MB) used to run on ONE CTA (160 us per call, 9.6 ms per
// train step).  Now one thread-block CLUSTER of 8 CTAs shares the work: columns (pass 1) and rows (pass 2) are
// split over the 8192 threads, the three norms are reduced through distributed shared memory, and the two
// cluster barriers order the global v / u hand-over between the CTAs.
constexpr int SN_CLUSTER = 8;
constexpr int SN_TPB = 1024;
__device__ float block_sum_1024(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
  return t;
}
__device__ __forceinline__ float cluster_total(cg::cluster_group& cl, float* slot, float mine) {
  // every CTA publishes its partial in its own smem slot, then sums the 8 slots over DSMEM
  if (threadIdx.x == 0) *slot = mine;
  cl.sync();
  float t = 0.f;
  for (int r = 0; r < SN_CLUSTER; ++r) t += *cl.map_shared_rank(slot, r);
  return t;
}
__device__ __forceinline__ void spectral_norm_body(const float* __restrict__ w, int rows, int cols, float* __restrict__ u,
                                                   float* __restrict__ v, int do_iter, float eps, float* __restrict__ sigma) {
  __shared__ float red[32];
  __shared__ float part[3];
  cg::cluster_group cl = cg::this_cluster();
  const int tid = threadIdx.x, cta = (int)cl.block_rank();
  const int gtid = cta * SN_TPB + tid, gthreads = SN_CLUSTER * SN_TPB;
  const int warp = tid >> 5, lane = tid & 31;
  const int gwarp = cta * (SN_TPB / 32) + warp, gwarps = SN_CLUSTER * (SN_TPB / 32);
  float vscale = 1.f;                     // v is kept un-normalised in global memory until the end
  if (do_iter) {
    // pass 1: v_raw = W^T u (column j: sum_r w[r][j] u[r]) -- coalesced over j
    float ss = 0.f;
    for (int j = gtid; j < cols; j += gthreads) {
      float a = 0.f;                                          // (an 8-way unrolled variant measured slower: 3.9 vs 3.4 ms/step)
      for (int r = 0; r < rows; ++r) a = fmaf(w[(size_t)r * cols + j], u[r], a);
      v[j] = a;
      ss = fmaf(a, a, ss);
    }
    ss = block_sum_1024(ss, red);
    ss = cluster_total(cl, &part[0], ss);                     // barrier: every v_raw[j] and u read is done
    vscale = 1.f / fmaxf(sqrtf(ss), eps);
  }
  // pass 2: wv = W v : one warp per row; raw W v_raw goes to a register, scaled by 1/|v_raw|
  float ss2 = 0.f, dot = 0.f;
  for (int r = gwarp; r < rows; r += gwarps) {
    float a = 0.f;
    for (int j = lane; j < cols; j += 32) a = fmaf(w[(size_t)r * cols + j], v[j], a);
    a = warp_sum(a) * vscale;
    if (do_iter) {
      if (lane == 0) { u[r] = a; ss2 = fmaf(a, a, ss2); }     // u_raw = W v; the old u is no longer needed
    } else if (lane == 0) {
      dot = fmaf(u[r], a, dot);
    }
  }
  if (do_iter) {
    ss2 = block_sum_1024(ss2, red);
    ss2 = cluster_total(cl, &part[1], ss2);                   // barrier: all of v_raw has been consumed
    const float inv = 1.f / fmaxf(sqrtf(ss2), eps);
    for (int r = gwarp; r < rows; r += gwarps)                // each warp rescales the rows it produced
      if (lane == 0) u[r] *= inv;
    for (int j = gtid; j < cols; j += gthreads) v[j] *= vscale;   // each thread rescales the columns it produced
    if (gtid == 0) *sigma = ss2 * inv;                        // u^T (W v) = |W v|^2 / |W v|
  } else {
    dot = block_sum_1024(dot, red);
    dot = cluster_total(cl, &part[2], dot);
    if (gtid == 0) *sigma = dot;
  }
  cl.sync();                                                  // no CTA may exit while its smem slot can still be read
}
__global__ void __cluster_dims__(SN_CLUSTER, 1, 1) __launch_bounds__(SN_TPB)
spectral_norm_kernel(const float* __restrict__ w, int rows, int cols, float* __restrict__ u,
                     float* __restrict__ v, int do_iter, float eps, float* __restrict__ sigma) {
  spectral_norm_body(w, rows, cols, u, v, do_iter, eps, sigma);
}
// all spectral-normalised convs of one discriminator forward in ONE launch: cluster c works on problem c (7 launches of
// ~46 us each per forward, 70 per train step, were 3.3 ms of the step)
__global__ void __cluster_dims__(SN_CLUSTER, 1, 1) __launch_bounds__(SN_TPB)
spectral_norm_batched_kernel(const msg_sn_batch b, int do_iter, float eps) {
  const int i = (int)blockIdx.x / SN_CLUSTER;
  spectral_norm_body(b.w[i], b.rows[i], b.cols[i], b.u[i], b.v[i], do_iter, eps, b.sigma[i]);
}
// scratch[0] = <dw, w_orig>
__global__ void __launch_bounds__(EW_TPB)
dot_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n, float* __restrict__ out) {
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) acc = fmaf(a[i], b[i], acc);
  block_atomic_add(acc, out);
}
__global__ void __launch_bounds__(EW_TPB)
sn_bwd_kernel(const float* __restrict__ dw, const float* __restrict__ u, const float* __restrict__ v,
              const float* __restrict__ sigma, const float* __restrict__ dotp, int rows, int cols,
              float* __restrict__ dwo) {
  const float s = *sigma;
  const float coef = *dotp / (s * s);   // <dw, w_orig> / sigma^2
  long long n = (long long)rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    int r = (int)(i / cols), c = (int)(i - (long long)r * cols);
    dwo[i] += dw[i] / s - coef * u[r] * v[c];
  }
}

// ---- activation backward, add, pooling ------------------------------------------------------
template <typename T>
__global__ void act_bwd_kernel(const T* __restrict__ y, const T* __restrict__ dy, long long n, int act,
                               T* __restrict__ dx) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float yy = to_f<T>(y[i]), g = to_f<T>(dy[i]);
    float d;
    if (act == MSG_ACT_TANH) d = 1.f - yy * yy;
    else d = act_grad_from_pre(yy, act);   // relu / lrelu: sign(y) == sign(pre-activation)
    dx[i] = from_f<T>(g * d);
  }
}
template <typename T>
__global__ void tanh_bwd_nchw_kernel(const float* __restrict__ y, const float* __restrict__ dy, int N,
                                     int C, int H, int W, int Cp, T* __restrict__ dz) {
  long long total = (long long)N * H * W * Cp;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int c = (int)(idx % Cp);
    long long p = idx / Cp;
    int w = (int)(p % W); p /= W;
    int h = (int)(p % H); int n = (int)(p / H);
    float v = 0.f;
    if (c < C) {
      size_t s = (((size_t)n * C + c) * H + h) * W + w;
      v = dy[s] * (1.f - y[s] * y[s]);
    }
    dz[idx] = from_f<T>(v);
  }
}
template <typename T>
__global__ void add_kernel(const T* __restrict__ a, const T* __restrict__ b, long long n, T* __restrict__ o) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    o[i] = from_f<T>(to_f<T>(a[i]) + to_f<T>(b[i]));
}
template <typename T>
__global__ void __launch_bounds__(EW_TPB)
avgpool_fwd_kernel(const T* __restrict__ x, long long HW, int C, float* __restrict__ y) {
  // grid (C, N): plane mean
  __shared__ float red[EW_TPB / 32];
  const int c = blockIdx.x, n = blockIdx.y;
  float s = 0.f;
  for (long long p = threadIdx.x; p < HW; p += EW_TPB) s += to_f<T>(x[((size_t)n * HW + p) * C + c]);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < EW_TPB / 32; ++i) t += red[i];
    y[(size_t)n * C + c] = t / (float)HW;
  }
}
template <typename T>
__global__ void avgpool_bwd_kernel(const float* __restrict__ dy, int N, long long HW, int C, T* __restrict__ dx) {
  long long total = (long long)N * HW * C;
  const float inv = 1.f / (float)HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    int n = (int)(i / (HW * C));
    dx[i] = from_f<T>(dy[(size_t)n * C + c] * inv);
  }
}
template <typename T>
__global__ void maxpool_fwd_kernel(const T* __restrict__ x, int N, int H, int W, int C, T* __restrict__ y) {
  int Ho = H / 2, Wo = W / 2;
  long long total = (long long)N * Ho * Wo * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C); long long p = i / C;
    int ow = (int)(p % Wo); p /= Wo;
    int oh = (int)(p % Ho); int n = (int)(p / Ho);
    const T* b = x + (((size_t)n * H + 2 * oh) * W + 2 * ow) * C + c;
    float m = fmaxf(fmaxf(to_f<T>(b[0]), to_f<T>(b[C])), fmaxf(to_f<T>(b[(size_t)W * C]), to_f<T>(b[(size_t)W * C + C])));
    y[i] = from_f<T>(m);
  }
}
template <typename T>
__global__ void maxpool_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy, int N, int H, int W,
                                   int C, T* __restrict__ dx) {
  int Ho = H / 2, Wo = W / 2;
  long long total = (long long)N * Ho * Wo * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C); long long p = i / C;
    int ow = (int)(p % Wo); p /= Wo;
    int oh = (int)(p % Ho); int n = (int)(p / Ho);
    size_t b = (((size_t)n * H + 2 * oh) * W + 2 * ow) * C + c;
    size_t offs[4] = {b, b + C, b + (size_t)W * C, b + (size_t)W * C + C};
    int best = 0; float m = to_f<T>(x[offs[0]]);
#pragma unroll
    for (int k = 1; k < 4; ++k) { float v = to_f<T>(x[offs[k]]); if (v > m) { m = v; best = k; } }  // first max wins (PyTorch)
    T g = dy[i];
#pragma unroll
    for (int k = 0; k < 4; ++k) dx[offs[k]] = (k == best) ? g : from_f<T>(0.f);
  }
}

}  // namespace
}  // namespace msg

using namespace msg;

#define DISPATCH_DTYPE(dtype, NAME, ...)                                  \
  if ((dtype) == MSG_F32) { using T = float; __VA_ARGS__; }               \
  else if ((dtype) == MSG_BF16) { using T = __nv_bfloat16; __VA_ARGS__; } \
  else MSG_REQUIRE(false, MSG_ERR_UNSUPPORTED, NAME ": bad dtype");

extern "C" int msg_nchw_to_nhwc(int dtype, const float* x, int N, int C, int H, int W, int Cp,
                                void* y, void* stream) {
  MSG_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0 && Cp >= C, MSG_ERR_SHAPE, "nchw_to_nhwc: bad shape");
  long long total = (long long)N * H * W * Cp;
  if (((dtype == MSG_BF16 && Cp == 8) || (dtype == MSG_F32 && Cp == 4)) && (((uintptr_t)y) & 15) == 0) {
    const long long pixels = (long long)N * H * W;
    if (dtype == MSG_BF16)
      nchw_to_nhwc_vec_kernel<__nv_bfloat16><<<ew_blocks(pixels), EW_TPB, 0, as_stream(stream)>>>(x, N, C, (long long)H * W, (__nv_bfloat16*)y);
    else
      nchw_to_nhwc_vec_kernel<float><<<ew_blocks(pixels), EW_TPB, 0, as_stream(stream)>>>(x, N, C, (long long)H * W, (float*)y);
    return check_launch("nchw_to_nhwc_vec_kernel");
  }
  DISPATCH_DTYPE(dtype, "nchw_to_nhwc", (nchw_to_nhwc_kernel<T><<<ew_blocks(total), EW_TPB, 0, as_stream(stream)>>>(x, N, C, H, W, Cp, (T*)y)));
  return check_launch("nchw_to_nhwc_kernel");
}
extern "C" int msg_nhwc_to_nchw(int dtype, const void* x, int N, int C, int H, int W, int Cp,
                                float* y, void* stream) {
  MSG_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0 && Cp >= C, MSG_ERR_SHAPE, "nhwc_to_nchw: bad shape");
  long long total = (long long)N * H * W * C;
  DISPATCH_DTYPE(dtype, "nhwc_to_nchw", (nhwc_to_nchw_kernel<T><<<ew_blocks(total), EW_TPB, 0, as_stream(stream)>>>((const T*)x, N, C, H, W, Cp, y)));
  return check_launch("nhwc_to_nchw_kernel");
}

extern "C" int msg_blend_outputs(const float* const* ys, const float* w, int S, const float* x,
                                 float w_x, float gain, int do_clip, float clip_lo, float clip_hi,
                                 long long numel, float* out_f32, uint8_t* out_u8, void* stream) {
  MSG_REQUIRE(S >= 1 && S <= MAX_STYLES, MSG_ERR_SHAPE, "blend: S must be in [1,%d]", MAX_STYLES);
  MSG_REQUIRE(numel > 0 && (out_f32 || out_u8), MSG_ERR_SHAPE, "blend: nothing to do");
  BlendArgs a;
  a.S = S;
  for (int s = 0; s < MAX_STYLES; ++s) { a.ys[s] = s < S ? ys[s] : nullptr; a.w[s] = s < S ? w[s] : 0.f; }
  blend_kernel<<<ew_blocks(numel), EW_TPB, 0, as_stream(stream)>>>(a, x, w_x, gain, do_clip, clip_lo, clip_hi, numel, out_f32, out_u8);
  return check_launch("blend_kernel");
}

extern "C" int msg_mse_loss(const float* a, const float* b, float b_const, long long n, float scale,
                            float* loss_out, float* grad_a, void* stream) {
  MSG_REQUIRE(n > 0, MSG_ERR_SHAPE, "mse: empty");
  mse_kernel<<<ew_blocks(n, 4), EW_TPB, 0, as_stream(stream)>>>(a, b, b_const, n, scale, loss_out, grad_a);
  return check_launch("mse_kernel");
}
extern "C" int msg_l1_loss(const float* a, const float* b, float b_const, long long n, float scale,
                           float* loss_out, float* grad_a, float* grad_b, void* stream) {
  MSG_REQUIRE(n > 0, MSG_ERR_SHAPE, "l1: empty");
  l1_kernel<<<ew_blocks(n, 4), EW_TPB, 0, as_stream(stream)>>>(a, b, b_const, n, scale, loss_out, grad_a, grad_b);
  return check_launch("l1_kernel");
}

// as adam_kernel with the step count read from device memory (a captured CUDA graph replays the same launch every step, so
// the bias corrections cannot be launch arguments); *step is incremented by adam_step_incr_kernel in front of it
__global__ void adam_step_incr_kernel(int* step) { *step += 1; }
__global__ void __launch_bounds__(EW_TPB)
adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                float* __restrict__ v, long long n, float lr, float b1, float b2, float eps,
                const int* __restrict__ step, float gscale) {
  const int t = *step;
  const float bc1 = (float)(1.0 - pow((double)b1, (double)t));
  const float sqrt_bc2 = (float)sqrt(1.0 - pow((double)b2, (double)t));
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float gi = g[i] * gscale;
    float mi = fmaf(b1, m[i], (1.f - b1) * gi);
    float vi = fmaf(b2, v[i], (1.f - b2) * gi * gi);
    m[i] = mi; v[i] = vi;
    float denom = sqrtf(vi) / sqrt_bc2 + eps;
    p[i] -= (lr / bc1) * (mi / denom);
  }
}

// ---- uint8 pre / post-processing on the device (batch_process_images.py:193-205, 287-291, 304-310) ----------------------
// canvas paste + ToTensor + Normalize(0.5, 0.5): out[n][c][Y][X] = (v / 255 - 0.5) / 0.5 with v = img[n][Y-oy][X-ox][c] inside
// the pasted image and `fill` (the white canvas) outside.  torchvision's arithmetic: ToTensor divides by 255 in fp32, Normalize
// subtracts 0.5 and divides by 0.5 in fp32.
__global__ void __launch_bounds__(EW_TPB)
u8_canvas_to_nchw_kernel(const uint8_t* __restrict__ img, int N, int h, int w, int H, int W, int oy, int ox, int fill,
                         float* __restrict__ out, uint8_t* __restrict__ canvas) {
  const long long total = (long long)N * H * W;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int X = (int)(idx % W);
    long long t = idx / W;
    const int Y = (int)(t % H), n = (int)(t / H);
    const int y = Y - oy, x = X - ox;
    int v[3] = {fill, fill, fill};
    if (y >= 0 && y < h && x >= 0 && x < w) {
      const uint8_t* s = img + (((size_t)n * h + y) * w + x) * 3;
      v[0] = s[0]; v[1] = s[1]; v[2] = s[2];
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float f = __fdiv_rn((float)v[c], 255.f);
      out[(((size_t)n * 3 + c) * H + Y) * W + X] = __fdiv_rn(f - 0.5f, 0.5f);
      if (canvas) canvas[(((size_t)n * H + Y) * W + X) * 3 + c] = (uint8_t)v[c];
    }
  }
}
// "simple" mode strength blend in uint8 space: out = uint8(clip(orig * (1 - s) + styled * s, 0, 255)) -- numpy's float64
// arithmetic and truncating astype, batch_process_images.py:304-310.  orig / out: NHWC uint8 (PIL layout); styled: NCHW uint8.
__global__ void __launch_bounds__(EW_TPB)
u8_strength_blend_kernel(const uint8_t* __restrict__ orig, const uint8_t* __restrict__ styled, int N, int H, int W, double s,
                         uint8_t* __restrict__ out) {
  const long long total = (long long)N * H * W;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int X = (int)(idx % W);
    long long t = idx / W;
    const int Y = (int)(t % H), n = (int)(t / H);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const size_t o = (size_t)idx * 3 + c;
      const double st = (double)styled[(((size_t)n * 3 + c) * H + Y) * W + X];
      double r = __dadd_rn(__dmul_rn((double)orig[o], 1.0 - s), __dmul_rn(st, s));     // (no fma contraction: numpy rounds each product)
      r = r < 0.0 ? 0.0 : (r > 255.0 ? 255.0 : r);
      out[o] = (uint8_t)r;
    }
  }
}

extern "C" int msg_u8_canvas_to_nchw(const uint8_t* img, int N, int h, int w, int H, int W, int off_y, int off_x, int fill,
                                     float* out, uint8_t* canvas, void* stream) {
  MSG_REQUIRE(img && out && N > 0 && h > 0 && w > 0 && H > 0 && W > 0, MSG_ERR_SHAPE, "u8_canvas_to_nchw: bad arguments");
  MSG_REQUIRE(off_y >= 0 && off_x >= 0 && off_y + h <= H && off_x + w <= W && fill >= 0 && fill <= 255, MSG_ERR_SHAPE,
              "u8_canvas_to_nchw: the %dx%d image at (%d, %d) does not fit the %dx%d canvas", h, w, off_y, off_x, H, W);
  const long long total = (long long)N * H * W;
  u8_canvas_to_nchw_kernel<<<ew_blocks(total, 2), EW_TPB, 0, as_stream(stream)>>>(img, N, h, w, H, W, off_y, off_x, fill, out, canvas);
  return check_launch("u8_canvas_to_nchw_kernel");
}

extern "C" int msg_u8_strength_blend(const uint8_t* orig_nhwc, const uint8_t* styled_nchw, int N, int H, int W, double strength,
                                     uint8_t* out_nhwc, void* stream) {
  MSG_REQUIRE(orig_nhwc && styled_nchw && out_nhwc && N > 0 && H > 0 && W > 0, MSG_ERR_SHAPE, "u8_strength_blend: bad arguments");
  const long long total = (long long)N * H * W;
  u8_strength_blend_kernel<<<ew_blocks(total, 2), EW_TPB, 0, as_stream(stream)>>>(orig_nhwc, styled_nchw, N, H, W, strength, out_nhwc);
  return check_launch("u8_strength_blend_kernel");
}

extern "C" int msg_adam_step_dev(float* p, const float* g, float* m, float* v, long long n, float lr,
                                 float beta1, float beta2, float eps, int* step_dev, float grad_scale,
                                 void* stream) {
  MSG_REQUIRE(n > 0 && step_dev != nullptr, MSG_ERR_SHAPE, "adam: bad arguments");
  adam_step_incr_kernel<<<1, 1, 0, as_stream(stream)>>>(step_dev);
  adam_dev_kernel<<<ew_blocks(n, 2), EW_TPB, 0, as_stream(stream)>>>(p, g, m, v, n, lr, beta1, beta2, eps, step_dev, grad_scale);
  return check_launch("adam_dev_kernel");
}

extern "C" int msg_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr,
                             float beta1, float beta2, float eps, int step, float grad_scale,
                             void* stream) {
  MSG_REQUIRE(n > 0 && step >= 1, MSG_ERR_SHAPE, "adam: bad arguments");
  double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
  adam_kernel<<<ew_blocks(n, 2), EW_TPB, 0, as_stream(stream)>>>(p, g, m, v, n, lr, beta1, beta2, eps, (float)bc1, (float)sqrt(bc2), grad_scale);
  return check_launch("adam_kernel");
}

extern "C" int msg_spectral_norm(const float* w, int rows, int cols, float* u, float* v,
                                 int do_power_iter, float eps, float* sigma, void* stream) {
  MSG_REQUIRE(rows > 0 && cols > 0, MSG_ERR_SHAPE, "spectral_norm: bad shape");
  spectral_norm_kernel<<<SN_CLUSTER, SN_TPB, 0, as_stream(stream)>>>(w, rows, cols, u, v, do_power_iter, eps, sigma);
  return check_launch("spectral_norm_kernel");
}
extern "C" int msg_spectral_norm_batched(const msg_sn_batch* b, int do_power_iter, float eps, void* stream) {
  MSG_REQUIRE(b != nullptr && b->n >= 1 && b->n <= MSG_SN_MAX_BATCH, MSG_ERR_SHAPE, "spectral_norm_batched: 1..%d problems", MSG_SN_MAX_BATCH);
  for (int i = 0; i < b->n; ++i)
    MSG_REQUIRE(b->rows[i] > 0 && b->cols[i] > 0 && b->w[i] && b->u[i] && b->v[i] && b->sigma[i], MSG_ERR_SHAPE,
                "spectral_norm_batched: bad problem %d", i);
  spectral_norm_batched_kernel<<<SN_CLUSTER * b->n, SN_TPB, 0, as_stream(stream)>>>(*b, do_power_iter, eps);
  return check_launch("spectral_norm_batched_kernel");
}

extern "C" int msg_spectral_norm_bwd(const float* dw, const float* w_orig, const float* u,
                                     const float* v, const float* sigma, int rows, int cols,
                                     float* dw_orig, float* scratch, void* stream) {
  long long n = (long long)rows * cols;
  cudaStream_t st = as_stream(stream);
  cudaMemsetAsync(scratch, 0, sizeof(float), st);
  dot_kernel<<<ew_blocks(n, 4), EW_TPB, 0, st>>>(dw, w_orig, n, scratch);
  sn_bwd_kernel<<<ew_blocks(n), EW_TPB, 0, st>>>(dw, u, v, sigma, scratch, rows, cols, dw_orig);
  return check_launch("spectral_norm_bwd");
}

extern "C" int msg_act_bwd(int dtype, const void* y, const void* dy, long long n, int act, void* dx,
                           void* stream) {
  MSG_REQUIRE(n > 0, MSG_ERR_SHAPE, "act_bwd: empty");
  DISPATCH_DTYPE(dtype, "act_bwd", (act_bwd_kernel<T><<<ew_blocks(n, 2), EW_TPB, 0, as_stream(stream)>>>((const T*)y, (const T*)dy, n, act, (T*)dx)));
  return check_launch("act_bwd_kernel");
}
extern "C" int msg_tanh_bwd_nchw(int dtype, const float* y, const float* dy, int N, int C, int H,
                                 int W, int Cp, void* dz, void* stream) {
  long long total = (long long)N * H * W * Cp;
  MSG_REQUIRE(total > 0 && Cp >= C, MSG_ERR_SHAPE, "tanh_bwd: bad shape");
  DISPATCH_DTYPE(dtype, "tanh_bwd", (tanh_bwd_nchw_kernel<T><<<ew_blocks(total), EW_TPB, 0, as_stream(stream)>>>(y, dy, N, C, H, W, Cp, (T*)dz)));
  return check_launch("tanh_bwd_nchw_kernel");
}
extern "C" int msg_add(int dtype, const void* a, const void* b, long long n, void* out, void* stream) {
  MSG_REQUIRE(n > 0, MSG_ERR_SHAPE, "add: empty");
  DISPATCH_DTYPE(dtype, "add", (add_kernel<T><<<ew_blocks(n, 2), EW_TPB, 0, as_stream(stream)>>>((const T*)a, (const T*)b, n, (T*)out)));
  return check_launch("add_kernel");
}
extern "C" int msg_avgpool_fwd(int dtype, const void* x, int N, long long HW, int C, float* y, void* stream) {
  MSG_REQUIRE(N > 0 && HW > 0 && C > 0, MSG_ERR_SHAPE, "avgpool: bad shape");
  dim3 grid(C, N);
  DISPATCH_DTYPE(dtype, "avgpool_fwd", (avgpool_fwd_kernel<T><<<grid, EW_TPB, 0, as_stream(stream)>>>((const T*)x, HW, C, y)));
  return check_launch("avgpool_fwd_kernel");
}
extern "C" int msg_avgpool_bwd(int dtype, const float* dy, int N, long long HW, int C, void* dx, void* stream) {
  MSG_REQUIRE(N > 0 && HW > 0 && C > 0, MSG_ERR_SHAPE, "avgpool: bad shape");
  long long total = (long long)N * HW * C;
  DISPATCH_DTYPE(dtype, "avgpool_bwd", (avgpool_bwd_kernel<T><<<ew_blocks(total), EW_TPB, 0, as_stream(stream)>>>(dy, N, HW, C, (T*)dx)));
  return check_launch("avgpool_bwd_kernel");
}
extern "C" int msg_maxpool2x2_fwd(int dtype, const void* x, int N, int H, int W, int C, void* y, void* stream) {
  MSG_REQUIRE(N > 0 && H % 2 == 0 && W % 2 == 0 && C > 0, MSG_ERR_SHAPE, "maxpool: bad shape");
  long long total = (long long)N * (H / 2) * (W / 2) * C;
  DISPATCH_DTYPE(dtype, "maxpool_fwd", (maxpool_fwd_kernel<T><<<ew_blocks(total), EW_TPB, 0, as_stream(stream)>>>((const T*)x, N, H, W, C, (T*)y)));
  return check_launch("maxpool_fwd_kernel");
}
extern "C" int msg_maxpool2x2_bwd(int dtype, const void* x, const void* dy, int N, int H, int W, int C,
                                  void* dx, void* stream) {
  MSG_REQUIRE(N > 0 && H % 2 == 0 && W % 2 == 0 && C > 0, MSG_ERR_SHAPE, "maxpool: bad shape");
  long long total = (long long)N * (H / 2) * (W / 2) * C;
  DISPATCH_DTYPE(dtype, "maxpool_bwd", (maxpool_bwd_kernel<T><<<ew_blocks(total), EW_TPB, 0, as_stream(stream)>>>((const T*)x, (const T*)dy, N, H, W, C, (T*)dx)));
  return check_launch("maxpool_bwd_kernel");
}
