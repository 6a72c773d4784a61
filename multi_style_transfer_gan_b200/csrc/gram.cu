// Gram matrix G = F F^T / (C*H*W) of NHWC features and the style loss around it.
// north_star addition (the reference has no VGG / Gram / perceptual loss); oracle =
// oracle/restate.py::gram / style_loss (Gatys / Johnson formulation), parity unpinned vs reference.
//
// Features are [N][HW][C] (channel-contiguous), so F F^T is a "pixels are K" GEMM with both
// operands MN-major -- exactly the weight-gradient GEMM of a 1x1 conv with x = dy = F.  bf16 features with
// C % 64 == 0 and HW % 128 == 0 therefore run on the tcgen05 wgrad kernel (conv_wgrad_tc.cu) in its per-image
// mode: ONE launch produces the whole batch of Gram matrices straight from TMA boxes of F.  Everything else
// (fp32 parity mode, odd shapes) takes the SIMT engine below: 64x64 (a,b) tiles, split over pixels, fp32 atomics.
// Backward:  dF = coef * (G - A) . F  is a 1x1 convolution with a per-image weight (G - A): one launch of the
// TMA conv kernel with MSG_CONV_PER_IMAGE_W (a per-image loop over the generic conv path otherwise).
#include "common.cuh"

namespace msg {
int conv2d_dispatch(const msg_conv_desc* d, const void* x, const void* w, const float* bias, void* y,
                    double* stats, const double* in_stats, cudaStream_t st);
bool conv2d_tma_supported(const msg_conv_desc* d, const void* x, const void* w, const void* y);
bool conv2d_wgrad_tc_supported(const msg_conv_desc* d, const void* x, const void* dy);
int conv2d_wgrad_tc_per_image(const msg_conv_desc* d, const void* x, const void* dy, float* dw, cudaStream_t st);
namespace {

// [N][HW][C] features as the input AND output of a 1x1 "convolution" over a 1 x HW plane
msg_conv_desc plane_desc(int dtype, int N, long long HW, int C) {
  msg_conv_desc d = {};
  d.dtype = dtype; d.N = N; d.Hi = 1; d.Wi = (int)HW; d.Ci_total = C; d.ci_off = 0; d.Cin = C;
  d.Ho = 1; d.Wo = (int)HW; d.Co_total = C; d.co_off = 0; d.Cout = C; d.Hg = 1; d.Wg = (int)HW;
  d.KH = 1; d.KW = 1; d.in_stride = 1; d.pad_h = 0; d.pad_w = 0; d.dil = 1;
  d.out_stride = 1; d.out_off_h = 0; d.out_off_w = 0; d.act = MSG_ACT_NONE; d.flags = 0;
  return d;
}

__global__ void scale_kernel(float* __restrict__ g, long long n, float s) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    g[i] *= s;
}

constexpr int GP = 16;
template <typename T>
__global__ void __launch_bounds__(256)
gram_kernel(const T* __restrict__ f, long long HW, int C, float inv_norm, float* __restrict__ gram,
            long long pixels_per_split) {
  __shared__ __align__(16) float As[GP][68];
  __shared__ __align__(16) float Bs[GP][68];
  // grid: x = tile pair index (ta * tiles + tb), y = pixel split, z = image
  const int tiles = (C + 63) / 64;
  const int ta = blockIdx.x / tiles, tb = blockIdx.x % tiles;
  if (tb < ta) return;  // symmetric: compute the upper triangle of tiles, mirror on store
  const int img = blockIdx.z;
  const T* fi = f + (size_t)img * HW * C;
  long long p_begin = (long long)blockIdx.y * pixels_per_split;
  long long p_end = p_begin + pixels_per_split < HW ? p_begin + pixels_per_split : HW;
  const int tid = threadIdx.x, pl = tid >> 4, c4 = (tid & 15) * 4, ty = tid >> 4, tx = tid & 15;
  const bool vec = (C & 3) == 0;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (long long pc = p_begin; pc < p_end; pc += GP) {
    long long p = pc + pl;
    float a[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0};
    if (p < p_end) {
      const T* row = fi + (size_t)p * C;
      int ca = ta * 64 + c4, cb = tb * 64 + c4;
      if (vec && ca + 3 < C) load4(row + ca, a);
      else for (int e = 0; e < 4; ++e) if (ca + e < C) a[e] = to_f<T>(row[ca + e]);
      if (vec && cb + 3 < C) load4(row + cb, b);
      else for (int e = 0; e < 4; ++e) if (cb + e < C) b[e] = to_f<T>(row[cb + e]);
    }
    *reinterpret_cast<float4*>(&As[pl][c4]) = make_float4(a[0], a[1], a[2], a[3]);
    *reinterpret_cast<float4*>(&Bs[pl][c4]) = make_float4(b[0], b[1], b[2], b[3]);
    __syncthreads();
#pragma unroll
    for (int q = 0; q < GP; ++q) {
      float4 a4 = *reinterpret_cast<const float4*>(&As[q][ty * 4]);
      float4 b4 = *reinterpret_cast<const float4*>(&Bs[q][tx * 4]);
      float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* g = gram + (size_t)img * C * C;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int ra = ta * 64 + ty * 4 + i;
    if (ra >= C) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int cb = tb * 64 + tx * 4 + j;
      if (cb >= C) continue;
      float v = acc[i][j] * inv_norm;
      atomicAdd(g + (size_t)ra * C + cb, v);
      if (ta != tb) atomicAdd(g + (size_t)cb * C + ra, v);
    }
  }
}

// loss += scale * mean((g*gs - a)^2); with gs != 1 the normalised Gram matrices are also written back (the
// tensor-core GEMM leaves raw sums: its normalisation is fused into this pass)
__global__ void __launch_bounds__(256)
gram_mse_kernel(float* __restrict__ g, const float* __restrict__ a, long long n, float gs, float scale,
                float* __restrict__ loss) {
  __shared__ float red[8];
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float gv = g[i];
    if (gs != 1.f) { gv *= gs; g[i] = gv; }
    float d = gv - a[i];
    acc = fmaf(d, d, acc);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    atomicAdd(loss, t * scale / (float)n);
  }
}

template <typename T>
__global__ void gram_diff_weight_kernel(const float* __restrict__ g, const float* __restrict__ a,
                                        long long n, float coef, T* __restrict__ w) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    w[i] = from_f<T>(coef * (g[i] - a[i]));
}

// pending_scale: when non-null and the tensor-core path ran, the 1/(C*HW) normalisation is NOT applied and
// *pending_scale receives it (the caller fuses it into its next pass over the matrices); else it is set to 1.
template <typename T>
int gram_impl(const T* f, int N, long long HW, int C, float* gram, cudaStream_t st, float* pending_scale = nullptr) {
  if (pending_scale) *pending_scale = 1.f;
  cudaMemsetAsync(gram, 0, (size_t)N * C * C * sizeof(float), st);
  if (sizeof(T) == 2 && HW <= 0x7fffffffLL && HW % 128 == 0 && C % 64 == 0) {
    const msg_conv_desc d = plane_desc(MSG_BF16, N, HW, C);
    if (conv2d_wgrad_tc_supported(&d, f, f)) {            // tcgen05: G[n] = F[n]^T F[n] for the whole batch, one launch
      int rc = conv2d_wgrad_tc_per_image(&d, f, f, gram, st);
      if (rc) return rc;
      const float inv_norm = 1.f / ((float)C * (float)HW);
      if (pending_scale) { *pending_scale = inv_norm; return MSG_OK; }
      const long long n = (long long)N * C * C;
      scale_kernel<<<(unsigned)((n + 255) / 256 < 1024 ? (n + 255) / 256 : 1024), 256, 0, st>>>(gram, n, inv_norm);
      return check_launch("gram scale_kernel");
    }
  }
  int tiles = (C + 63) / 64;
  long long pairs = (long long)tiles * tiles;
  long long want = (4LL * sm_count()) / (pairs * N) + 1;
  long long maxs = (HW + GP - 1) / GP;
  if (want > maxs) want = maxs;
  long long per = ((HW + want - 1) / want + GP - 1) / GP * GP;
  unsigned gy = (unsigned)((HW + per - 1) / per);
  dim3 grid((unsigned)pairs, gy, N);
  float inv_norm = 1.f / ((float)C * (float)HW);
  gram_kernel<T><<<grid, 256, 0, st>>>(f, HW, C, inv_norm, gram, per);
  return check_launch("gram_kernel");
}

}  // namespace
}  // namespace msg

using namespace msg;

extern "C" int msg_gram(int dtype, const void* feat, int N, long long HW, int C, float* gram,
                        void* stream) {
  MSG_REQUIRE(N > 0 && HW > 0 && C > 0, MSG_ERR_SHAPE, "gram: bad shape");
  if (dtype == MSG_F32) return gram_impl<float>((const float*)feat, N, HW, C, gram, as_stream(stream));
  if (dtype == MSG_BF16) return gram_impl<__nv_bfloat16>((const __nv_bfloat16*)feat, N, HW, C, gram, as_stream(stream));
  MSG_REQUIRE(false, MSG_ERR_UNSUPPORTED, "gram: bad dtype");
}

extern "C" int msg_gram_loss_fwd(int dtype, const void* feat, int N, long long HW, int C,
                                 const float* target, float scale, float* gram, float* loss_out,
                                 void* stream) {
  MSG_REQUIRE(N > 0 && HW > 0 && C > 0, MSG_ERR_SHAPE, "gram: bad shape");
  float gs = 1.f;
  int rc;
  if (dtype == MSG_F32) rc = gram_impl<float>((const float*)feat, N, HW, C, gram, as_stream(stream), &gs);
  else if (dtype == MSG_BF16) rc = gram_impl<__nv_bfloat16>((const __nv_bfloat16*)feat, N, HW, C, gram, as_stream(stream), &gs);
  else MSG_REQUIRE(false, MSG_ERR_UNSUPPORTED, "gram: bad dtype");
  if (rc) return rc;
  long long n = (long long)N * C * C;
  unsigned blocks = (unsigned)((n + 255) / 256 < 1024 ? (n + 255) / 256 : 1024);
  gram_mse_kernel<<<blocks, 256, 0, as_stream(stream)>>>(gram, target, n, gs, scale, loss_out);
  return check_launch("gram_mse_kernel");
}

// dfeat[n][p][a] = coef * sum_b (G - A)[n][a][b] * F[n][p][b],  coef = 4*scale / (N*C*C * C*HW)
// wscratch: caller-owned [N][C][C] buffer of `dtype` that receives the per-image weights.
extern "C" int msg_gram_loss_bwd(int dtype, const void* feat, int N, long long HW, int C,
                                 const float* gram, const float* target, float scale,
                                 void* wscratch, void* dfeat, void* stream) {
  MSG_REQUIRE(N > 0 && HW > 0 && C > 0, MSG_ERR_SHAPE, "gram_bwd: bad shape");
  MSG_REQUIRE(HW <= 0x7fffffffLL, MSG_ERR_SHAPE, "gram_bwd: plane too large");
  MSG_REQUIRE(wscratch != nullptr, MSG_ERR_SHAPE, "gram_bwd: null scratch");
  cudaStream_t st = as_stream(stream);
  size_t esz = dtype == MSG_F32 ? 4 : 2;
  void* wbuf = wscratch;
  long long n = (long long)N * C * C;
  float coef = 4.f * scale / ((float)N * (float)C * (float)C * (float)C * (float)HW);
  unsigned blocks = (unsigned)((n + 255) / 256 < 2048 ? (n + 255) / 256 : 2048);
  if (dtype == MSG_F32) gram_diff_weight_kernel<float><<<blocks, 256, 0, st>>>(gram, target, n, coef, (float*)wbuf);
  else if (dtype == MSG_BF16) gram_diff_weight_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(gram, target, n, coef, (__nv_bfloat16*)wbuf);
  else MSG_REQUIRE(false, MSG_ERR_UNSUPPORTED, "gram_bwd: bad dtype");
  int rc = check_launch("gram_diff_weight_kernel");
  if (rc) return rc;
  // [HW x C] = F[HW x C] . W^T per image, W = coef*(G-A) [Cout=C][Cin=C]
  msg_conv_desc db = plane_desc(dtype, N, HW, C);
  db.flags = MSG_CONV_PER_IMAGE_W;
  if (dtype == MSG_BF16 && conv2d_tma_supported(&db, feat, wbuf, dfeat))
    return conv2d_dispatch(&db, feat, wbuf, nullptr, dfeat, nullptr, nullptr, st);     // whole batch, one launch
  msg_conv_desc d = plane_desc(dtype, 1, HW, C);
  for (int i = 0; i < N && rc == MSG_OK; ++i) {
    const char* fi = (const char*)feat + (size_t)i * HW * C * esz;
    char* di = (char*)dfeat + (size_t)i * HW * C * esz;
    const char* wi = (const char*)wbuf + (size_t)i * C * C * esz;
    rc = conv2d_dispatch(&d, fi, wi, nullptr, di, nullptr, nullptr, st);
  }
  return rc;
}
