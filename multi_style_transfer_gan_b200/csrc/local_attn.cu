// LocalAttention core (enhanced_generator.py:22-35), SIMT fp32 engine: forward and backward.
// Input is the [N,H,W,3C] qkv map produced by the 1x1 qkv conv (pointwise, so windowing before or
// after it is equivalent -- SURVEY.md section 4 invariant iii).  Per ws x ws window (P = ws*ws pixels):
//   qh = q / max(||q||_C, 1e-12) per pixel, kh likewise; S = qh kh^T (C x C); A = softmax_rows(S);
//   out = A v.
// |S| <= 1 because qh, kh are unit vectors, so exp() needs no max subtraction.  The C x C logits
// never leave the SM (the reference materialises them: [nW,C,C] is larger than the activation).
// One CTA per window (grid-stride), one warp per attention row.
#include "common.cuh"

namespace msg {
namespace {

constexpr int LA_TPB = 256;
constexpr int LA_WARPS = LA_TPB / 32;
constexpr int P = 16;  // pixels per window (ws = 4)
constexpr float kNormEps = 1e-12f;

struct WinCoord { int n, h0, w0; };
__device__ __forceinline__ WinCoord win_coord(long long wi, int H, int W) {
  int wpr = W / 4, wpi = (H / 4) * wpr;
  WinCoord c;
  c.n = (int)(wi / wpi);
  int r = (int)(wi - (long long)c.n * wpi);
  c.h0 = (r / wpr) * 4;
  c.w0 = (r % wpr) * 4;
  return c;
}

// loads q, k, v of one window into smem as float [P][C] each, normalises q and k in place and
// leaves 1/max(norm,eps) (and whether the clamp was active) in inv[2*P] / clamped[2*P].
template <typename T>
__device__ __forceinline__ void load_window(const T* __restrict__ qkv, const WinCoord& wc, int H,
                                            int W, int C, float* qh, float* kh, float* vs,
                                            float* inv, int* clamped) {
  const int tid = threadIdx.x;
  for (int idx = tid; idx < P * 3 * C; idx += LA_TPB) {
    int p = idx / (3 * C), c3 = idx - p * 3 * C;
    int hh = wc.h0 + (p >> 2), ww = wc.w0 + (p & 3);
    float v = to_f<T>(qkv[(((size_t)wc.n * H + hh) * W + ww) * 3 * C + c3]);
    if (c3 < C) qh[p * C + c3] = v;
    else if (c3 < 2 * C) kh[p * C + c3 - C] = v;
    else vs[p * C + c3 - 2 * C] = v;
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31;
  for (int pr = warp; pr < 2 * P; pr += LA_WARPS) {
    const float* src = (pr < P) ? qh + pr * C : kh + (pr - P) * C;
    float ss = 0.f;
    for (int c = lane; c < C; c += 32) ss = fmaf(src[c], src[c], ss);
    ss = warp_sum(ss);
    if (lane == 0) {
      float nrm = sqrtf(ss);
      clamped[pr] = nrm < kNormEps;
      inv[pr] = 1.f / fmaxf(nrm, kNormEps);
    }
  }
  __syncthreads();
  for (int idx = tid; idx < P * C; idx += LA_TPB) {
    int p = idx / C;
    qh[idx] *= inv[p];
    kh[idx] *= inv[P + p];
  }
  __syncthreads();
}

template <typename T>
__global__ void __launch_bounds__(LA_TPB)
local_attn_fwd_kernel(const T* __restrict__ qkv, int N, int H, int W, int C, T* __restrict__ out) {
  extern __shared__ float sm[];
  float* qh = sm;
  float* kh = qh + P * C;
  float* vs = kh + P * C;
  float* os = vs + P * C;
  float* inv = os + P * C;
  int* clamped = reinterpret_cast<int*>(inv + 2 * P);
  const long long nwin = (long long)N * (H / 4) * (W / 4);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (long long wi = blockIdx.x; wi < nwin; wi += gridDim.x) {
    WinCoord wc = win_coord(wi, H, W);
    load_window<T>(qkv, wc, H, W, C, qh, kh, vs, inv, clamped);
    for (int i = warp; i < C; i += LA_WARPS) {
      float qi[P];
#pragma unroll
      for (int p = 0; p < P; ++p) qi[p] = qh[p * C + i];
      float acc[P], den = 0.f;
#pragma unroll
      for (int p = 0; p < P; ++p) acc[p] = 0.f;
      for (int j = lane; j < C; j += 32) {
        float s = 0.f;
#pragma unroll
        for (int p = 0; p < P; ++p) s = fmaf(qi[p], kh[p * C + j], s);
        float e = expf(s);
        den += e;
#pragma unroll
        for (int p = 0; p < P; ++p) acc[p] = fmaf(e, vs[p * C + j], acc[p]);
      }
      den = warp_sum(den);
      float r = 1.f / den;
#pragma unroll
      for (int p = 0; p < P; ++p) {
        float a = warp_sum(acc[p]);
        if (lane == p) os[p * C + i] = a * r;
      }
    }
    __syncthreads();
    for (int idx = tid; idx < P * C; idx += LA_TPB) {
      int p = idx / C, c = idx - p * C;
      int hh = wc.h0 + (p >> 2), ww = wc.w0 + (p & 3);
      out[(((size_t)wc.n * H + hh) * W + ww) * C + c] = from_f<T>(os[idx]);
    }
    __syncthreads();
  }
}

template <typename T>
__global__ void __launch_bounds__(LA_TPB)
local_attn_bwd_kernel(const T* __restrict__ qkv, const T* __restrict__ dout, int N, int H, int W,
                      int C, T* __restrict__ dqkv) {
  extern __shared__ float sm[];
  float* qh = sm;
  float* kh = qh + P * C;
  float* vs = kh + P * C;
  float* dos = vs + P * C;
  float* dqh = dos + P * C;
  float* dkh = dqh + P * C;
  float* dvs = dkh + P * C;
  float* rowsum = dvs + P * C;
  float* rowdot = rowsum + C;
  float* inv = rowdot + C;
  float* dots = inv + 2 * P;
  int* clamped = reinterpret_cast<int*>(dots + 2 * P);
  const long long nwin = (long long)N * (H / 4) * (W / 4);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (long long wi = blockIdx.x; wi < nwin; wi += gridDim.x) {
    WinCoord wc = win_coord(wi, H, W);
    for (int idx = tid; idx < P * C; idx += LA_TPB) {
      int p = idx / C, c = idx - p * C;
      int hh = wc.h0 + (p >> 2), ww = wc.w0 + (p & 3);
      dos[idx] = to_f<T>(dout[(((size_t)wc.n * H + hh) * W + ww) * C + c]);
    }
    load_window<T>(qkv, wc, H, W, C, qh, kh, vs, inv, clamped);  // ends with __syncthreads
    // ---- row pass: softmax normalisers and d(qh)
    for (int i = warp; i < C; i += LA_WARPS) {
      float qi[P], di[P];
#pragma unroll
      for (int p = 0; p < P; ++p) { qi[p] = qh[p * C + i]; di[p] = dos[p * C + i]; }
      float den = 0.f, dot = 0.f;
      for (int j = lane; j < C; j += 32) {
        float s = 0.f, dp = 0.f;
#pragma unroll
        for (int p = 0; p < P; ++p) { s = fmaf(qi[p], kh[p * C + j], s); dp = fmaf(di[p], vs[p * C + j], dp); }
        float e = expf(s);
        den += e; dot = fmaf(e, dp, dot);
      }
      den = warp_sum(den); dot = warp_sum(dot);
      float r = 1.f / den;
      dot *= r;   // sum_j A_ij dP_ij
      if (lane == 0) { rowsum[i] = den; rowdot[i] = dot; }
      float acc[P];
#pragma unroll
      for (int p = 0; p < P; ++p) acc[p] = 0.f;
      for (int j = lane; j < C; j += 32) {
        float s = 0.f, dp = 0.f;
#pragma unroll
        for (int p = 0; p < P; ++p) { s = fmaf(qi[p], kh[p * C + j], s); dp = fmaf(di[p], vs[p * C + j], dp); }
        float ds = expf(s) * r * (dp - dot);
#pragma unroll
        for (int p = 0; p < P; ++p) acc[p] = fmaf(ds, kh[p * C + j], acc[p]);
      }
#pragma unroll
      for (int p = 0; p < P; ++p) {
        float a = warp_sum(acc[p]);
        if (lane == p) dqh[p * C + i] = a;
      }
    }
    __syncthreads();
    // ---- column pass: d(kh) and dv
    for (int j = warp; j < C; j += LA_WARPS) {
      float kj[P], vj[P], dk[P], dv[P];
#pragma unroll
      for (int p = 0; p < P; ++p) { kj[p] = kh[p * C + j]; vj[p] = vs[p * C + j]; dk[p] = 0.f; dv[p] = 0.f; }
      for (int i = lane; i < C; i += 32) {
        float s = 0.f, dp = 0.f;
#pragma unroll
        for (int p = 0; p < P; ++p) { s = fmaf(qh[p * C + i], kj[p], s); dp = fmaf(dos[p * C + i], vj[p], dp); }
        float a = expf(s) / rowsum[i];
        float ds = a * (dp - rowdot[i]);
#pragma unroll
        for (int p = 0; p < P; ++p) { dk[p] = fmaf(ds, qh[p * C + i], dk[p]); dv[p] = fmaf(a, dos[p * C + i], dv[p]); }
      }
#pragma unroll
      for (int p = 0; p < P; ++p) {
        float a = warp_sum(dk[p]), b = warp_sum(dv[p]);
        if (lane == p) { dkh[p * C + j] = a; dvs[p * C + j] = b; }
      }
    }
    __syncthreads();
    // ---- L2-normalise backward: dq = (dqh - qh <qh,dqh>) / ||q||   (or dqh/eps when clamped)
    for (int pr = warp; pr < 2 * P; pr += LA_WARPS) {
      const float* a = (pr < P) ? qh + pr * C : kh + (pr - P) * C;
      const float* b = (pr < P) ? dqh + pr * C : dkh + (pr - P) * C;
      float d = 0.f;
      for (int c = lane; c < C; c += 32) d = fmaf(a[c], b[c], d);
      d = warp_sum(d);
      if (lane == 0) dots[pr] = clamped[pr] ? 0.f : d;
    }
    __syncthreads();
    for (int idx = tid; idx < P * 3 * C; idx += LA_TPB) {
      int p = idx / (3 * C), c3 = idx - p * 3 * C;
      int hh = wc.h0 + (p >> 2), ww = wc.w0 + (p & 3);
      float v;
      if (c3 < C) v = (dqh[p * C + c3] - qh[p * C + c3] * dots[p]) * inv[p];
      else if (c3 < 2 * C) v = (dkh[p * C + c3 - C] - kh[p * C + c3 - C] * dots[P + p]) * inv[P + p];
      else v = dvs[p * C + c3 - 2 * C];
      dqkv[(((size_t)wc.n * H + hh) * W + ww) * 3 * C + c3] = from_f<T>(v);
    }
    __syncthreads();
  }
}

}  // namespace
}  // namespace msg

namespace msg {
bool local_attn_tc_supported(int dtype, int C, const void* qkv, const void* out);
int local_attn_fwd_tc(const void* qkv, int N, int H, int W, int C, void* out, cudaStream_t st);
bool local_attn_bwd_tc_supported(int dtype, int C, const void* qkv, const void* dout, const void* dqkv);
int local_attn_bwd_tc(const void* qkv, const void* dout, int N, int H, int W, int C, void* dqkv, cudaStream_t st);
}  // namespace msg

using namespace msg;

static unsigned la_grid(long long nwin) {
  long long g = 8LL * sm_count();
  return (unsigned)(nwin < g ? nwin : g);
}

extern "C" int msg_local_attn_fwd(int dtype, const void* qkv, int N, int H, int W, int C, int ws,
                                  void* out, void* stream) {
  MSG_REQUIRE(ws == 4, MSG_ERR_UNSUPPORTED, "local_attn: only window_size=4 (enhanced_generator.py:102)");
  MSG_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0 && H % 4 == 0 && W % 4 == 0, MSG_ERR_SHAPE,
              "local_attn: H, W must be multiples of the window size (got %dx%d)", H, W);
  cudaStream_t st = as_stream(stream);
  if (local_attn_tc_supported(dtype, C, qkv, out)) return local_attn_fwd_tc(qkv, N, H, W, C, out, st);
  size_t smem = (size_t)(4 * P * C + 2 * P) * sizeof(float) + 2 * P * sizeof(int);
  MSG_REQUIRE(smem <= 227 * 1024, MSG_ERR_UNSUPPORTED, "local_attn: C=%d too large", C);
  long long nwin = (long long)N * (H / 4) * (W / 4);
  if (dtype == MSG_F32) {
    cudaFuncSetAttribute(local_attn_fwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    local_attn_fwd_kernel<float><<<la_grid(nwin), LA_TPB, smem, st>>>((const float*)qkv, N, H, W, C, (float*)out);
  } else if (dtype == MSG_BF16) {
    cudaFuncSetAttribute(local_attn_fwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    local_attn_fwd_kernel<__nv_bfloat16><<<la_grid(nwin), LA_TPB, smem, st>>>((const __nv_bfloat16*)qkv, N, H, W, C, (__nv_bfloat16*)out);
  } else {
    MSG_REQUIRE(false, MSG_ERR_UNSUPPORTED, "local_attn: bad dtype");
  }
  return check_launch("local_attn_fwd_kernel");
}

extern "C" int msg_local_attn_bwd(int dtype, const void* qkv, const void* dout, int N, int H, int W,
                                  int C, int ws, void* dqkv, void* stream) {
  MSG_REQUIRE(ws == 4, MSG_ERR_UNSUPPORTED, "local_attn: only window_size=4");
  MSG_REQUIRE(N > 0 && C > 0 && H % 4 == 0 && W % 4 == 0, MSG_ERR_SHAPE, "local_attn_bwd: bad shape");
  if (local_attn_bwd_tc_supported(dtype, C, qkv, dout, dqkv))
    return local_attn_bwd_tc(qkv, dout, N, H, W, C, dqkv, as_stream(stream));
  size_t smem = (size_t)(7 * P * C + 2 * C + 4 * P) * sizeof(float) + 2 * P * sizeof(int);
  MSG_REQUIRE(smem <= 227 * 1024, MSG_ERR_UNSUPPORTED, "local_attn_bwd: C=%d too large", C);
  long long nwin = (long long)N * (H / 4) * (W / 4);
  cudaStream_t st = as_stream(stream);
  if (dtype == MSG_F32) {
    cudaFuncSetAttribute(local_attn_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    local_attn_bwd_kernel<float><<<la_grid(nwin), LA_TPB, smem, st>>>((const float*)qkv, (const float*)dout, N, H, W, C, (float*)dqkv);
  } else if (dtype == MSG_BF16) {
    cudaFuncSetAttribute(local_attn_bwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    local_attn_bwd_kernel<__nv_bfloat16><<<la_grid(nwin), LA_TPB, smem, st>>>((const __nv_bfloat16*)qkv, (const __nv_bfloat16*)dout, N, H, W, C, (__nv_bfloat16*)dqkv);
  } else {
    MSG_REQUIRE(false, MSG_ERR_UNSUPPORTED, "local_attn_bwd: bad dtype");
  }
  return check_launch("local_attn_bwd_kernel");
}
