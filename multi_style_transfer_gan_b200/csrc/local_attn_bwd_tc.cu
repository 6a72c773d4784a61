// LocalAttention core backward (autograd of enhanced_generator.py:22-35), bf16 tensor-core kernel.
//
// Per 4x4 window (16 pixels p, C channels; i = query channel, j = key channel):
//   qh = q/|q|, kh = k/|k| (over channels, per pixel)       S[i][j] = sum_p qh[p][i] kh[p][j]
//   A = softmax_j(S)                                         out[p][i] = sum_j A[i][j] v[p][j]
// and with dO = d out:
//   dA[i][j] = sum_p dO[p][i] v[p][j]        D[i] = sum_j A[i][j] dA[i][j]      dS = A (dA - D)
//   dv[p][j]  = sum_i dO[p][i] A[i][j]       dqh[p][i] = sum_j dS[i][j] kh[p][j]
//   dkh[p][j] = sum_i dS[i][j] qh[p][i]      dq = (dqh - qh <qh,dqh>) / |q|     (same for k)
// One CTA (4 warps) per window.  Everything C x C lives in mma.sync accumulator fragments:
//   phase R: each warp owns 16-row slabs of S (all j): softmax sums, D, dS, and dqh (contraction over j runs
//            along the accumulator's column pairs, which is exactly the B-fragment layout of dS^T);
//   phase C: each warp owns 16-column slabs, recomputed TRANSPOSED (S^T = Kh^T Qh with the row sums and D of
//            phase R read from smem), so the contraction over i for dv and dkh again runs along column pairs.
// S and dA are recomputed where needed (K = 16 pixels: one mma per 16x8 tile) instead of being stored.
// exp is a packed bf16 ex2 (log2(e) is folded into the stored qh), identical in both phases.
#include "common.cuh"
#include "la_mma.cuh"

namespace msg {
namespace {
using namespace la;

constexpr int LB_THREADS = 128;
constexpr int P16 = 16;
constexpr float LOG2E = 1.4426950408889634f;

template <int C>
struct BwdCfg {
  static constexpr int PITCH = 2 * C + 16;          // bytes per pixel row of q / k / v / dO
  static constexpr int MAT = P16 * PITCH;
  static constexpr int OPITCH = 6 * C + 16;         // bytes per pixel row of the dq | dk | dv staging tile
  static constexpr int JW = C < 64 ? C : 64;        // columns per register chunk
  static constexpr int NCHK = C / JW;
  static constexpr int NSLAB = C / 16;
  static constexpr int SPW = (NSLAB + 3) / 4;       // slabs per warp
  static constexpr int FLOATS = 2 * C + 4 * P16;    // rinv, rdot, invq, invk, dotq, dotk
  static constexpr size_t SMEM = 4 * MAT + P16 * OPITCH + FLOATS * 4 + 2 * P16 * 4;
};

// acc[nt][4] (16 rows x JW cols) = slab-A (rows = channels c0.., K = 16 pixels) x chunk of Y (cols cb..cb+JW)
template <int C>
__device__ __forceinline__ void slab_times_chunk(float (&acc)[BwdCfg<C>::JW / 8][4], const uint32_t (&afr)[4],
                                                 uint32_t y_a, int cb, int r8, int mi) {
  constexpr int PITCH = BwdCfg<C>::PITCH, JW = BwdCfg<C>::JW;
#pragma unroll
  for (int nt = 0; nt < JW / 8; nt += 2) {
    acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
    acc[nt + 1][0] = acc[nt + 1][1] = acc[nt + 1][2] = acc[nt + 1][3] = 0.f;
    uint32_t bfr[4];
    ldsm_x4_t(y_a + (r8 + 8 * (mi & 1)) * PITCH + (cb + 8 * (nt + (mi >> 1))) * 2, bfr);
    mma_bf16(acc[nt], afr, bfr[0], bfr[1]);
    mma_bf16(acc[nt + 1], afr, bfr[2], bfr[3]);
  }
}

template <int C>
__global__ void __launch_bounds__(LB_THREADS, (C <= 64 ? 4 : C == 128 ? 3 : 2))
local_attn_bwd_tc_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ dout,
                         int N, int H, int W, __nv_bfloat16* __restrict__ dqkv) {
  using Cfg = BwdCfg<C>;
  constexpr int PITCH = Cfg::PITCH, MAT = Cfg::MAT, OPITCH = Cfg::OPITCH, JW = Cfg::JW, NCHK = Cfg::NCHK;
  constexpr int NSLAB = Cfg::NSLAB, SPW = Cfg::SPW, CH = C / 8;
  extern __shared__ __align__(16) uint8_t sm[];
  uint8_t* qs = sm;
  uint8_t* ks = qs + MAT;
  uint8_t* vs = ks + MAT;
  uint8_t* ds = vs + MAT;                           // dO
  uint8_t* ot = ds + MAT;                           // [16][OPITCH] staged dq | dk | dv
  float* rinv = reinterpret_cast<float*>(ot + P16 * OPITCH);
  float* rdot = rinv + C;
  float* invq = rdot + C;
  float* invk = invq + P16;
  float* dotq = invk + P16;
  float* dotk = dotq + P16;
  int* clampq = reinterpret_cast<int*>(dotk + P16);
  int* clampk = clampq + P16;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mi = lane >> 3, r8 = lane & 7, g = lane >> 2, q4 = lane & 3;
  const uint32_t qs_a = s_u32(qs), ks_a = s_u32(ks), vs_a = s_u32(vs), ds_a = s_u32(ds);
  const uint32_t ONES = 0x3F803F80u;
  const int wpr = W / 4, wpi = (H / 4) * wpr;
  const long long nwin = (long long)N * wpi;

  for (long long wi = blockIdx.x; wi < nwin; wi += gridDim.x) {
    const int n = (int)(wi / wpi);
    const int rw = (int)(wi - (long long)n * wpi);
    const int h0 = (rw / wpr) * 4, w0 = (rw % wpr) * 4;
    const size_t pix0 = ((size_t)n * H + h0) * W + w0;
    // ---- stage q, k, v, dO: 16 pixels x (3C + C) bf16
    for (int c = tid; c < P16 * 4 * CH; c += LB_THREADS) {
      const int p = c / (4 * CH), cc = c - p * (4 * CH);
      const int part = cc / CH, off = cc - part * CH;
      const size_t pix = pix0 + (size_t)(p >> 2) * W + (p & 3);
      const __nv_bfloat16* src = part < 3 ? qkv + pix * (3 * C) + cc * 8 : dout + pix * C + off * 8;
      cpa16(qs_a + part * MAT + p * PITCH + off * 16, src);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    if (tid < P16) { dotq[tid] = 0.f; dotk[tid] = 0.f; }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    // ---- L2-normalise q and k over C per pixel, in place: qs <- log2(e) q/|q|, ks <- k/|k|
    {
      const int p = tid >> 3, part = tid & 7;
      constexpr int VPT = (C + 63) / 64;
      const bool act = part * 8 < C;                 // C = 32: only 4 chunks per pixel
      uint4* qp = reinterpret_cast<uint4*>(qs + p * PITCH) + part;
      uint4* kp = reinterpret_cast<uint4*>(ks + p * PITCH) + part;
      uint4 qv[VPT], kv[VPT];
      float sq = 0.f, sk = 0.f;
#pragma unroll
      for (int v = 0; v < VPT; ++v) {
        qv[v] = act ? qp[8 * v] : make_uint4(0, 0, 0, 0);
        kv[v] = act ? kp[8 * v] : make_uint4(0, 0, 0, 0);
        const uint32_t qw[4] = {qv[v].x, qv[v].y, qv[v].z, qv[v].w}, kw[4] = {kv[v].x, kv[v].y, kv[v].z, kv[v].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          sq = fmaf(bf_lo(qw[e]), bf_lo(qw[e]), fmaf(bf_hi(qw[e]), bf_hi(qw[e]), sq));
          sk = fmaf(bf_lo(kw[e]), bf_lo(kw[e]), fmaf(bf_hi(kw[e]), bf_hi(kw[e]), sk));
        }
      }
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
        sk += __shfl_xor_sync(0xffffffffu, sk, o);
      }
      const float nq = sqrtf(sq), nk = sqrtf(sk);
      const float iq = 1.f / fmaxf(nq, 1e-12f), ik = 1.f / fmaxf(nk, 1e-12f);
      if (part == 0) {
        invq[p] = iq; invk[p] = ik;
        clampq[p] = nq < 1e-12f; clampk[p] = nk < 1e-12f;
      }
      const float sq2 = iq * LOG2E;
#pragma unroll
      for (int v = 0; v < VPT; ++v) {
        uint32_t qw[4] = {qv[v].x, qv[v].y, qv[v].z, qv[v].w}, kw[4] = {kv[v].x, kv[v].y, kv[v].z, kv[v].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          qw[e] = pack_bf16x2(bf_lo(qw[e]) * sq2, bf_hi(qw[e]) * sq2);
          kw[e] = pack_bf16x2(bf_lo(kw[e]) * ik, bf_hi(kw[e]) * ik);
        }
        if (act) {
          qp[8 * v] = make_uint4(qw[0], qw[1], qw[2], qw[3]);
          kp[8 * v] = make_uint4(kw[0], kw[1], kw[2], kw[3]);
        }
      }
    }
    __syncthreads();

    // ================= phase R: row slabs i0..i0+15, all j =================
    float dq_acc[SPW][2][4];
#pragma unroll
    for (int s = 0; s < SPW; ++s) {
      const int rt = warp + 4 * s;
#pragma unroll
      for (int a = 0; a < 2; ++a) dq_acc[s][a][0] = dq_acc[s][a][1] = dq_acc[s][a][2] = dq_acc[s][a][3] = 0.f;
      if (rt < NSLAB) {
        const int i0 = rt * 16;
        uint32_t qfr[4], dfr[4];
        ldsm_x4_t(qs_a + (r8 + 8 * (mi >> 1)) * PITCH + (i0 + 8 * (mi & 1)) * 2, qfr);   // Qh^T slab
        ldsm_x4_t(ds_a + (r8 + 8 * (mi >> 1)) * PITCH + (i0 + 8 * (mi & 1)) * 2, dfr);   // dO^T slab
        uint32_t pk[C / 8][2];                       // P = exp(S) as packed bf16: [nt][row g | row g+8]
        float rs[4] = {0.f, 0.f, 0.f, 0.f};
        float acc[JW / 8][4];
#pragma unroll
        for (int ch = 0; ch < NCHK; ++ch) {
          slab_times_chunk<C>(acc, qfr, ks_a, ch * JW, r8, mi);
#pragma unroll
          for (int nt = 0; nt < JW / 8; ++nt) {
            pk[ch * (JW / 8) + nt][0] = ex2_bf16x2(pack_bf16x2(acc[nt][0], acc[nt][1]));
            pk[ch * (JW / 8) + nt][1] = ex2_bf16x2(pack_bf16x2(acc[nt][2], acc[nt][3]));
          }
#pragma unroll
          for (int k2 = 0; k2 < JW / 16; ++k2) {     // row sums: P (as A operand) x ones
            const int t = ch * (JW / 8) + 2 * k2;
            const uint32_t pa[4] = {pk[t][0], pk[t][1], pk[t + 1][0], pk[t + 1][1]};
            mma_bf16(rs, pa, ONES, ONES);
          }
        }
        const float inv0 = 1.f / rs[0], inv1 = 1.f / rs[2];
        float D0 = 0.f, D1 = 0.f;
#pragma unroll
        for (int ch = 0; ch < NCHK; ++ch) {
          slab_times_chunk<C>(acc, dfr, vs_a, ch * JW, r8, mi);     // dA chunk
#pragma unroll
          for (int nt = 0; nt < JW / 8; ++nt) {
            const int t = ch * (JW / 8) + nt;
            D0 = fmaf(bf_lo(pk[t][0]), acc[nt][0], fmaf(bf_hi(pk[t][0]), acc[nt][1], D0));
            D1 = fmaf(bf_lo(pk[t][1]), acc[nt][2], fmaf(bf_hi(pk[t][1]), acc[nt][3], D1));
          }
        }
        D0 += __shfl_xor_sync(0xffffffffu, D0, 1); D0 += __shfl_xor_sync(0xffffffffu, D0, 2);
        D1 += __shfl_xor_sync(0xffffffffu, D1, 1); D1 += __shfl_xor_sync(0xffffffffu, D1, 2);
        D0 *= inv0; D1 *= inv1;
        if (q4 == 0) {
          rinv[i0 + g] = inv0; rinv[i0 + g + 8] = inv1;
          rdot[i0 + g] = D0; rdot[i0 + g + 8] = D1;
        }
#pragma unroll
        for (int ch = 0; ch < NCHK; ++ch) {
          if (NCHK > 1) slab_times_chunk<C>(acc, dfr, vs_a, ch * JW, r8, mi);   // (single chunk: still in registers)
          uint32_t dsb[JW / 8][2];
#pragma unroll
          for (int nt = 0; nt < JW / 8; ++nt) {
            const int t = ch * (JW / 8) + nt;
            dsb[nt][0] = pack_bf16x2(bf_lo(pk[t][0]) * inv0 * (acc[nt][0] - D0), bf_hi(pk[t][0]) * inv0 * (acc[nt][1] - D0));
            dsb[nt][1] = pack_bf16x2(bf_lo(pk[t][1]) * inv1 * (acc[nt][2] - D1), bf_hi(pk[t][1]) * inv1 * (acc[nt][3] - D1));
          }
#pragma unroll
          for (int k2 = 0; k2 < JW / 16; ++k2) {     // dqh[p][i] += sum_j kh[p][j] dS[i][j]
            uint32_t kfr[4];
            ldsm_x4(ks_a + (r8 + 8 * (mi & 1)) * PITCH + (ch * JW + 16 * k2 + 8 * (mi >> 1)) * 2, kfr);
            mma_bf16(dq_acc[s][0], kfr, dsb[2 * k2][0], dsb[2 * k2 + 1][0]);
            mma_bf16(dq_acc[s][1], kfr, dsb[2 * k2][1], dsb[2 * k2 + 1][1]);
          }
        }
      }
    }
    __syncthreads();

    // ================= phase C: column slabs j0..j0+15, all i (computed transposed) =================
    float dk_acc[SPW][2][4];
#pragma unroll
    for (int s = 0; s < SPW; ++s) {
      const int ct = warp + 4 * s;
#pragma unroll
      for (int a = 0; a < 2; ++a) dk_acc[s][a][0] = dk_acc[s][a][1] = dk_acc[s][a][2] = dk_acc[s][a][3] = 0.f;
      if (ct < NSLAB) {
        const int j0 = ct * 16;
        uint32_t kfrT[4], vfrT[4];
        ldsm_x4_t(ks_a + (r8 + 8 * (mi >> 1)) * PITCH + (j0 + 8 * (mi & 1)) * 2, kfrT);  // Kh^T slab
        ldsm_x4_t(vs_a + (r8 + 8 * (mi >> 1)) * PITCH + (j0 + 8 * (mi & 1)) * 2, vfrT);  // V^T slab
        float dv[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
        for (int ch = 0; ch < NCHK; ++ch) {
          float st[JW / 8][4], da[JW / 8][4];
          slab_times_chunk<C>(st, kfrT, qs_a, ch * JW, r8, mi);     // S^T chunk: rows j, cols i
          slab_times_chunk<C>(da, vfrT, ds_a, ch * JW, r8, mi);     // dA^T chunk
          uint32_t abf[JW / 8][2], dsb[JW / 8][2];
#pragma unroll
          for (int nt = 0; nt < JW / 8; ++nt) {
            const int i = ch * JW + 8 * nt + 2 * q4;
            const float2 ri = *reinterpret_cast<const float2*>(rinv + i);
            const float2 rd = *reinterpret_cast<const float2*>(rdot + i);
            const uint32_t p0 = ex2_bf16x2(pack_bf16x2(st[nt][0], st[nt][1]));
            const uint32_t p1 = ex2_bf16x2(pack_bf16x2(st[nt][2], st[nt][3]));
            const float a00 = bf_lo(p0) * ri.x, a01 = bf_hi(p0) * ri.y;
            const float a10 = bf_lo(p1) * ri.x, a11 = bf_hi(p1) * ri.y;
            abf[nt][0] = pack_bf16x2(a00, a01);
            abf[nt][1] = pack_bf16x2(a10, a11);
            dsb[nt][0] = pack_bf16x2(a00 * (da[nt][0] - rd.x), a01 * (da[nt][1] - rd.y));
            dsb[nt][1] = pack_bf16x2(a10 * (da[nt][2] - rd.x), a11 * (da[nt][3] - rd.y));
          }
#pragma unroll
          for (int k2 = 0; k2 < JW / 16; ++k2) {
            uint32_t dofr[4], qfrp[4];
            const uint32_t o = (r8 + 8 * (mi & 1)) * PITCH + (ch * JW + 16 * k2 + 8 * (mi >> 1)) * 2;
            ldsm_x4(ds_a + o, dofr);                 // dO  [M = p][K = i]
            ldsm_x4(qs_a + o, qfrp);                 // qh' [M = p][K = i]
            mma_bf16(dv[0], dofr, abf[2 * k2][0], abf[2 * k2 + 1][0]);
            mma_bf16(dv[1], dofr, abf[2 * k2][1], abf[2 * k2 + 1][1]);
            mma_bf16(dk_acc[s][0], qfrp, dsb[2 * k2][0], dsb[2 * k2 + 1][0]);
            mma_bf16(dk_acc[s][1], qfrp, dsb[2 * k2][1], dsb[2 * k2 + 1][1]);
          }
        }
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          const int col = (2 * C + j0 + 8 * nt + 2 * q4) * 2;
          *reinterpret_cast<uint32_t*>(ot + g * OPITCH + col) = pack_bf16x2(dv[nt][0], dv[nt][1]);
          *reinterpret_cast<uint32_t*>(ot + (g + 8) * OPITCH + col) = pack_bf16x2(dv[nt][2], dv[nt][3]);
        }
      }
    }

    // ---- <qh, dqh> and <kh, dkh> per pixel: partial sums from the accumulators
    {
      float dq0 = 0.f, dq1 = 0.f, dk0 = 0.f, dk1 = 0.f;
#pragma unroll
      for (int s = 0; s < SPW; ++s) {
        const int rt = warp + 4 * s;
        if (rt < NSLAB) {
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) {
            const int col = (rt * 16 + 8 * nt + 2 * q4) * 2;
            const uint32_t qa = *reinterpret_cast<const uint32_t*>(qs + g * PITCH + col);
            const uint32_t qb = *reinterpret_cast<const uint32_t*>(qs + (g + 8) * PITCH + col);
            const uint32_t ka = *reinterpret_cast<const uint32_t*>(ks + g * PITCH + col);
            const uint32_t kb = *reinterpret_cast<const uint32_t*>(ks + (g + 8) * PITCH + col);
            dq0 = fmaf(bf_lo(qa), dq_acc[s][nt][0], fmaf(bf_hi(qa), dq_acc[s][nt][1], dq0));
            dq1 = fmaf(bf_lo(qb), dq_acc[s][nt][2], fmaf(bf_hi(qb), dq_acc[s][nt][3], dq1));
            dk0 = fmaf(bf_lo(ka), dk_acc[s][nt][0], fmaf(bf_hi(ka), dk_acc[s][nt][1], dk0));
            dk1 = fmaf(bf_lo(kb), dk_acc[s][nt][2], fmaf(bf_hi(kb), dk_acc[s][nt][3], dk1));
          }
        }
      }
#pragma unroll
      for (int o = 1; o < 4; o <<= 1) {
        dq0 += __shfl_xor_sync(0xffffffffu, dq0, o); dq1 += __shfl_xor_sync(0xffffffffu, dq1, o);
        dk0 += __shfl_xor_sync(0xffffffffu, dk0, o); dk1 += __shfl_xor_sync(0xffffffffu, dk1, o);
      }
      if (q4 == 0 && warp < NSLAB) {
        atomicAdd(dotq + g, dq0); atomicAdd(dotq + g + 8, dq1);
        atomicAdd(dotk + g, dk0); atomicAdd(dotk + g + 8, dk1);
      }
    }
    __syncthreads();
    // ---- normalisation backward from the accumulators into the staging tile.  qs holds L qh (L = log2 e), dk_acc holds
    //      L dkh:  dq = (dqh - (L qh) <L qh, dqh> / L^2) / |q|,   dk = (L dkh - kh <kh, L dkh>) / (L |k|)
    {
      const float cq0 = clampq[g] ? 0.f : dotq[g] * (1.f / (LOG2E * LOG2E));
      const float cq1 = clampq[g + 8] ? 0.f : dotq[g + 8] * (1.f / (LOG2E * LOG2E));
      const float ck0 = clampk[g] ? 0.f : dotk[g], ck1 = clampk[g + 8] ? 0.f : dotk[g + 8];
      const float sq0 = invq[g], sq1 = invq[g + 8];
      const float sk0 = invk[g] * (1.f / LOG2E), sk1 = invk[g + 8] * (1.f / LOG2E);
#pragma unroll
      for (int s = 0; s < SPW; ++s) {
        const int rt = warp + 4 * s;
        if (rt < NSLAB) {
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) {
            const int c = rt * 16 + 8 * nt + 2 * q4;
            const uint32_t qa = *reinterpret_cast<const uint32_t*>(qs + g * PITCH + c * 2);
            const uint32_t qb = *reinterpret_cast<const uint32_t*>(qs + (g + 8) * PITCH + c * 2);
            const uint32_t ka = *reinterpret_cast<const uint32_t*>(ks + g * PITCH + c * 2);
            const uint32_t kb = *reinterpret_cast<const uint32_t*>(ks + (g + 8) * PITCH + c * 2);
            *reinterpret_cast<uint32_t*>(ot + g * OPITCH + c * 2) =
                pack_bf16x2((dq_acc[s][nt][0] - bf_lo(qa) * cq0) * sq0, (dq_acc[s][nt][1] - bf_hi(qa) * cq0) * sq0);
            *reinterpret_cast<uint32_t*>(ot + (g + 8) * OPITCH + c * 2) =
                pack_bf16x2((dq_acc[s][nt][2] - bf_lo(qb) * cq1) * sq1, (dq_acc[s][nt][3] - bf_hi(qb) * cq1) * sq1);
            *reinterpret_cast<uint32_t*>(ot + g * OPITCH + (C + c) * 2) =
                pack_bf16x2((dk_acc[s][nt][0] - bf_lo(ka) * ck0) * sk0, (dk_acc[s][nt][1] - bf_hi(ka) * ck0) * sk0);
            *reinterpret_cast<uint32_t*>(ot + (g + 8) * OPITCH + (C + c) * 2) =
                pack_bf16x2((dk_acc[s][nt][2] - bf_lo(kb) * ck1) * sk1, (dk_acc[s][nt][3] - bf_hi(kb) * ck1) * sk1);
          }
        }
      }
    }
    __syncthreads();
    // ---- staged dq | dk | dv -> dqkv (16-byte coalesced stores)
    for (int c = tid; c < P16 * 3 * CH; c += LB_THREADS) {
      const int p = c / (3 * CH), off = c - p * (3 * CH);
      const uint4 val = *reinterpret_cast<const uint4*>(ot + p * OPITCH + off * 16);
      *reinterpret_cast<uint4*>(dqkv + (pix0 + (size_t)(p >> 2) * W + (p & 3)) * (3 * C) + off * 8) = val;
    }
    __syncthreads();   // qs..ds, ot and the per-pixel scalars are rewritten by the next window
  }
}

template <int C>
int launch_bwd(const __nv_bfloat16* qkv, const __nv_bfloat16* dout, int N, int H, int W, __nv_bfloat16* dqkv,
               cudaStream_t st) {
  const size_t smem = BwdCfg<C>::SMEM;
  cudaError_t e = cudaFuncSetAttribute(local_attn_bwd_tc_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("local_attn_bwd_tc: smem attribute: %s", cudaGetErrorString(e)); return MSG_ERR_CUDA; }
  const long long nwin = (long long)N * (H / 4) * (W / 4);
  const int per_sm = C <= 64 ? 4 : (C == 128 ? 3 : 2);
  long long grid = (long long)per_sm * sm_count();
  if (grid > nwin) grid = nwin;
  local_attn_bwd_tc_kernel<C><<<(unsigned)grid, LB_THREADS, smem, st>>>(qkv, dout, N, H, W, dqkv);
  return check_launch("local_attn_bwd_tc_kernel");
}

}  // namespace

bool local_attn_bwd_tc_supported(int dtype, int C, const void* qkv, const void* dout, const void* dqkv) {
  if (dtype != MSG_BF16) return false;
  if (C != 32 && C != 64 && C != 128 && C != 256) return false;
  return (((uintptr_t)qkv | (uintptr_t)dout | (uintptr_t)dqkv) & 15) == 0;
}

int local_attn_bwd_tc(const void* qkv, const void* dout, int N, int H, int W, int C, void* dqkv, cudaStream_t st) {
  auto q = (const __nv_bfloat16*)qkv;
  auto d = (const __nv_bfloat16*)dout;
  auto o = (__nv_bfloat16*)dqkv;
  switch (C) {
    case 32: return launch_bwd<32>(q, d, N, H, W, o, st);
    case 64: return launch_bwd<64>(q, d, N, H, W, o, st);
    case 128: return launch_bwd<128>(q, d, N, H, W, o, st);
    case 256: return launch_bwd<256>(q, d, N, H, W, o, st);
  }
  set_error("local_attn_bwd_tc: unsupported C=%d", C);
  return MSG_ERR_UNSUPPORTED;
}

}  // namespace msg
