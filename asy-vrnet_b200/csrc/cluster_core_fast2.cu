// Cluster core, forward, compile-time path for THE live geometry of the backbone and of neck level p4: 16x16 regions,
// 2x2 centre proposal, D = 32 (every backbone stage) or D = 24 (neck).  Same maths as core_fwd_kernel (cluster_core.cu;
// reference backbone/fusion/vr_coc.py:158-190).
//
// Why a second fast kernel: ncu on core_fwd_fast_kernel (profiles/r01_ncu_final_cm_mlpf_core.csv) shows 15.6 K warp
// instructions per region-head at 64 % issue utilisation and 24 % DRAM utilisation - the run-time geometry (loop bounds,
// shifts, quadrant selects) and scalar FMAs cost 4x the instructions the arithmetic needs.  Here everything is a
// compile-time constant and the FMA streams are packed fp32x2 (FFMA2 / FADD2):
//   * 128 threads per CTA, 4 CTAs (fp32 value: 3) per SM, one region-head per iteration; feat/value tiles by 4-D TMA box.
//   * passes over channels: a thread owns TWO channels and every 8th item (item = 4 consecutive points of a region row), so
//     the four one-hot weight vectors of an item are fetched once per two channels, and both the row half (loop index) and
//     the column half (lane bit 1) of its items are static: quadrant sums are plain adds.
//   * similarity pass: a thread owns two neighbouring points; the normalised centres are stored duplicated (c,c) so that
//     one FFMA2 advances one centre for both points: 5 FFMA2 + 3 LDS per channel per point pair.
//   * the output tile aliases the VALUE tile (a channel's plane is rewritten by the lanes that consumed it), so the feat
//     tile is dead after the similarity pass and the next region-head's feat streams in under passes 3-4.
#include <stdlib.h>

#include "tma.cuh"

namespace vrcoc {
namespace {

constexpr int F2_THREADS = 128;
constexpr int F2_RS = 16;                  // region side
constexpr int F2_N = F2_RS * F2_RS;        // points per region
constexpr float F2_EPS = 1e-12f;

struct Fast2Cfg { int E, F1, F2, H, W, R; };

__device__ __forceinline__ uint64_t p2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void u2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ float hsum2(uint64_t v) { float a, b; u2(v, a, b); return a + b; }
__device__ __forceinline__ uint64_t fma2x(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t mul2x(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t add2x(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// one item (4 consecutive points) of a value plane as two fp32 pairs
template <typename T> __device__ __forceinline__ void ld_item(const T* plane, int item, uint64_t& v01, uint64_t& v23);
template <> __device__ __forceinline__ void ld_item<float>(const float* plane, int item, uint64_t& v01, uint64_t& v23) {
  const float4 x = *reinterpret_cast<const float4*>(plane + 4 * item);
  v01 = p2(x.x, x.y); v23 = p2(x.z, x.w);
}
template <> __device__ __forceinline__ void ld_item<__nv_bfloat16>(const __nv_bfloat16* plane, int item, uint64_t& v01, uint64_t& v23) {
  const uint2 raw = *reinterpret_cast<const uint2*>(plane + 4 * item);
  v01 = p2(__uint_as_float(raw.x << 16), __uint_as_float(raw.x & 0xffff0000u));
  v23 = p2(__uint_as_float(raw.y << 16), __uint_as_float(raw.y & 0xffff0000u));
}
template <typename T> __device__ __forceinline__ void st_item(T* plane, int item, uint64_t o01, uint64_t o23);
template <> __device__ __forceinline__ void st_item<float>(float* plane, int item, uint64_t o01, uint64_t o23) {
  float4 o;
  u2(o01, o.x, o.y); u2(o23, o.z, o.w);
  *reinterpret_cast<float4*>(plane + 4 * item) = o;
}
template <> __device__ __forceinline__ void st_item<__nv_bfloat16>(__nv_bfloat16* plane, int item, uint64_t o01, uint64_t o23) {
  float a, b, c, d;
  u2(o01, a, b); u2(o23, c, d);
  uint2 raw;
  *reinterpret_cast<__nv_bfloat162*>(&raw.x) = __floats2bfloat162_rn(a, b);
  *reinterpret_cast<__nv_bfloat162*>(&raw.y) = __floats2bfloat162_rn(c, d);
  *reinterpret_cast<uint2*>(plane + 4 * item) = raw;
}

template <int D, typename TV> struct F2Smem {
  static constexpr int f_bytes = D * F2_N * 4;
  static constexpr int v_bytes = D * F2_N * (int)sizeof(TV);
  static constexpr int off_f = 0;
  static constexpr int off_v = off_f + f_bytes;
  static constexpr int off_w = off_v + v_bytes;               // [4][256] one-hot weights, centre-major
  static constexpr int off_cm = off_w + 4 * F2_N * 4;         // [D][4] centre means
  static constexpr int off_cd = off_cm + D * 16;              // [D][4][2] normalised centres, duplicated
  static constexpr int off_misc = off_cd + D * 32;            // cnt[2] (packed 16-bit counters), bar_f, bar_v
  static constexpr int total = off_misc + 64 + 128;           // + alignment slack
};

template <int D, typename TV>
__global__ void __launch_bounds__(F2_THREADS, sizeof(TV) == 2 ? 4 : 3)
core_fwd_fast2_kernel(const __grid_constant__ CUtensorMap tmF, const __grid_constant__ CUtensorMap tmV,
                      const __grid_constant__ CUtensorMap tmO, uint8_t* __restrict__ idx_out, float* __restrict__ smax_out,
                      const float* __restrict__ alpha_p, const float* __restrict__ beta_p, Fast2Cfg G) {
  using S = F2Smem<D, TV>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t pad = (128u - (smem_u32(smem_raw) & 127u)) & 127u;
  unsigned char* smem = smem_raw + pad;
  const float* ft = reinterpret_cast<const float*>(smem + S::off_f);
  TV* vt = reinterpret_cast<TV*>(smem + S::off_v);            // value tile, rewritten in place as the output tile
  float* wq = reinterpret_cast<float*>(smem + S::off_w);
  float* cm = reinterpret_cast<float*>(smem + S::off_cm);
  float* cd = reinterpret_cast<float*>(smem + S::off_cd);
  int* cnt = reinterpret_cast<int*>(smem + S::off_misc);
  uint64_t* bar_f = reinterpret_cast<uint64_t*>(smem + S::off_misc + 16);
  uint64_t* bar_v = bar_f + 1;

  const int tid = threadIdx.x, lane = tid & 31;
  const int cp = tid >> 3, l = tid & 7;                       // channel pair, lane inside the pair's group
  const bool chan_live = cp < D / 2;                          // warp-uniform (4 pairs per warp, D/2 % 4 == 0)
  const int col = (l >> 1) & 1;                               // column half of every item this thread touches
  const int jx = cp & 1;                                      // odd pairs walk the item blocks pairwise swapped: the two
                                                              // pairs of a half-warp hit disjoint banks on 8-byte accesses
  const float alpha = __ldg(alpha_p), beta = __ldg(beta_p);
  constexpr float inv_quadrant = 4.0f / (float)F2_N;
  const int64_t HW = (int64_t)G.H * G.W;

  if (tid == 0) {
    mbar_init(bar_f, 1);
    mbar_init(bar_v, 1);
    mbar_fence_init();
    tma_prefetch_desc(&tmF); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmO);
  }
  __syncthreads();

  auto issue_f = [&](int r) {
    const int f2 = r % G.F2, f1 = (r / G.F2) % G.F1, be = r / (G.F1 * G.F2);
    mbar_expect_tx(bar_f, (uint32_t)S::f_bytes);
    tma_load_4d(smem + S::off_f, &tmF, f2 * F2_RS, f1 * F2_RS, (be % G.E) * D, be / G.E, bar_f);
  };
  auto issue_v = [&](int r) {
    const int f2 = r % G.F2, f1 = (r / G.F2) % G.F1, be = r / (G.F1 * G.F2);
    mbar_expect_tx(bar_v, (uint32_t)S::v_bytes);
    tma_load_4d(smem + S::off_v, &tmV, f2 * F2_RS, f1 * F2_RS, (be % G.E) * D, be / G.E, bar_v);
  };

  int r = blockIdx.x;
  if (tid == 0 && r < G.R) { issue_f(r); issue_v(r); }

  for (int it = 0; r < G.R; ++it, r += gridDim.x) {
    const uint32_t parity = (uint32_t)it & 1u;
    if (tid < 2) cnt[tid] = 0;
    mbar_wait(bar_f, parity);

    // ---- pass 1: centre proposal of feat = four quadrant means per channel --------------------------------------------
    if (chan_live) {
      float s[4];
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        const float* plane = ft + (2 * cp + ch) * F2_N;
        uint64_t top = 0ull, bot = 0ull;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 x = *reinterpret_cast<const float4*>(plane + 4 * (l + 8 * j));
          const uint64_t t = add2x(p2(x.x, x.y), p2(x.z, x.w));
          if (j < 4) top = add2x(top, t); else bot = add2x(bot, t);
        }
        s[2 * ch] = hsum2(top);
        s[2 * ch + 1] = hsum2(bot);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {                             // lanes l, l^1, l^4, l^5 share the column half
        s[k] += __shfl_xor_sync(0xffffffffu, s[k], 1);
        s[k] += __shfl_xor_sync(0xffffffffu, s[k], 4);
      }
      if ((l & 5) == 0) {                                       // l = 0: left quadrants (m = 0, 2); l = 2: right (m = 1, 3)
        float* c0 = cm + (2 * cp) * 4 + col;
        c0[0] = s[0] * inv_quadrant; c0[2] = s[1] * inv_quadrant;
        c0[4] = s[2] * inv_quadrant; c0[6] = s[3] * inv_quadrant;
      }
    }
    if (tid == 0 && it > 0) {                                   // previous output tile has left the value tile by now
      tma_store_wait_read();
      issue_v(r);
    }
    __syncthreads();
    if (tid < 32) {                                             // normalise the four centres: lane = (channel group, centre)
      const int m = lane & 3, g = lane >> 2;
      float c[D / 8], ss = 0.f;
#pragma unroll
      for (int t = 0; t < D / 8; ++t) { c[t] = cm[(g + 8 * t) * 4 + m]; ss = fmaf(c[t], c[t], ss); }
      ss += __shfl_xor_sync(0xffffffffu, ss, 4);
      ss += __shfl_xor_sync(0xffffffffu, ss, 8);
      ss += __shfl_xor_sync(0xffffffffu, ss, 16);
      const float inv = 1.0f / fmaxf(sqrtf(ss), F2_EPS);
#pragma unroll
      for (int t = 0; t < D / 8; ++t) {
        const float v = c[t] * inv;
        *reinterpret_cast<float2*>(cd + (g + 8 * t) * 8 + 2 * m) = make_float2(v, v);
      }
    }
    __syncthreads();

    // ---- pass 2: similarity, arg-max, gate -> one-hot weights; two neighbouring points per thread -------------------------
    {
      uint64_t ss = 0ull, d0 = 0ull, d1 = 0ull, d2 = 0ull, d3 = 0ull;
      const float* fp = ft + 2 * tid;
#pragma unroll 8
      for (int d = 0; d < D; ++d) {
        const uint64_t x = *reinterpret_cast<const uint64_t*>(fp + d * F2_N);
        const ulonglong2 c01 = *reinterpret_cast<const ulonglong2*>(cd + d * 8);
        const ulonglong2 c23 = *reinterpret_cast<const ulonglong2*>(cd + d * 8 + 4);
        ss = fma2x(x, x, ss);
        d0 = fma2x(c01.x, x, d0); d1 = fma2x(c01.y, x, d1);
        d2 = fma2x(c23.x, x, d2); d3 = fma2x(c23.y, x, d3);
      }
      float ssv[2], dv[4][2];
      u2(ss, ssv[0], ssv[1]);
      u2(d0, dv[0][0], dv[0][1]); u2(d1, dv[1][0], dv[1][1]); u2(d2, dv[2][0], dv[2][1]); u2(d3, dv[3][0], dv[3][1]);
      int kb[2];
      float gv[2];
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        const float inv = fminf(rsqrtf(ssv[p]), 1.0f / F2_EPS);       // = 1 / max(|f|, eps) to 2 ulp (MUFU.RSQ)
        float tb = alpha * dv[0][p], db = dv[0][p];
        int k = 0;
        if (alpha * dv[1][p] > tb) { tb = alpha * dv[1][p]; db = dv[1][p]; k = 1; }
        if (alpha * dv[2][p] > tb) { tb = alpha * dv[2][p]; db = dv[2][p]; k = 2; }
        if (alpha * dv[3][p] > tb) { tb = alpha * dv[3][p]; db = dv[3][p]; k = 3; }
        kb[p] = k;
        gv[p] = __fdividef(1.0f, 1.0f + __expf(-fmaf(alpha, db * inv, beta)));   // sigmoid to ~3 ulp; the arg-max above does
                                                                                 // not depend on it
      }
#pragma unroll
      for (int m = 0; m < 4; ++m)
        *reinterpret_cast<float2*>(wq + m * F2_N + 2 * tid) = make_float2(kb[0] == m ? gv[0] : 0.f, kb[1] == m ? gv[1] : 0.f);
      if (idx_out || smax_out) {
        const int f2 = r % G.F2, f1 = (r / G.F2) % G.F1, be = r / (G.F1 * G.F2);
        const int n = 2 * tid;
        const int64_t io = (int64_t)be * HW + (int64_t)(f1 * F2_RS + (n >> 4)) * G.W + f2 * F2_RS + (n & 15);
        if (idx_out) *reinterpret_cast<uint16_t*>(idx_out + io) = (uint16_t)(kb[0] | (kb[1] << 8));
        if (smax_out) *reinterpret_cast<float2*>(smax_out + io) = make_float2(gv[0], gv[1]);
      }
      // members per centre: two packed words of 16-bit counters (a region has 256 points)
      const unsigned pa = (unsigned)((kb[0] == 0) + (kb[1] == 0)) | ((unsigned)((kb[0] == 1) + (kb[1] == 1)) << 16);
      const unsigned pb = (unsigned)((kb[0] == 2) + (kb[1] == 2)) | ((unsigned)((kb[0] == 3) + (kb[1] == 3)) << 16);
      const unsigned ra = __reduce_add_sync(0xffffffffu, pa), rb = __reduce_add_sync(0xffffffffu, pb);
      if (lane == 0) { atomicAdd(&cnt[0], (int)ra); atomicAdd(&cnt[1], (int)rb); }
    }
    __syncthreads();                                            // weights + counters complete; the feat tile is dead
    const int rn = r + (int)gridDim.x;
    if (tid == 0 && rn < G.R) issue_f(rn);
    mbar_wait(bar_v, parity);

    // ---- pass 3 + 4: aggregate value to the centres, dispatch back to the points (in place) ---------------------------------
    if (chan_live) {
      const unsigned ca = (unsigned)cnt[0], cb = (unsigned)cnt[1];
      // MUFU.RCP of a small integer (1 ulp): the IEEE division costs 13 instructions and a slow-path call each
      const float den0 = __fdividef(1.0f, (float)((ca & 0xffffu) + 1u)), den1 = __fdividef(1.0f, (float)((ca >> 16) + 1u));
      const float den2 = __fdividef(1.0f, (float)((cb & 0xffffu) + 1u)), den3 = __fdividef(1.0f, (float)((cb >> 16) + 1u));
      TV* plane0 = vt + (2 * cp) * F2_N;
      TV* plane1 = plane0 + F2_N;
      uint64_t A0[4] = {0ull, 0ull, 0ull, 0ull}, A1[4] = {0ull, 0ull, 0ull, 0ull};
      uint64_t qt0 = 0ull, qb0 = 0ull, qt1 = 0ull, qb1 = 0ull;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int i = l + 8 * (j ^ jx);
        uint64_t wlo[4], whi[4];
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const float4 w = *reinterpret_cast<const float4*>(wq + m * F2_N + 4 * i);
          wlo[m] = p2(w.x, w.y); whi[m] = p2(w.z, w.w);
        }
        uint64_t v01, v23, u01, u23;
        ld_item<TV>(plane0, i, v01, v23);
        ld_item<TV>(plane1, i, u01, u23);
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          A0[m] = fma2x(wlo[m], v01, A0[m]); A0[m] = fma2x(whi[m], v23, A0[m]);
          A1[m] = fma2x(wlo[m], u01, A1[m]); A1[m] = fma2x(whi[m], u23, A1[m]);
        }
        const uint64_t q0 = add2x(v01, v23), q1 = add2x(u01, u23);
        if (j < 4) { qt0 = add2x(qt0, q0); qt1 = add2x(qt1, q1); } else { qb0 = add2x(qb0, q0); qb1 = add2x(qb1, q1); }
      }
      float a0[4], a1[4];
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        a0[m] = hsum2(A0[m]); a1[m] = hsum2(A1[m]);
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
          a0[m] += __shfl_xor_sync(0xffffffffu, a0[m], o);
          a1[m] += __shfl_xor_sync(0xffffffffu, a1[m], o);
        }
      }
      float q[4] = {hsum2(qt0), hsum2(qb0), hsum2(qt1), hsum2(qb1)}, qo[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        q[k] += __shfl_xor_sync(0xffffffffu, q[k], 1);
        q[k] += __shfl_xor_sync(0xffffffffu, q[k], 4);
        qo[k] = __shfl_xor_sync(0xffffffffu, q[k], 2);          // the other column half
      }
      // quadrant m = 2 * (bottom) + (right)
      const float Q00 = col ? qo[0] : q[0], Q01 = col ? q[0] : qo[0], Q02 = col ? qo[1] : q[1], Q03 = col ? q[1] : qo[1];
      const float Q10 = col ? qo[2] : q[2], Q11 = col ? q[2] : qo[2], Q12 = col ? qo[3] : q[3], Q13 = col ? q[3] : qo[3];
      asm volatile("" ::: "memory");                            // pass 4 re-reads the weights: carrying 128 registers of
                                                                // them across the reduction spills to local memory
      uint64_t e0[4], e1[4];
      {
        const float x0 = fmaf(Q00, inv_quadrant, a0[0]) * den0, x1 = fmaf(Q01, inv_quadrant, a0[1]) * den1;
        const float x2 = fmaf(Q02, inv_quadrant, a0[2]) * den2, x3 = fmaf(Q03, inv_quadrant, a0[3]) * den3;
        const float y0 = fmaf(Q10, inv_quadrant, a1[0]) * den0, y1 = fmaf(Q11, inv_quadrant, a1[1]) * den1;
        const float y2 = fmaf(Q12, inv_quadrant, a1[2]) * den2, y3 = fmaf(Q13, inv_quadrant, a1[3]) * den3;
        e0[0] = p2(x0, x0); e0[1] = p2(x1, x1); e0[2] = p2(x2, x2); e0[3] = p2(x3, x3);
        e1[0] = p2(y0, y0); e1[1] = p2(y1, y1); e1[2] = p2(y2, y2); e1[3] = p2(y3, y3);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int i = l + 8 * (j ^ jx);
        uint64_t wlo[4], whi[4];
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const float4 w = *reinterpret_cast<const float4*>(wq + m * F2_N + 4 * i);
          wlo[m] = p2(w.x, w.y); whi[m] = p2(w.z, w.w);
        }
        const uint64_t o01 = fma2x(wlo[0], e0[0], fma2x(wlo[1], e0[1], fma2x(wlo[2], e0[2], mul2x(wlo[3], e0[3]))));
        const uint64_t o23 = fma2x(whi[0], e0[0], fma2x(whi[1], e0[1], fma2x(whi[2], e0[2], mul2x(whi[3], e0[3]))));
        const uint64_t r01 = fma2x(wlo[0], e1[0], fma2x(wlo[1], e1[1], fma2x(wlo[2], e1[2], mul2x(wlo[3], e1[3]))));
        const uint64_t r23 = fma2x(whi[0], e1[0], fma2x(whi[1], e1[1], fma2x(whi[2], e1[2], mul2x(whi[3], e1[3]))));
        st_item<TV>(plane0, i, o01, o23);
        st_item<TV>(plane1, i, r01, r23);
      }
    }
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      const int f2 = r % G.F2, f1 = (r / G.F2) % G.F1, be = r / (G.F1 * G.F2);
      tma_store_4d(&tmO, vt, f2 * F2_RS, f1 * F2_RS, (be % G.E) * D, be / G.E);
      tma_store_commit();                                       // the next value tile is requested after pass 1 of the next
    }                                                           // iteration, when this store has long read the tile
  }
  if (tid == 0) tma_store_wait_all();
}

template <int D, typename TV>
int launch_fast2(const CUtensorMap& tf, const CUtensorMap& tv, const CUtensorMap& to, uint8_t* idx, float* smax, const float* alpha,
                 const float* beta, const Fast2Cfg& G, cudaStream_t st) {
  using S = F2Smem<D, TV>;
  auto kern = core_fwd_fast2_kernel<D, TV>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::total);
  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  int per_sm = (227 * 1024) / (S::total + 1024);
  if (per_sm > 4) per_sm = 4;
  int grid = sm_count() * per_sm;
  if (grid > G.R) grid = G.R;
  kern<<<grid, F2_THREADS, S::total, st>>>(tf, tv, to, idx, smax, alpha, beta, G);
  return check_launch("cluster_core_fwd_fast2");
}

}  // namespace

// returns VRCOC_OK when launched, 1 when the geometry / storage is not covered (caller falls back), <0 on error
int cluster_core_fwd_fast2(const void* feat, int fdt, const void* value, int vdt, void* out, int odt, uint8_t* idx, float* smax,
                           const float* alpha, const float* beta, int B, int E, int D, int H, int W, int F1, int F2, int pw, int ph,
                           int64_t bs_f, int64_t bs_v, int64_t bs_o, cudaStream_t st) {
  const char* knob = getenv("VRCOC_CORE_FAST2");                 // "0": A/B switch for tools/microbench.py and the parity tests
  if ((knob && knob[0] == '0') || pw != 2 || ph != 2 || tma_encode_fn() == nullptr) return 1;
  if (H % F1 || W % F2 || H / F1 != F2_RS || W / F2 != F2_RS || (D != 32 && D != 24)) return 1;
  if (fdt != VRCOC_F32 || vdt != odt) return 1;
  if ((reinterpret_cast<uintptr_t>(idx) & 1) || (reinterpret_cast<uintptr_t>(smax) & 7)) return 1;
  auto ok = [&](const void* p, int dt, int64_t bs) {
    const int es = dt == VRCOC_F32 ? 4 : 2;
    return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && ((int64_t)W * es) % 16 == 0 && ((int64_t)H * W * es) % 16 == 0 && (bs * es) % 16 == 0;
  };
  if (!ok(feat, fdt, bs_f) || !ok(value, vdt, bs_v) || !ok(out, odt, bs_o)) return 1;
  Fast2Cfg G{E, F1, F2, H, W, B * E * F1 * F2};
  CUtensorMap tf, tv, to;
  auto enc = [&](CUtensorMap* tm, const void* base, int dt, int64_t bs) {
    const int es = dt == VRCOC_F32 ? 4 : 2;
    cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)(E * D), (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)W * es, (cuuint64_t)H * W * es, (cuuint64_t)bs * es};
    cuuint32_t box[4] = {(cuuint32_t)F2_RS, (cuuint32_t)F2_RS, (cuuint32_t)D, 1};
    return tma_encode(tm, dt, base, 4, dims, strides, box, false);
  };
  int rc;
  if ((rc = enc(&tf, feat, fdt, bs_f)) || (rc = enc(&tv, value, vdt, bs_v)) || (rc = enc(&to, out, odt, bs_o))) return rc;
#define F2_GO(DD, TT) return launch_fast2<DD, TT>(tf, tv, to, idx, smax, alpha, beta, G, st)
  if (vdt == VRCOC_BF16) { if (D == 32) F2_GO(32, __nv_bfloat16); else F2_GO(24, __nv_bfloat16); }
  if (vdt == VRCOC_F32) { if (D == 32) F2_GO(32, float); else F2_GO(24, float); }
#undef F2_GO
  return 1;
}

}  // namespace vrcoc
