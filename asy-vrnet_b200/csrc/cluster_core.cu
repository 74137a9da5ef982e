// Cluster core: centre proposal, cosine similarity, sigmoid gate, arg-max assignment, aggregation, dispatch.
// Replaces the eager-op chain of Cluster.forward between the projections
// (reference backbone/fusion/vr_coc.py:158-190 == backbone/vision/context_cluster.py:130-162), which materialises two
// [R,M,N,D] temporaries; here one CTA owns one region-head, stages its feat/value tile in shared memory once
// (HBM traffic = read feat + read value + write out) and runs every pass on chip.
//
// Maths: SURVEY appendix A.  Per region-head, points n (row-major inside the region), centres m = i*ph + j:
//   c_m = binmean(f), vc_m = binmean(v); z_mn = <c_m/|c_m|, f_n/|f_n|>; s_mn = sigmoid(beta + alpha z_mn)
//   k_n = argmax_m s_mn (lowest m on ties), g_n = s_{k_n n}
//   a_m = (sum_{k_n=m} g_n v_n + vc_m) / (cnt_m + 1);  o_n = g_n a_{k_n}
//
// Two pass shapes are used:
//   point-major   : thread = point, loop over d   (norms, dots, arg-max, dispatch) — smem reads conflict-free
//   channel-major : thread = (d, sub), TPD lanes per d, loop over points, warp-shuffle tree over the TPD lanes
//                   (centre sums, aggregation) — the d-stride NS is padded so the 32/TPD channels of a warp
//                   fall in disjoint banks.
// STAGED=false is the general fallback for regions that do not fit in shared memory: the same passes read
// feat/value straight from global/L2 through a per-point offset table.
#include <stdlib.h>

#include "common.cuh"

namespace vrcoc {

int cluster_core_fwd_fast(const void* feat, int fdt, const void* value, int vdt, void* out, int odt, uint8_t* idx, float* smax,
                          const float* alpha, const float* beta, int B, int E, int D, int H, int W, int F1, int F2, int pw, int ph,
                          int64_t bs_f, int64_t bs_v, int64_t bs_o, cudaStream_t st);   // cluster_core_fast.cu
int cluster_core_fwd_fast2(const void* feat, int fdt, const void* value, int vdt, void* out, int odt, uint8_t* idx, float* smax,
                           const float* alpha, const float* beta, int B, int E, int D, int H, int W, int F1, int F2, int pw, int ph,
                           int64_t bs_f, int64_t bs_v, int64_t bs_o, cudaStream_t st);  // cluster_core_fast2.cu

constexpr int CORE_THREADS = 256;
constexpr float NORM_EPS = 1e-12f;  // F.normalize eps (vr_coc.py:121-122)

struct CoreGeom {
  int B, E, D, H, W;
  int F1, F2;      // effective folds (1,1 when the reference's `fold_w>1 and fold_h>1` test fails)
  int rw, rh;      // region extent along dim2 / dim3
  int pw, ph;      // proposal bins along dim2 / dim3
  int N, M, NS, TPD;
  int64_t bs_f, bs_v, bs_o, bs_df, bs_dv;   // batch strides (elements) of feat, value, out|dout, dfeat, dvalue
};

struct CoreSmem {
  // byte offsets into dynamic shared memory
  int f, v, dout, g, z, inv, dz, off, k, mask, chat, cnorm, vc, agg, dA, dchat, cnt, binv, red;
  int total;
};

static CoreSmem core_layout(const CoreGeom& q, bool staged, bool bwd, int maxm) {
  CoreSmem s{};
  int p = 0;
  auto take = [&](int bytes) { int o = p; p += (bytes + 15) & ~15; return o; };
  int tile = staged ? q.D * q.NS * 4 : 0;
  s.f = take(tile);
  s.v = take(tile);
  s.dout = take(bwd ? tile : 0);
  s.g = take(q.N * 4);
  s.z = take(bwd ? q.N * 4 : 0);
  s.inv = take(bwd ? q.N * 4 : 0);
  s.dz = take(bwd ? q.N * 4 : 0);
  s.off = take(q.N * 4);
  s.k = take(q.N);
  s.mask = take(q.N * (maxm > 16 ? 8 : 2));        // bin-membership bit mask per point: 16 bits, 64 for the wide-proposal path
  s.chat = take(q.D * maxm * 4);
  s.cnorm = take(maxm * 4);
  s.vc = take(q.D * maxm * 4);
  s.agg = take(q.D * maxm * 4);
  s.dA = take(bwd ? q.D * maxm * 4 : 0);
  s.dchat = take(bwd ? q.D * maxm * 4 : 0);
  s.cnt = take(maxm * 4);
  s.binv = take(maxm * 4);
  s.red = take(64 * 4);
  s.total = p;
  return s;
}

template <typename T>
struct TileAccess {
  const float* sm;   // staged tile [D][NS] (nullptr in direct mode)
  const T* gbase;    // global base of this region-head's channel 0
  const int* off;    // per-point global offset
  int NS;
  int64_t HW;
  template <bool STAGED>
  __device__ __forceinline__ float at(int d, int n) const {
    if (STAGED) return sm[d * NS + n];
    return ldf<T>(gbase + (int64_t)d * HW + off[n]);
  }
};

// Cooperative tile load global -> smem (fp32), 128-bit vectorised when the region rows allow it.
template <typename T>
__device__ __forceinline__ void stage_tile(float* sm, const T* gbase, const CoreGeom& q, int row0, int col0) {
  const int64_t HW = (int64_t)q.H * q.W;
  const bool vec = (q.rh % 4 == 0) && (q.W % 4 == 0) && ((reinterpret_cast<uintptr_t>(gbase) & 15) == 0) &&
                   (col0 % 4 == 0) && ((HW * sizeof(T)) % 16 == 0);
  if (vec) {
    const int upr = q.rh / 4;               // 4-element units per region row
    const int units = q.D * q.rw * upr;
    for (int u = threadIdx.x; u < units; u += blockDim.x) {
      int c4 = u % upr;
      int r = (u / upr) % q.rw;
      int d = u / (upr * q.rw);
      const T* p = gbase + (int64_t)d * HW + (int64_t)(row0 + r) * q.W + col0 + c4 * 4;
      float4 x;
      if (sizeof(T) == 4) {
        x = __ldg(reinterpret_cast<const float4*>(p));
      } else {
        uint2 raw = __ldg(reinterpret_cast<const uint2*>(p));
        float2 a = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&raw.x));
        float2 b = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&raw.y));
        x = make_float4(a.x, a.y, b.x, b.y);
      }
      float* dst = sm + d * q.NS + r * q.rh + c4 * 4;   // NS % 4 == 0 is guaranteed by the host when vec holds
      *reinterpret_cast<float4*>(dst) = x;
    }
  } else {
    const int total = q.D * q.N;
    for (int u = threadIdx.x; u < total; u += blockDim.x) {
      int n = u % q.N, d = u / q.N;
      int r = n / q.rh, c = n % q.rh;
      sm[d * q.NS + n] = ldf<T>(gbase + (int64_t)d * HW + (int64_t)(row0 + r) * q.W + col0 + c);
    }
  }
}

__device__ __forceinline__ void region_origin(const CoreGeom& q, int rh_idx, int& b, int& e, int& row0, int& col0) {
  int f2 = rh_idx % q.F2;
  int f1 = (rh_idx / q.F2) % q.F1;
  int be = rh_idx / (q.F1 * q.F2);
  e = be % q.E;
  b = be / q.E;
  row0 = f1 * q.rw;
  col0 = f2 * q.rh;
}

// bins of AdaptiveAvgPool2d: [floor(i*n/p), ceil((i+1)*n/p))
__device__ __forceinline__ int bin_lo(int i, int n, int p) { return (i * n) / p; }
__device__ __forceinline__ int bin_hi(int i, int n, int p) { return ((i + 1) * n + p - 1) / p; }

// Shared prologue: per-point offsets and bin-membership masks, 1/|bin|.
// bin-membership masks: one bit per centre (AdaptiveAvgPool2d bins overlap when the region is not divisible by the proposal)
template <int MAXM> struct MaskOf { using type = uint16_t; };
template <> struct MaskOf<64> { using type = uint64_t; };

template <typename MT>
__device__ __forceinline__ void setup_points(const CoreGeom& q, int row0, int col0, int* off, MT* mask,
                                             float* binv, int* cnt) {
  for (int n = threadIdx.x; n < q.N; n += blockDim.x) {
    int r = n / q.rh, c = n % q.rh;
    off[n] = (row0 + r) * q.W + col0 + c;
    uint64_t mk = 0;
    for (int i = 0; i < q.pw; ++i) {
      if (r < bin_lo(i, q.rw, q.pw) || r >= bin_hi(i, q.rw, q.pw)) continue;
      for (int j = 0; j < q.ph; ++j)
        if (c >= bin_lo(j, q.rh, q.ph) && c < bin_hi(j, q.rh, q.ph)) mk |= 1ull << (i * q.ph + j);
    }
    mask[n] = (MT)mk;
  }
  if (threadIdx.x < q.M) {
    int i = threadIdx.x / q.ph, j = threadIdx.x % q.ph;
    int cntp = (bin_hi(i, q.rw, q.pw) - bin_lo(i, q.rw, q.pw)) * (bin_hi(j, q.rh, q.ph) - bin_lo(j, q.rh, q.ph));
    binv[threadIdx.x] = 1.0f / (float)cntp;
    cnt[threadIdx.x] = 0;
  }
}

// channel-major masked sums:  dst[d][m] = scale_m * sum_{n in bin m} A(d,n)      (centre proposal)
template <int MAXM, bool STAGED, typename T>
__device__ __forceinline__ void bin_means(const TileAccess<T>& A, const CoreGeom& q, const typename MaskOf<MAXM>::type* mask,
                                          const float* binv, float* dst /*[D][MAXM]*/) {
  const int sub = threadIdx.x % q.TPD;
  const int dpb = blockDim.x / q.TPD;             // channels per sweep
  for (int d0 = 0; d0 < q.D; d0 += dpb) {
    const int d = d0 + threadIdx.x / q.TPD;
    float acc[MAXM];
#pragma unroll
    for (int m = 0; m < MAXM; ++m) acc[m] = 0.f;
    if (d < q.D) {
      for (int n = sub; n < q.N; n += q.TPD) {
        float x = A.template at<STAGED>(d, n);
        const uint64_t mk = mask[n];
#pragma unroll
        for (int m = 0; m < MAXM; ++m) acc[m] += ((mk >> m) & 1ull) ? x : 0.f;
      }
    }
#pragma unroll
    for (int m = 0; m < MAXM; ++m)
      for (int o = q.TPD >> 1; o > 0; o >>= 1) acc[m] += __shfl_xor_sync(0xffffffffu, acc[m], o);
    if (d < q.D && sub == 0) {
#pragma unroll
      for (int m = 0; m < MAXM; ++m)
        if (m < q.M) dst[d * MAXM + m] = acc[m] * binv[m];
    }
  }
}

// channel-major assignment-weighted sums: dst[d][m] = sum_{k_n = m} wgt_n * A(d,n)
template <int MAXM, bool STAGED, typename T>
__device__ __forceinline__ void assigned_sums(const TileAccess<T>& A, const CoreGeom& q, const uint8_t* k,
                                              const float* wgt, float* dst /*[D][MAXM]*/) {
  const int sub = threadIdx.x % q.TPD;
  const int dpb = blockDim.x / q.TPD;
  for (int d0 = 0; d0 < q.D; d0 += dpb) {
    const int d = d0 + threadIdx.x / q.TPD;
    float acc[MAXM];
#pragma unroll
    for (int m = 0; m < MAXM; ++m) acc[m] = 0.f;
    if (d < q.D) {
      for (int n = sub; n < q.N; n += q.TPD) {
        float x = A.template at<STAGED>(d, n) * wgt[n];
        int kk = k[n];
#pragma unroll
        for (int m = 0; m < MAXM; ++m) acc[m] += (kk == m) ? x : 0.f;
      }
    }
#pragma unroll
    for (int m = 0; m < MAXM; ++m)
      for (int o = q.TPD >> 1; o > 0; o >>= 1) acc[m] += __shfl_xor_sync(0xffffffffu, acc[m], o);
    if (d < q.D && sub == 0) {
#pragma unroll
      for (int m = 0; m < MAXM; ++m)
        if (m < q.M) dst[d * MAXM + m] = acc[m];
    }
  }
}

// normalise centres in place: chat[d][m] = c/max(|c|,eps); cnorm[m] = |c|
template <int MAXM>
__device__ __forceinline__ void normalise_centres(const CoreGeom& q, float* chat, float* cnorm) {
  if (threadIdx.x < q.M) {
    int m = threadIdx.x;
    float ss = 0.f;
    for (int d = 0; d < q.D; ++d) { float c = chat[d * MAXM + m]; ss += c * c; }
    float nrm = sqrtf(ss);
    cnorm[m] = nrm;
    float inv = 1.0f / fmaxf(nrm, NORM_EPS);
    for (int d = 0; d < q.D; ++d) chat[d * MAXM + m] *= inv;
  }
}

template <typename TF, typename TV, typename TO, int MAXM, bool STAGED>
__global__ void __launch_bounds__(CORE_THREADS)
core_fwd_kernel(const TF* __restrict__ feat, const TV* __restrict__ value, TO* __restrict__ out,
                uint8_t* __restrict__ idx_out, float* __restrict__ smax_out,
                const float* __restrict__ alpha_p, const float* __restrict__ beta_p, CoreGeom q, CoreSmem L) {
  extern __shared__ __align__(16) unsigned char smem[];
  float* sf = reinterpret_cast<float*>(smem + L.f);
  float* sv = reinterpret_cast<float*>(smem + L.v);
  float* g = reinterpret_cast<float*>(smem + L.g);
  int* off = reinterpret_cast<int*>(smem + L.off);
  uint8_t* k = smem + L.k;
  using MT = typename MaskOf<MAXM>::type;
  MT* mask = reinterpret_cast<MT*>(smem + L.mask);
  float* chat = reinterpret_cast<float*>(smem + L.chat);
  float* cnorm = reinterpret_cast<float*>(smem + L.cnorm);
  float* vc = reinterpret_cast<float*>(smem + L.vc);
  float* agg = reinterpret_cast<float*>(smem + L.agg);
  int* cnt = reinterpret_cast<int*>(smem + L.cnt);
  float* binv = reinterpret_cast<float*>(smem + L.binv);

  int b, e, row0, col0;
  region_origin(q, blockIdx.x, b, e, row0, col0);
  const int64_t HW = (int64_t)q.H * q.W;
  const int64_t chan0 = (int64_t)e * q.D;
  const TF* fbase = feat + b * q.bs_f + chan0 * HW;
  const TV* vbase = value + b * q.bs_v + chan0 * HW;
  const float alpha = __ldg(alpha_p), beta = __ldg(beta_p);

  setup_points(q, row0, col0, off, mask, binv, cnt);
  if (STAGED) {
    stage_tile<TF>(sf, fbase, q, row0, col0);
    stage_tile<TV>(sv, vbase, q, row0, col0);
  }
  __syncthreads();
  TileAccess<TF> AF{sf, fbase, off, q.NS, HW};
  TileAccess<TV> AV{sv, vbase, off, q.NS, HW};

  // pass 1: centre proposal (AdaptiveAvgPool2d of feat and value)
  bin_means<MAXM, STAGED>(AF, q, mask, binv, chat);
  bin_means<MAXM, STAGED>(AV, q, mask, binv, vc);
  __syncthreads();
  normalise_centres<MAXM>(q, chat, cnorm);
  __syncthreads();

  // pass 2: similarity, gate, arg-max
  for (int n = threadIdx.x; n < q.N; n += blockDim.x) {
    float ss = 0.f, dot[MAXM];
#pragma unroll
    for (int m = 0; m < MAXM; ++m) dot[m] = 0.f;
    for (int d = 0; d < q.D; ++d) {
      float x = AF.template at<STAGED>(d, n);
      ss = fmaf(x, x, ss);
#pragma unroll
      for (int m = 0; m < MAXM; ++m) dot[m] = fmaf(chat[d * MAXM + m], x, dot[m]);
    }
    float inv = 1.0f / fmaxf(sqrtf(ss), NORM_EPS);
    float best = -1.f;
    int bi = 0;
#pragma unroll
    for (int m = 0; m < MAXM; ++m) {
      if (m < q.M) {
        float s = sigmoidf_exact(fmaf(alpha, dot[m] * inv, beta));
        if (s > best) { best = s; bi = m; }
      }
    }
    g[n] = best;
    k[n] = (uint8_t)bi;
    atomicAdd(&cnt[bi], 1);
    int64_t io = ((int64_t)b * q.E + e) * HW + off[n];
    if (idx_out) idx_out[io] = (uint8_t)bi;
    if (smax_out) smax_out[io] = best;
  }
  __syncthreads();

  // pass 3: aggregate to centres
  assigned_sums<MAXM, STAGED>(AV, q, k, g, agg);
  __syncthreads();
  for (int i = threadIdx.x; i < q.D * q.M; i += blockDim.x) {
    int d = i / q.M, m = i % q.M;
    agg[d * MAXM + m] = (agg[d * MAXM + m] + vc[d * MAXM + m]) / ((float)cnt[m] + 1.0f);
  }
  __syncthreads();

  // pass 4: dispatch back to the points
  TO* obase = out + b * q.bs_o + chan0 * HW;
  for (int n = threadIdx.x; n < q.N; n += blockDim.x) {
    float gn = g[n];
    int kk = k[n];
    int o = off[n];
    for (int d = 0; d < q.D; ++d) stf<TO>(obase + (int64_t)d * HW + o, gn * agg[d * MAXM + kk]);
  }
}

// Backward.  Inputs: feat, value, dout, saved idx + sim_max.  Outputs: dfeat, dvalue, per-CTA (dalpha, dbeta).
// FD > 0: the live geometry at compile time (16x16 regions, 2x2 proposals, head_dim FD): the same code with every loop bound,
// divisor and stride a constant.  ncu on the run-time-geometry instance at stage 1 (profiles/r02_ncu_core_bwd.txt): 147.5 M warp
// instructions per launch = 281 per (point, channel) element, the hot lines being the integer divisions / modulos by q.rh, q.TPD,
// q.N in the inner loops; 254 us where the tensors it touches (235 MB) would take 39 us at the HBM rate.
template <typename TF, typename TV, typename TG, typename TDF, typename TDV, int MAXM, bool STAGED, int FD = 0>
__global__ void __launch_bounds__(CORE_THREADS)
core_bwd_kernel(const TF* __restrict__ feat, const TV* __restrict__ value, const TG* __restrict__ dout,
                const uint8_t* __restrict__ idx_in, const float* __restrict__ smax_in,
                const float* __restrict__ alpha_p, TDF* __restrict__ dfeat, TDV* __restrict__ dvalue,
                float* __restrict__ partials, CoreGeom q_in, CoreSmem L) {
  CoreGeom q = q_in;
  if (FD > 0) { q.D = FD; q.rw = 16; q.rh = 16; q.pw = 2; q.ph = 2; q.N = 256; q.M = 4; q.NS = 264; q.TPD = 8; }
  extern __shared__ __align__(16) unsigned char smem[];
  float* sf = reinterpret_cast<float*>(smem + L.f);
  float* sv = reinterpret_cast<float*>(smem + L.v);
  float* sg = reinterpret_cast<float*>(smem + L.dout);
  float* g = reinterpret_cast<float*>(smem + L.g);
  float* z = reinterpret_cast<float*>(smem + L.z);
  float* inv = reinterpret_cast<float*>(smem + L.inv);
  float* dz = reinterpret_cast<float*>(smem + L.dz);
  int* off = reinterpret_cast<int*>(smem + L.off);
  uint8_t* k = smem + L.k;
  using MT = typename MaskOf<MAXM>::type;
  MT* mask = reinterpret_cast<MT*>(smem + L.mask);
  float* chat = reinterpret_cast<float*>(smem + L.chat);
  float* cnorm = reinterpret_cast<float*>(smem + L.cnorm);
  float* vc = reinterpret_cast<float*>(smem + L.vc);
  float* agg = reinterpret_cast<float*>(smem + L.agg);
  float* dA = reinterpret_cast<float*>(smem + L.dA);
  float* dch = reinterpret_cast<float*>(smem + L.dchat);
  int* cnt = reinterpret_cast<int*>(smem + L.cnt);
  float* binv = reinterpret_cast<float*>(smem + L.binv);
  float* red = reinterpret_cast<float*>(smem + L.red);

  int b, e, row0, col0;
  region_origin(q, blockIdx.x, b, e, row0, col0);
  const int64_t HW = (int64_t)q.H * q.W;
  const int64_t chan0 = (int64_t)e * q.D;
  const TF* fbase = feat + b * q.bs_f + chan0 * HW;
  const TV* vbase = value + b * q.bs_v + chan0 * HW;
  const TG* gbase = dout + b * q.bs_o + chan0 * HW;
  const float alpha = __ldg(alpha_p);

  setup_points(q, row0, col0, off, mask, binv, cnt);
  if (STAGED) {
    stage_tile<TF>(sf, fbase, q, row0, col0);
    stage_tile<TV>(sv, vbase, q, row0, col0);
    stage_tile<TG>(sg, gbase, q, row0, col0);
  }
  __syncthreads();
  TileAccess<TF> AF{sf, fbase, off, q.NS, HW};
  TileAccess<TV> AV{sv, vbase, off, q.NS, HW};
  TileAccess<TG> AG{sg, gbase, off, q.NS, HW};

  // recompute centres
  bin_means<MAXM, STAGED>(AF, q, mask, binv, chat);
  bin_means<MAXM, STAGED>(AV, q, mask, binv, vc);
  __syncthreads();
  normalise_centres<MAXM>(q, chat, cnorm);
  // saved assignment
  for (int n = threadIdx.x; n < q.N; n += blockDim.x) {
    int64_t io = ((int64_t)b * q.E + e) * HW + off[n];
    int kk = idx_in[io];
    k[n] = (uint8_t)kk;
    g[n] = smax_in[io];
    atomicAdd(&cnt[kk], 1);
  }
  __syncthreads();

  // B2: per-point norm and cosine to the assigned centre
  for (int n = threadIdx.x; n < q.N; n += blockDim.x) {
    int kk = k[n];
    float ss = 0.f, dt = 0.f;
    for (int d = 0; d < q.D; ++d) {
      float x = AF.template at<STAGED>(d, n);
      ss = fmaf(x, x, ss);
      dt = fmaf(chat[d * MAXM + kk], x, dt);
    }
    float nrm = sqrtf(ss);
    float iv = 1.0f / fmaxf(nrm, NORM_EPS);
    inv[n] = (nrm >= NORM_EPS) ? iv : -iv;   // sign flags the clamped branch (no projection term)
    z[n] = dt * iv;
  }
  __syncthreads();

  // B3: A_m and da_m
  assigned_sums<MAXM, STAGED>(AV, q, k, g, agg);
  assigned_sums<MAXM, STAGED>(AG, q, k, g, dA);
  __syncthreads();
  for (int i = threadIdx.x; i < q.D * q.M; i += blockDim.x) {
    int d = i / q.M, m = i % q.M;
    float den = 1.0f / ((float)cnt[m] + 1.0f);
    agg[d * MAXM + m] = (agg[d * MAXM + m] + vc[d * MAXM + m]) * den;
    dA[d * MAXM + m] *= den;
  }
  __syncthreads();

  // B4: dg, t, dz and the scalar gradients
  float t_sum = 0.f, tz_sum = 0.f;
  for (int n = threadIdx.x; n < q.N; n += blockDim.x) {
    int kk = k[n];
    float dg = 0.f;
    for (int d = 0; d < q.D; ++d) {
      dg = fmaf(AG.template at<STAGED>(d, n), agg[d * MAXM + kk], dg);
      dg = fmaf(dA[d * MAXM + kk], AV.template at<STAGED>(d, n), dg);
    }
    float gn = g[n];
    float t = dg * gn * (1.0f - gn);
    t_sum += t;
    tz_sum += t * z[n];
    dz[n] = alpha * t;
  }
  // block reduce (fixed order -> deterministic)
  t_sum = warp_sum(t_sum);
  tz_sum = warp_sum(tz_sum);
  if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = tz_sum; red[32 + (threadIdx.x >> 5)] = t_sum; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, bb = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += red[w]; bb += red[32 + w]; }
    partials[2 * blockIdx.x] = a;       // dalpha contribution
    partials[2 * blockIdx.x + 1] = bb;  // dbeta contribution
  }

  // B5: dchat_m = sum_{k_n=m} dz_n * fhat_n.  First fold the per-point coefficients the last two passes need:
  //   df_n[d] = dz*|inv| * (chat_k[d] - z*|inv|*f[d])      (un-clamped);   df_n[d] = dz*|inv|*chat_k[d]  (clamped)
  // coefA = dz*|inv| -> dz[] (and -> z[] as the weight of f in dchat), coefB = z*|inv| (0 when clamped) -> inv[]
  float* wgt = z;
  __syncthreads();
  for (int n = threadIdx.x; n < q.N; n += blockDim.x) {
    float iv = inv[n];
    float aiv = fabsf(iv);
    float cA = dz[n] * aiv;
    float cB = (iv > 0.f) ? z[n] * aiv : 0.f;
    dz[n] = cA;
    inv[n] = cB;
    wgt[n] = cA;   // dz * |inv|: weight of f in dchat
  }
  __syncthreads();
  assigned_sums<MAXM, STAGED>(AF, q, k, wgt, dch);
  __syncthreads();
  // dc_m = (dchat_m - chat_m <chat_m, dchat_m>) / |c_m|   (clamped: dchat_m / eps), pre-divided by |bin m|
  if (threadIdx.x < q.M) {
    int m = threadIdx.x;
    float nrm = cnorm[m];
    float dt = 0.f;
    for (int d = 0; d < q.D; ++d) dt = fmaf(chat[d * MAXM + m], dch[d * MAXM + m], dt);
    bool clamped = nrm < NORM_EPS;
    float iv = 1.0f / fmaxf(nrm, NORM_EPS);
    for (int d = 0; d < q.D; ++d) {
      float v = dch[d * MAXM + m] - (clamped ? 0.f : chat[d * MAXM + m] * dt);
      dch[d * MAXM + m] = v * iv * binv[m];
    }
  }
  __syncthreads();

  // B6: write dfeat, dvalue
  TDF* dfb = dfeat + b * q.bs_df + chan0 * HW;
  TDV* dvb = dvalue + b * q.bs_dv + chan0 * HW;
  for (int n = threadIdx.x; n < q.N; n += blockDim.x) {
    int kk = k[n];
    float cA = dz[n], cB = inv[n], gn = g[n];
    const uint64_t mk = mask[n];
    int o = off[n];
    for (int d = 0; d < q.D; ++d) {
      float x = AF.template at<STAGED>(d, n);
      float df = cA * (chat[d * MAXM + kk] - cB * x);
      float dv = gn * dA[d * MAXM + kk];
#pragma unroll
      for (int m = 0; m < MAXM; ++m) {
        if ((mk >> m) & 1ull) {
          df += dch[d * MAXM + m];
          dv = fmaf(dA[d * MAXM + m], binv[m], dv);
        }
      }
      stf<TDF>(dfb + (int64_t)d * HW + o, df);
      stf<TDV>(dvb + (int64_t)d * HW + o, dv);
    }
  }
}

__global__ void core_reduce_partials(const float* __restrict__ partials, int n, float* __restrict__ out2) {
  // single CTA, fixed order: deterministic
  __shared__ double sa[256], sb[256];
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) { a += partials[2 * i]; b += partials[2 * i + 1]; }
  sa[threadIdx.x] = a; sb[threadIdx.x] = b;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) { sa[threadIdx.x] += sa[threadIdx.x + s]; sb[threadIdx.x] += sb[threadIdx.x + s]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { out2[0] = (float)sa[0]; out2[1] = (float)sb[0]; }
}

// ---- host side -------------------------------------------------------------------------------------------
static int make_geom(CoreGeom& q, int B, int E, int D, int H, int W, int fold_w, int fold_h, int pw, int ph) {
  VRCOC_REQUIRE(B > 0 && E > 0 && D > 0 && H > 0 && W > 0, "cluster_core: non-positive dimension");
  VRCOC_REQUIRE(pw > 0 && ph > 0 && pw * ph <= 64, "cluster_core: proposal %dx%d unsupported (M must be <= 64)", pw, ph);
  VRCOC_REQUIRE(D <= 256, "cluster_core: head_dim %d > 256 unsupported", D);
  bool folded = fold_w > 1 && fold_h > 1;  // vr_coc.py:160
  q.B = B; q.E = E; q.D = D; q.H = H; q.W = W;
  q.F1 = folded ? fold_w : 1;
  q.F2 = folded ? fold_h : 1;
  // mirrors the reference assert (vr_coc.py:163)
  VRCOC_REQUIRE(H % q.F1 == 0 && W % q.F2 == 0, "Ensure the feature map size (%d*%d) can be divided by fold %d*%d", H, W,
                fold_w, fold_h);
  q.rw = H / q.F1; q.rh = W / q.F2;
  q.pw = pw; q.ph = ph;
  q.N = q.rw * q.rh; q.M = pw * ph;
  VRCOC_REQUIRE(q.rw >= pw && q.rh >= ph, "cluster_core: region %dx%d smaller than proposal %dx%d", q.rw, q.rh, pw, ph);
  VRCOC_REQUIRE(q.N <= 16384, "cluster_core: region of %d points unsupported (max 16384)", q.N);
  int tpd = 1;
  while (tpd * 2 <= 32 && tpd * 2 * D <= CORE_THREADS) tpd *= 2;
  q.TPD = tpd;
  int pad = ((tpd - q.N) % 32 + 32) % 32;   // NS == TPD (mod 32): channels of one warp land in disjoint banks
  q.NS = q.N + pad;
  if (q.NS % 4) q.NS += 4 - q.NS % 4;        // keep float4 staging stores aligned (only matters for odd N)
  return VRCOC_OK;
}

template <typename F>
static int dispatch_dtype(int dt, F&& f) {
  if (dt == VRCOC_F32) return f((float*)nullptr);
  if (dt == VRCOC_BF16) return f((__nv_bfloat16*)nullptr);
  return fail(VRCOC_EINVAL, "unknown dtype %d", dt);
}

constexpr int SMEM_LIMIT = 200 * 1024;

template <typename TF, typename TV, typename TO>
static int launch_fwd(const void* feat, const void* value, void* out, uint8_t* idx, float* smax, const float* alpha,
                      const float* beta, const CoreGeom& q, cudaStream_t st) {
  int maxm = q.M <= 4 ? 4 : (q.M <= 16 ? 16 : 64);
  CoreSmem Ls = core_layout(q, true, false, maxm);
  bool staged = Ls.total <= SMEM_LIMIT;
  CoreSmem L = staged ? Ls : core_layout(q, false, false, maxm);
  VRCOC_REQUIRE(L.total <= SMEM_LIMIT, "cluster_core: region too large for shared memory (%d bytes)", L.total);
  int R = q.B * q.E * q.F1 * q.F2;
#define LAUNCH(MM, ST)                                                                                              \
  do {                                                                                                              \
    auto kern = core_fwd_kernel<TF, TV, TO, MM, ST>;                                                                \
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total);                               \
    kern<<<R, CORE_THREADS, L.total, st>>>((const TF*)feat, (const TV*)value, (TO*)out, idx, smax, alpha, beta, q, L); \
  } while (0)
  if (maxm == 4) { if (staged) LAUNCH(4, true); else LAUNCH(4, false); }
  else if (maxm == 16) { if (staged) LAUNCH(16, true); else LAUNCH(16, false); }
  else { if (staged) LAUNCH(64, true); else LAUNCH(64, false); }        // wide proposals (coc_tiny2: 7x7 = 49 centres, vr_coc.py:734-756)
#undef LAUNCH
  return check_launch("cluster_core_fwd");
}

template <typename TF, typename TV, typename TG, typename TDF, typename TDV>
static int launch_bwd(const void* feat, const void* value, const void* dout, const uint8_t* idx, const float* smax,
                      const float* alpha, void* dfeat, void* dvalue, float* dab, float* partials, const CoreGeom& q,
                      cudaStream_t st) {
  int maxm = q.M <= 4 ? 4 : (q.M <= 16 ? 16 : 64);
  CoreSmem Ls = core_layout(q, true, true, maxm);
  bool staged = Ls.total <= SMEM_LIMIT;
  CoreSmem L = staged ? Ls : core_layout(q, false, true, maxm);
  VRCOC_REQUIRE(L.total <= SMEM_LIMIT, "cluster_core_bwd: region too large for shared memory (%d bytes)", L.total);
  int R = q.B * q.E * q.F1 * q.F2;
#define LAUNCH(MM, ST)                                                                                             \
  do {                                                                                                             \
    auto kern = core_bwd_kernel<TF, TV, TG, TDF, TDV, MM, ST>;                                                     \
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total);                              \
    kern<<<R, CORE_THREADS, L.total, st>>>((const TF*)feat, (const TV*)value, (const TG*)dout, idx, smax, alpha,   \
                                           (TDF*)dfeat, (TDV*)dvalue, partials, q, L);                             \
  } while (0)
  const bool live16 = maxm == 4 && staged && q.rw == 16 && q.rh == 16 && q.pw == 2 && q.ph == 2 && q.NS == 264 && q.TPD == 8;
#define LAUNCH_FIXED(FDV)                                                                                          \
  do {                                                                                                             \
    auto kern = core_bwd_kernel<TF, TV, TG, TDF, TDV, 4, true, FDV>;                                               \
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total);                              \
    kern<<<R, CORE_THREADS, L.total, st>>>((const TF*)feat, (const TV*)value, (const TG*)dout, idx, smax, alpha,   \
                                           (TDF*)dfeat, (TDV*)dvalue, partials, q, L);                             \
  } while (0)
  static const bool fixed_on = []() { const char* k = getenv("VRCOC_CORE_BWD_FIXED"); return !(k && k[0] == '0'); }();   // "0": A/B switch
  if (live16 && fixed_on && q.D == 32) LAUNCH_FIXED(32);
  else if (live16 && fixed_on && q.D == 24) LAUNCH_FIXED(24);
  else if (maxm == 4) { if (staged) LAUNCH(4, true); else LAUNCH(4, false); }
  else if (maxm == 16) { if (staged) LAUNCH(16, true); else LAUNCH(16, false); }
  else { if (staged) LAUNCH(64, true); else LAUNCH(64, false); }
#undef LAUNCH_FIXED
#undef LAUNCH
  int rc = check_launch("cluster_core_bwd");
  if (rc) return rc;
  core_reduce_partials<<<1, 256, 0, st>>>(partials, R, dab);
  return check_launch("cluster_core_bwd.reduce");
}

}  // namespace vrcoc

using namespace vrcoc;

extern "C" int vrcoc_cluster_core_fwd(const void* feat, int feat_dtype, const void* value, int value_dtype, void* out,
                                      int out_dtype, uint8_t* idx, float* sim_max, const float* alpha, const float* beta,
                                      int B, int E, int D, int H, int W, int fold_w, int fold_h, int proposal_w,
                                      int proposal_h, int64_t feat_bstride, int64_t value_bstride, int64_t out_bstride,
                                      void* stream) {
  VRCOC_REQUIRE(feat && value && out && alpha && beta, "cluster_core_fwd: null pointer");
  CoreGeom q;
  int rc = make_geom(q, B, E, D, H, W, fold_w, fold_h, proposal_w, proposal_h);
  if (rc) return rc;
  const int64_t dense = (int64_t)E * D * H * W;
  q.bs_f = feat_bstride ? feat_bstride : dense;
  q.bs_v = value_bstride ? value_bstride : dense;
  q.bs_o = out_bstride ? out_bstride : dense;
  q.bs_df = q.bs_dv = dense;
  cudaStream_t st = (cudaStream_t)stream;
  // compile-time path for 16x16 regions / 2x2 proposals / D in {24, 32}, then the TMA fast path for 2x2 proposals on
  // power-of-two regions (every live configuration); 1 = not covered -> next kernel
  rc = cluster_core_fwd_fast2(feat, feat_dtype, value, value_dtype, out, out_dtype, idx, sim_max, alpha, beta, B, E, D, H, W, q.F1,
                              q.F2, proposal_w, proposal_h, q.bs_f, q.bs_v, q.bs_o, st);
  if (rc != 1) return rc;
  rc = cluster_core_fwd_fast(feat, feat_dtype, value, value_dtype, out, out_dtype, idx, sim_max, alpha, beta, B, E, D, H, W, q.F1,
                             q.F2, proposal_w, proposal_h, q.bs_f, q.bs_v, q.bs_o, st);
  if (rc != 1) return rc;
  // supported storage combinations: all-fp32, or (feat fp32|bf16, value bf16, out bf16)
  if (feat_dtype == VRCOC_F32 && value_dtype == VRCOC_F32 && out_dtype == VRCOC_F32)
    return launch_fwd<float, float, float>(feat, value, out, idx, sim_max, alpha, beta, q, st);
  if (feat_dtype == VRCOC_F32 && value_dtype == VRCOC_BF16 && out_dtype == VRCOC_BF16)
    return launch_fwd<float, __nv_bfloat16, __nv_bfloat16>(feat, value, out, idx, sim_max, alpha, beta, q, st);
  if (feat_dtype == VRCOC_BF16 && value_dtype == VRCOC_BF16 && out_dtype == VRCOC_BF16)
    return launch_fwd<__nv_bfloat16, __nv_bfloat16, __nv_bfloat16>(feat, value, out, idx, sim_max, alpha, beta, q, st);
  return fail(VRCOC_EINVAL, "cluster_core_fwd: unsupported dtype combination feat=%d value=%d out=%d", feat_dtype,
              value_dtype, out_dtype);
}

extern "C" int vrcoc_cluster_core_bwd(const void* feat, int feat_dtype, const void* value, int value_dtype,
                                      const void* dout, int dout_dtype, const uint8_t* idx, const float* sim_max,
                                      const float* alpha, const float* beta, void* dfeat, int dfeat_dtype, void* dvalue,
                                      int dvalue_dtype, float* dalpha_beta, float* partials, int B, int E, int D, int H,
                                      int W, int fold_w, int fold_h, int proposal_w, int proposal_h, int64_t feat_bstride,
                                      int64_t value_bstride, int64_t dout_bstride, int64_t dfeat_bstride,
                                      int64_t dvalue_bstride, void* stream) {
  (void)beta;
  VRCOC_REQUIRE(feat && value && dout && idx && sim_max && alpha && dfeat && dvalue && dalpha_beta && partials,
                "cluster_core_bwd: null pointer");
  CoreGeom q;
  int rc = make_geom(q, B, E, D, H, W, fold_w, fold_h, proposal_w, proposal_h);
  if (rc) return rc;
  const int64_t dense = (int64_t)E * D * H * W;
  q.bs_f = feat_bstride ? feat_bstride : dense;
  q.bs_v = value_bstride ? value_bstride : dense;
  q.bs_o = dout_bstride ? dout_bstride : dense;
  q.bs_df = dfeat_bstride ? dfeat_bstride : dense;
  q.bs_dv = dvalue_bstride ? dvalue_bstride : dense;
  cudaStream_t st = (cudaStream_t)stream;
  if (feat_dtype == VRCOC_F32 && value_dtype == VRCOC_F32 && dout_dtype == VRCOC_F32 && dfeat_dtype == VRCOC_F32 &&
      dvalue_dtype == VRCOC_F32)
    return launch_bwd<float, float, float, float, float>(feat, value, dout, idx, sim_max, alpha, dfeat, dvalue,
                                                         dalpha_beta, partials, q, st);
  if (feat_dtype == VRCOC_F32 && value_dtype == VRCOC_BF16 && dout_dtype == VRCOC_BF16 && dfeat_dtype == VRCOC_F32 &&
      dvalue_dtype == VRCOC_BF16)
    return launch_bwd<float, __nv_bfloat16, __nv_bfloat16, float, __nv_bfloat16>(feat, value, dout, idx, sim_max, alpha,
                                                                                 dfeat, dvalue, dalpha_beta, partials, q, st);
  return fail(VRCOC_EINVAL, "cluster_core_bwd: unsupported dtype combination");
}
