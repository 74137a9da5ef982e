// Cluster core, fast forward path for the live geometry class: 2x2 centre proposal on even regions whose rows are a
// power-of-two multiple of 8 points (every backbone stage: 16x16 regions, D=32; neck: 8x8 / 16x16 / 32x32, D=24).
// Same maths as core_fwd_kernel (cluster_core.cu; reference backbone/fusion/vr_coc.py:158-190), restructured after the
// ncu profile of the generic kernel (profiles/): that kernel is issue-bound at ~5.3 K instructions per warp per
// region-head (integer div/mod tile staging, per-element bin masks, scalar 2-byte stores) and reaches 1 TB/s.
//
//  * persistent CTAs, one region-head per iteration; the feat/value tiles of the NEXT region-head are fetched by TMA
//    (cp.async.bulk.tensor.4d over the NCHW tensor: box = region rows x region cols x D channels) into the other smem
//    stage while the current one is processed; the output tile leaves through a TMA store.  No load/store instruction
//    and no address arithmetic per element.
//  * AdaptiveAvgPool2d((2,2)) on an even region = four quadrant means: the channel-major passes read float4 "items"
//    (4 consecutive points of a row never straddle a quadrant) and need one predicated add per item.
//  * the assignment is stored as a one-hot weight vector w_n = g_n * e_{k_n} (float4 per point), which turns both the
//    aggregation  A_m = sum_n w_n[m] v_n  and the dispatch  o_n = sum_m w_n[m] a_m  into pure FMA streams; the centre
//    aggregates a_m live in registers of the (channel, lane-group) threads between the two.
//  * arg-max is taken on alpha*<c_hat_m, f_n> (sigmoid is monotone, the positive 1/|f_n| is common to all m): one
//    sigmoid per point instead of four.  Exact ties still resolve to the lowest centre index.
#include "tma.cuh"

namespace vrcoc {

constexpr int FAST_THREADS = 256;
constexpr float FAST_EPS = 1e-12f;

struct FastCfg {
  int B, E, D, H, W, F1, F2, rw, rh, N;
  int lcpr;          // log2(items per region row), item = 4 consecutive points
  int lrh;           // log2(rh)
  int items;         // N / 4
  int TPD, dpb;      // lanes per channel, channels per sweep
  int stages;
  int f_bytes, v_bytes, o_bytes;
  int off_f, off_v, off_o, off_w, off_chat, off_misc, off_bar, total;
  int R;
};

template <typename T> __device__ __forceinline__ float4 load4(const T* plane, int item);
template <> __device__ __forceinline__ float4 load4<float>(const float* plane, int item) {
  return *reinterpret_cast<const float4*>(plane + 4 * item);
}
template <> __device__ __forceinline__ float4 load4<__nv_bfloat16>(const __nv_bfloat16* plane, int item) {
  const uint2 raw = *reinterpret_cast<const uint2*>(plane + 4 * item);
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
template <typename T> __device__ __forceinline__ void store4(T* plane, int item, float4 v);
template <> __device__ __forceinline__ void store4<float>(float* plane, int item, float4 v) {
  *reinterpret_cast<float4*>(plane + 4 * item) = v;
}
template <> __device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* plane, int item, float4 v) {
  uint2 raw;
  *reinterpret_cast<__nv_bfloat162*>(&raw.x) = __floats2bfloat162_rn(v.x, v.y);
  *reinterpret_cast<__nv_bfloat162*>(&raw.y) = __floats2bfloat162_rn(v.z, v.w);
  *reinterpret_cast<uint2*>(plane + 4 * item) = raw;
}

__device__ __forceinline__ void add_quadrant(float (&acc)[4], int m, float s) {
  acc[0] += (m == 0) ? s : 0.f;
  acc[1] += (m == 1) ? s : 0.f;
  acc[2] += (m == 2) ? s : 0.f;
  acc[3] += (m == 3) ? s : 0.f;
}

template <typename TF, typename TV, typename TO>
__global__ void __launch_bounds__(FAST_THREADS)
core_fwd_fast_kernel(const __grid_constant__ CUtensorMap tmF, const __grid_constant__ CUtensorMap tmV,
                     const __grid_constant__ CUtensorMap tmO, uint8_t* __restrict__ idx_out, float* __restrict__ smax_out,
                     const float* __restrict__ alpha_p, const float* __restrict__ beta_p, FastCfg G) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t pad = (128u - (smem_u32(smem_raw) & 127u)) & 127u;
  unsigned char* smem = smem_raw + pad;
  float* wq = reinterpret_cast<float*>(smem + G.off_w);            // [4][N] one-hot * g, centre-major: every LDS.128 of
                                                                   // 4 consecutive points is bank-conflict-free across a warp
  float* chat = reinterpret_cast<float*>(smem + G.off_chat);       // [D][4] normalised centres
  float* cnorm = reinterpret_cast<float*>(smem + G.off_misc);      // [4]
  int* cnt = reinterpret_cast<int*>(smem + G.off_misc + 16);       // [4]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + G.off_bar);  // [stages]

  const int tid = threadIdx.x;
  const int sub = tid % G.TPD, dsel = tid / G.TPD;
  const int half_rows = G.rw >> 1, half_items = 1 << (G.lcpr - 1), cpr_mask = (1 << G.lcpr) - 1;
  const float alpha = __ldg(alpha_p), beta = __ldg(beta_p);
  const float inv_quadrant = 4.0f / (float)G.N;
  const int64_t HW = (int64_t)G.H * G.W;

  if (tid == 0) {
    for (int i = 0; i < G.stages; ++i) mbar_init(&bars[i], 1);
    mbar_fence_init();
    tma_prefetch_desc(&tmF); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmO);
  }
  __syncthreads();

  auto issue_load = [&](int r, int stage) {
    const int f2 = r % G.F2, f1 = (r / G.F2) % G.F1, be = r / (G.F1 * G.F2);
    const int e = be % G.E, b = be / G.E;
    mbar_expect_tx(&bars[stage], (uint32_t)(G.f_bytes + G.v_bytes));
    tma_load_4d(smem + G.off_f + stage * G.f_bytes, &tmF, f2 * G.rh, f1 * G.rw, e * G.D, b, &bars[stage]);
    tma_load_4d(smem + G.off_v + stage * G.v_bytes, &tmV, f2 * G.rh, f1 * G.rw, e * G.D, b, &bars[stage]);
  };

  int r = blockIdx.x;
  if (tid == 0 && r < G.R) issue_load(r, 0);

  for (int it = 0; r < G.R; ++it, r += gridDim.x) {
    const int s = G.stages > 1 ? (it & 1) : 0;
    const uint32_t parity = G.stages > 1 ? (uint32_t)(it >> 1) & 1u : (uint32_t)it & 1u;
    if (tid == 0) {
      tma_store_wait_read();                                       // previous output tile has left shared memory
      if (G.stages > 1 && r + (int)gridDim.x < G.R) issue_load(r + gridDim.x, s ^ 1);
    }
    if (tid < 4) cnt[tid] = 0;
    mbar_wait(&bars[s], parity);
    const TF* ft = reinterpret_cast<const TF*>(smem + G.off_f + s * G.f_bytes);
    const TV* vt = reinterpret_cast<const TV*>(smem + G.off_v + s * G.v_bytes);
    // the output tile reuses this stage's feat tile: feat is dead after pass 2 (sizeof(TO) <= sizeof(TF), host-checked)
    TO* otile = reinterpret_cast<TO*>(smem + G.off_f + s * G.f_bytes);

    // ---- pass 1: centre proposal of feat (quadrant means) -----------------------------------------------------------
    for (int d0 = 0; d0 < G.D; d0 += G.dpb) {
      const int d = d0 + dsel;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      if (d < G.D) {
        const TF* plane = ft + d * G.N;
#pragma unroll 4
        for (int i = sub; i < G.items; i += G.TPD) {
          const float4 x = load4<TF>(plane, i);
          const int m = (((i >> G.lcpr) >= half_rows) ? 2 : 0) | (((i & cpr_mask) >= half_items) ? 1 : 0);
          add_quadrant(acc, m, (x.x + x.y) + (x.z + x.w));
        }
      }
#pragma unroll
      for (int m = 0; m < 4; ++m)
        for (int o = G.TPD >> 1; o > 0; o >>= 1) acc[m] += __shfl_xor_sync(0xffffffffu, acc[m], o);
      if (d < G.D && sub == 0)
        *reinterpret_cast<float4*>(chat + 4 * d) = make_float4(acc[0] * inv_quadrant, acc[1] * inv_quadrant, acc[2] * inv_quadrant, acc[3] * inv_quadrant);
    }
    __syncthreads();
    if (tid < 4) {
      float ss = 0.f;
      for (int d = 0; d < G.D; ++d) { const float c = chat[4 * d + tid]; ss = fmaf(c, c, ss); }
      const float nrm = sqrtf(ss);
      cnorm[tid] = nrm;
      const float inv = 1.0f / fmaxf(nrm, FAST_EPS);
      for (int d = 0; d < G.D; ++d) chat[4 * d + tid] *= inv;
    }
    __syncthreads();

    // ---- pass 2: similarity, arg-max, gate -> one-hot weights ---------------------------------------------------------
    const int f2 = r % G.F2, f1 = (r / G.F2) % G.F1, be = r / (G.F1 * G.F2);
    for (int n0 = 0; n0 < G.N; n0 += FAST_THREADS) {
      const int n = n0 + tid;
      const bool in = n < G.N;
      int kbest = 0;
      if (in) {
        float ss = 0.f, d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll 8
        for (int d = 0; d < G.D; ++d) {
          const float x = (float)ft[d * G.N + n];
          const float4 c = *reinterpret_cast<const float4*>(chat + 4 * d);
          ss = fmaf(x, x, ss);
          d0 = fmaf(c.x, x, d0); d1 = fmaf(c.y, x, d1); d2 = fmaf(c.z, x, d2); d3 = fmaf(c.w, x, d3);
        }
        const float inv = 1.0f / fmaxf(sqrtf(ss), FAST_EPS);
        float tb = alpha * d0, db = d0;
        if (alpha * d1 > tb) { tb = alpha * d1; db = d1; kbest = 1; }
        if (alpha * d2 > tb) { tb = alpha * d2; db = d2; kbest = 2; }
        if (alpha * d3 > tb) { tb = alpha * d3; db = d3; kbest = 3; }
        const float g = sigmoidf_exact(fmaf(alpha, db * inv, beta));
        wq[n] = kbest == 0 ? g : 0.f;
        wq[G.N + n] = kbest == 1 ? g : 0.f;
        wq[2 * G.N + n] = kbest == 2 ? g : 0.f;
        wq[3 * G.N + n] = kbest == 3 ? g : 0.f;
        if (idx_out || smax_out) {
          const int64_t io = (int64_t)be * HW + (int64_t)(f1 * G.rw + (n >> G.lrh)) * G.W + f2 * G.rh + (n & (G.rh - 1));
          if (idx_out) idx_out[io] = (uint8_t)kbest;
          if (smax_out) smax_out[io] = g;
        }
      }
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const int c = __popc(__ballot_sync(0xffffffffu, in && kbest == m));
        if ((tid & 31) == 0 && c) atomicAdd(&cnt[m], c);
      }
    }
    __syncthreads();

    // ---- pass 3 + 4: aggregate value to the centres, dispatch back to the points ----------------------------------------
    const float den0 = 1.0f / ((float)cnt[0] + 1.0f), den1 = 1.0f / ((float)cnt[1] + 1.0f);
    const float den2 = 1.0f / ((float)cnt[2] + 1.0f), den3 = 1.0f / ((float)cnt[3] + 1.0f);
    for (int d0 = 0; d0 < G.D; d0 += G.dpb) {
      const int d = d0 + dsel;
      float A[4] = {0.f, 0.f, 0.f, 0.f}, Q[4] = {0.f, 0.f, 0.f, 0.f};
      if (d < G.D) {
        const TV* plane = vt + d * G.N;
#pragma unroll 4
        for (int i = sub; i < G.items; i += G.TPD) {
          const float4 v = load4<TV>(plane, i);
          const float4 w0 = *reinterpret_cast<const float4*>(wq + 4 * i), w1 = *reinterpret_cast<const float4*>(wq + G.N + 4 * i);
          const float4 w2 = *reinterpret_cast<const float4*>(wq + 2 * G.N + 4 * i), w3 = *reinterpret_cast<const float4*>(wq + 3 * G.N + 4 * i);
          A[0] = fmaf(w0.x, v.x, fmaf(w0.y, v.y, fmaf(w0.z, v.z, fmaf(w0.w, v.w, A[0]))));
          A[1] = fmaf(w1.x, v.x, fmaf(w1.y, v.y, fmaf(w1.z, v.z, fmaf(w1.w, v.w, A[1]))));
          A[2] = fmaf(w2.x, v.x, fmaf(w2.y, v.y, fmaf(w2.z, v.z, fmaf(w2.w, v.w, A[2]))));
          A[3] = fmaf(w3.x, v.x, fmaf(w3.y, v.y, fmaf(w3.z, v.z, fmaf(w3.w, v.w, A[3]))));
          const int m = (((i >> G.lcpr) >= half_rows) ? 2 : 0) | (((i & cpr_mask) >= half_items) ? 1 : 0);
          add_quadrant(Q, m, (v.x + v.y) + (v.z + v.w));
        }
      }
#pragma unroll
      for (int m = 0; m < 4; ++m)
        for (int o = G.TPD >> 1; o > 0; o >>= 1) {
          A[m] += __shfl_xor_sync(0xffffffffu, A[m], o);
          Q[m] += __shfl_xor_sync(0xffffffffu, Q[m], o);
        }
      if (d < G.D) {
        const float a0 = fmaf(Q[0], inv_quadrant, A[0]) * den0, a1 = fmaf(Q[1], inv_quadrant, A[1]) * den1;
        const float a2 = fmaf(Q[2], inv_quadrant, A[2]) * den2, a3 = fmaf(Q[3], inv_quadrant, A[3]) * den3;
        TO* oplane = otile + d * G.N;
#pragma unroll 4
        for (int i = sub; i < G.items; i += G.TPD) {
          const float4 w0 = *reinterpret_cast<const float4*>(wq + 4 * i), w1 = *reinterpret_cast<const float4*>(wq + G.N + 4 * i);
          const float4 w2 = *reinterpret_cast<const float4*>(wq + 2 * G.N + 4 * i), w3 = *reinterpret_cast<const float4*>(wq + 3 * G.N + 4 * i);
          float4 o;
          o.x = fmaf(w0.x, a0, fmaf(w1.x, a1, fmaf(w2.x, a2, w3.x * a3)));
          o.y = fmaf(w0.y, a0, fmaf(w1.y, a1, fmaf(w2.y, a2, w3.y * a3)));
          o.z = fmaf(w0.z, a0, fmaf(w1.z, a1, fmaf(w2.z, a2, w3.z * a3)));
          o.w = fmaf(w0.w, a0, fmaf(w1.w, a1, fmaf(w2.w, a2, w3.w * a3)));
          store4<TO>(oplane, i, o);
        }
      }
    }
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      const int e = be % G.E, b = be / G.E;
      tma_store_4d(&tmO, otile, f2 * G.rh, f1 * G.rw, e * G.D, b);
      tma_store_commit();
      if (G.stages == 1 && r + (int)gridDim.x < G.R) {     // single stage: the tiles are free once the store has read them
        tma_store_wait_read();
        issue_load(r + gridDim.x, 0);
      }
    }
  }
  if (tid == 0) tma_store_wait_all();
}

static int ilog2_exact(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return (1 << l) == v ? l : -1;
}

static int esz(int dt) { return dt == VRCOC_F32 ? 4 : 2; }

template <typename TF, typename TV, typename TO>
static int launch_fast(const CUtensorMap& tf, const CUtensorMap& tv, const CUtensorMap& to, uint8_t* idx, float* smax,
                       const float* alpha, const float* beta, const FastCfg& G, int grid, cudaStream_t st) {
  auto kern = core_fwd_fast_kernel<TF, TV, TO>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G.total);
  kern<<<grid, FAST_THREADS, G.total, st>>>(tf, tv, to, idx, smax, alpha, beta, G);
  return check_launch("cluster_core_fwd_fast");
}

// returns VRCOC_OK when launched, 1 when the geometry is not covered by the fast path (caller falls back), <0 on error
int cluster_core_fwd_fast(const void* feat, int fdt, const void* value, int vdt, void* out, int odt, uint8_t* idx, float* smax,
                          const float* alpha, const float* beta, int B, int E, int D, int H, int W, int F1, int F2, int pw, int ph,
                          int64_t bs_f, int64_t bs_v, int64_t bs_o, cudaStream_t st) {
  if (pw != 2 || ph != 2 || tma_encode_fn() == nullptr) return 1;
  FastCfg G{};
  G.B = B; G.E = E; G.D = D; G.H = H; G.W = W; G.F1 = F1; G.F2 = F2;
  G.rw = H / F1; G.rh = W / F2; G.N = G.rw * G.rh;
  G.lrh = ilog2_exact(G.rh);
  if ((G.rw & 1) || (G.rh % 8) || G.lrh < 0 || G.rh > 256 || G.rw > 256 || D > 256) return 1;
  G.lcpr = G.lrh - 2;
  G.items = G.N / 4;
  int tpd = 1;
  while (tpd * 2 <= 32 && tpd * 2 * D <= FAST_THREADS) tpd *= 2;
  G.TPD = tpd; G.dpb = FAST_THREADS / tpd;
  G.R = B * E * F1 * F2;
  // TMA constraints: 16-byte aligned bases and global strides
  auto ok = [&](const void* p, int dt, int64_t bs) {
    const int es = esz(dt);
    return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && ((int64_t)W * es) % 16 == 0 && ((int64_t)H * W * es) % 16 == 0 && (bs * es) % 16 == 0;
  };
  if (!ok(feat, fdt, bs_f) || !ok(value, vdt, bs_v) || !ok(out, odt, bs_o)) return 1;
  auto r128 = [](int v) { return (v + 127) & ~127; };
  G.f_bytes = r128(D * G.N * esz(fdt)); G.v_bytes = r128(D * G.N * esz(vdt)); G.o_bytes = r128(D * G.N * esz(odt));
  if (G.f_bytes != D * G.N * esz(fdt) || G.v_bytes != D * G.N * esz(vdt)) return 1;   // tiles must be whole 128-byte units
  if (G.o_bytes > G.f_bytes) return 1;                                                 // the output tile aliases the feat tile
  const int fixed = r128(G.N * 16) + r128(D * 16) + 128 + 128 + 256;
  // small tiles: single stage and four resident CTAs per SM (the loads of one CTA overlap the passes of the others and
  // 32 warps hide the shared-memory latency of the FMA streams); larger tiles: two stages inside one CTA
  const int one = G.f_bytes + G.v_bytes + fixed;
  G.stages = (one <= 56 * 1024) ? 1 : ((2 * (G.f_bytes + G.v_bytes) + fixed <= 200 * 1024) ? 2 : 1);
  if (G.stages * (G.f_bytes + G.v_bytes) + fixed > 220 * 1024) return 1;
  int p = 0;
  G.off_f = p; p += G.stages * G.f_bytes;
  G.off_v = p; p += G.stages * G.v_bytes;
  G.off_o = 0;
  G.off_w = p; p += r128(G.N * 16);
  G.off_chat = p; p += r128(D * 16);
  G.off_misc = p; p += 128;
  G.off_bar = p; p += 128;
  G.total = p + 128;   // alignment slack

  CUtensorMap tf, tv, to;
  auto enc = [&](CUtensorMap* tm, const void* base, int dt, int64_t bs) {
    const int es = esz(dt);
    cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)(E * D), (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)W * es, (cuuint64_t)H * W * es, (cuuint64_t)bs * es};
    cuuint32_t box[4] = {(cuuint32_t)G.rh, (cuuint32_t)G.rw, (cuuint32_t)D, 1};
    return tma_encode(tm, dt, base, 4, dims, strides, box, false);
  };
  int rc;
  if ((rc = enc(&tf, feat, fdt, bs_f)) || (rc = enc(&tv, value, vdt, bs_v)) || (rc = enc(&to, out, odt, bs_o))) return rc;
  int per_sm = (220 * 1024) / G.total;
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 4) per_sm = 4;
  int grid = sm_count() * per_sm;
  if (grid > G.R) grid = G.R;
  if (fdt == VRCOC_F32 && vdt == VRCOC_F32 && odt == VRCOC_F32)
    return launch_fast<float, float, float>(tf, tv, to, idx, smax, alpha, beta, G, grid, st);
  if (fdt == VRCOC_F32 && vdt == VRCOC_BF16 && odt == VRCOC_BF16)
    return launch_fast<float, __nv_bfloat16, __nv_bfloat16>(tf, tv, to, idx, smax, alpha, beta, G, grid, st);
  if (fdt == VRCOC_BF16 && vdt == VRCOC_BF16 && odt == VRCOC_BF16)
    return launch_fast<__nv_bfloat16, __nv_bfloat16, __nv_bfloat16>(tf, tv, to, idx, smax, alpha, beta, G, grid, st);
  return 1;
}

}  // namespace vrcoc
