// Statistics, table builders and the bandwidth-bound elementwise passes of the fusion modules.
#include <stdarg.h>

#include "common.cuh"

namespace vrcoc {

// ---- error plumbing ------------------------------------------------------------------------------------------
char* err_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}
int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}
int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(VRCOC_ECUDA, "%s: %s", what, cudaGetErrorString(e));
  return VRCOC_OK;
}

// ---- block reduction of two floats -----------------------------------------------------------------------------
__device__ __forceinline__ void block_sum2(float& a, float& b, float* red /*>=64 floats*/) {
  a = warp_sum(a);
  b = warp_sum(b);
  const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) { red[w] = a; red[32 + w] = b; }
  __syncthreads();
  if (threadIdx.x < 32) {
    float x = (int)threadIdx.x < nw ? red[threadIdx.x] : 0.f;
    float y = (int)threadIdx.x < nw ? red[32 + threadIdx.x] : 0.f;
    x = warp_sum(x);
    y = warp_sum(y);
    if (threadIdx.x == 0) { red[0] = x; red[32] = y; }
  }
  __syncthreads();
  a = red[0];
  b = red[32];
}

// Plane kernels below use grid = (chunks, planes): planes with many points (the 512x512 ingest maps have only B*3..B*4
// planes) are split into PLANE_CHUNK-element chunks so the grid fills the chip; per-plane sums then go through float
// atomics into a zeroed buffer (single-chunk planes store directly and stay bit-reproducible).
constexpr int PLANE_CHUNK = 16384;

template <typename T, typename F>
__device__ __forceinline__ void for_chunk8(const T* p, int lo, int hi, F&& f) {
  // f(index, value[8], count): 8-wide when the chunk allows it
  const bool vec = ((hi - lo) % 8 == 0) && (lo % 8 == 0) && ((reinterpret_cast<uintptr_t>(p) & 15) == 0);
  if (vec) {
    for (int i = lo + threadIdx.x * 8; i < hi; i += blockDim.x * 8) {
      float v[8];
      ld8<T>(p + i, v);
      f(i, v, 8);
    }
  } else {
    for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
      float v[8];
      v[0] = ldf<T>(p + i);
      f(i, v, 1);
    }
  }
}

__device__ __forceinline__ void publish_sums(float s, float s2, float* cs, double* ss, int plane, int C, float* red) {
  block_sum2(s, s2, red);
  if (threadIdx.x == 0) {
    if (cs) {
      if (gridDim.x == 1) { cs[2 * plane] = s; cs[2 * plane + 1] = s2; }
      else { atomicAdd(&cs[2 * plane], s); atomicAdd(&cs[2 * plane + 1], s2); }
    }
    if (ss) {
      double* dst = ss + ((int64_t)(plane / C) * VRCOC_STAT_SLOTS + ((plane + 5 * blockIdx.x) & (VRCOC_STAT_SLOTS - 1))) * 2;
      atomicAdd(dst, (double)s);
      atomicAdd(dst + 1, (double)s2);
    }
  }
}

// ---- per-(sample, channel) sums --------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) channel_sums_kernel(const T* __restrict__ x, int HW, float* __restrict__ cs,
                                                           double* __restrict__ ss, int C) {
  __shared__ float red[64];
  const int plane = blockIdx.y;
  const int lo = blockIdx.x * PLANE_CHUNK, hi = min(HW, lo + PLANE_CHUNK);
  const T* p = x + (int64_t)plane * HW;
  float s = 0.f, s2 = 0.f;
  for_chunk8<T>(p, lo, hi, [&](int, const float (&v)[8], int n) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (j < n) { s += v[j]; s2 = fmaf(v[j], v[j], s2); }
  });
  publish_sums(s, s2, cs, ss, plane, C, red);
}

// ---- y = act(x*s1+t1) + res; y = y*s2+t2, with optional per-plane sums and global min/max ------------------------
template <typename T>
__global__ void __launch_bounds__(256)
chan_affine_kernel(const T* __restrict__ x, const T* __restrict__ res, T* __restrict__ out, const float* __restrict__ s1,
                   const float* __restrict__ t1, int act, const float* __restrict__ s2, const float* __restrict__ t2, int C,
                   int HW, float* __restrict__ cs, uint32_t* __restrict__ minmax) {
  __shared__ float red[64];
  const int plane = blockIdx.y, c = plane % C;
  const int lo = blockIdx.x * PLANE_CHUNK, hi = min(HW, lo + PLANE_CHUNK);
  const float a1 = s1 ? s1[c] : 1.f, b1 = t1 ? t1[c] : 0.f, a2 = s2 ? s2[c] : 1.f, b2 = t2 ? t2[c] : 0.f;
  const int64_t base = (int64_t)plane * HW;
  float s = 0.f, sq = 0.f, mx = 0.f, mn = __int_as_float(0x7f800000);
  const bool rvec = !res || ((reinterpret_cast<uintptr_t>(res + base) & 15) == 0);
  const bool ovec = (reinterpret_cast<uintptr_t>(out + base) & 15) == 0;
  for_chunk8<T>(x + base, lo, hi, [&](int i, const float (&v)[8], int n) {
    float y[8], r[8];
    if (res) {
      if (n == 8 && rvec) ld8<T>(res + base + i, r);
      else r[0] = ldf<T>(res + base + i);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (j < n) {
        float t = apply_act(fmaf(v[j], a1, b1), act);
        if (res) t += r[j];
        t = fmaf(t, a2, b2);
        y[j] = t;
        s += t; sq = fmaf(t, t, sq);
        mx = fmaxf(mx, t); mn = fminf(mn, t);
      }
    }
    if (n == 8 && ovec) st8<T>(out + base + i, y);
    else for (int j = 0; j < n; ++j) stf<T>(out + base + i + j, y[j]);
  });
  if (cs) publish_sums(s, sq, cs, nullptr, plane, C, red);
  if (minmax) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    }
    if ((threadIdx.x & 31) == 0) {
      atomicMax(&minmax[0], __float_as_uint(mx));
      atomicMax(&minmax[1], ~__float_as_uint(mn));
    }
  }
}

// ---- ImageEnhanceByRadar tail -------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
img_enh_finish_kernel(const T* __restrict__ k, const T* __restrict__ img, T* __restrict__ out,
                      const uint32_t* __restrict__ minmax, const float* __restrict__ sc, const float* __restrict__ sh, int C,
                      int HW, float* __restrict__ cs) {
  __shared__ float red[64];
  const int plane = blockIdx.y, c = plane % C;
  const int lo = blockIdx.x * PLANE_CHUNK, hi = min(HW, lo + PLANE_CHUNK);
  const float mx = __uint_as_float(minmax[0]), mn = __uint_as_float(~minmax[1]);
  const float inv_dst = 1.0f / (mx - mn);  // constant map: 0 * inf -> NaN, like the reference's 0/0 true_divide (vr_coc.py:66)
  const float a = sc ? sc[c] : 1.f, b = sh ? sh[c] : 0.f;
  const int64_t base = (int64_t)plane * HW;
  float s = 0.f, sq = 0.f;
  const bool ivec = (reinterpret_cast<uintptr_t>(img + base) & 15) == 0 && (reinterpret_cast<uintptr_t>(out + base) & 15) == 0;
  for_chunk8<T>(k + base, lo, hi, [&](int i, const float (&v)[8], int n) {
    float y[8], im[8];
    if (n == 8 && ivec) ld8<T>(img + base + i, im);
    else im[0] = ldf<T>(img + base + i);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (j < n) {
        float t = fmaf(v[j] - mn, inv_dst, 1.0f) * im[j];
        t = fmaf(t, a, b);
        y[j] = t;
        s += t; sq = fmaf(t, t, sq);
      }
    }
    if (n == 8 && ivec) st8<T>(out + base + i, y);
    else for (int j = 0; j < n; ++j) stf<T>(out + base + i + j, y[j]);
  });
  if (cs) publish_sums(s, sq, cs, nullptr, plane, C, red);
}

// ---- ShuffleAttention parameters + attended-channel means (shuffle_attention.py:48-66) -------------------------------
// image channel c = g*(2q) + half*q + j.  attn[b][c] = {scale, gate_a, gate_c, mean_hw(attended channel)}
//   half 0: attended = x * sigmoid(cw_j*mean + cb_j)                -> scale = that sigmoid, gate off (a=0, c=+88)
//   half 1: attended = x * sigmoid(sw_j*(gw_j*(x-mu)*rstd + gb_j) + sb_j) -> scale = 1, gate_a = sw*gw*rstd, gate_c = sw*(gb - gw*mu*rstd)+sb
// G == 0: no attention (RadarEnhanceByImage initial=True): scale 1, gate off, mean = channel mean.
template <typename T>
__global__ void __launch_bounds__(256)
sa_gate_sums_kernel(const T* __restrict__ img, int Ci, int HW, int G, const float* __restrict__ cs,
                    const float* __restrict__ cw, const float* __restrict__ cb, const float* __restrict__ sw,
                    const float* __restrict__ sb, const float* __restrict__ gw, const float* __restrict__ gb,
                    float* __restrict__ attn, const T* __restrict__ radar, int Cr, float* __restrict__ cs_radar, int n_img_planes) {
  __shared__ float red[64];
  if ((int)blockIdx.x >= n_img_planes) {
    // extra blocks (vrcoc_fusion_stats): plane sums of the radar map for the ECA means, in the same launch
    const int rp = blockIdx.x - n_img_planes;
    const T* p = radar + (int64_t)rp * HW;
    float s = 0.f, s2 = 0.f;
    for_chunk8<T>(p, 0, HW, [&](int, const float (&v)[8], int n) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < n) { s += v[j]; s2 = fmaf(v[j], v[j], s2); }
    });
    block_sum2(s, s2, red);
    if (threadIdx.x == 0) { cs_radar[2 * rp] = s; cs_radar[2 * rp + 1] = s2; }
    return;
  }
  const int plane = blockIdx.x, c = plane % Ci;
  const float inv_hw = 1.0f / (float)HW;
  float psum, psq;
  if (cs) {
    psum = cs[2 * plane];
    psq = cs[2 * plane + 1];
  } else {
    // no precomputed sums: first pass over the block's own plane (it stays in L1/L2 for the gated pass below)
    const T* p = img + (int64_t)plane * HW;
    float s = 0.f, s2 = 0.f;
    for_chunk8<T>(p, 0, HW, [&](int, const float (&v)[8], int n) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < n) { s += v[j]; s2 = fmaf(v[j], v[j], s2); }
    });
    block_sum2(s, s2, red);
    psum = s;
    psq = s2;
    __syncthreads();                                                   // red is reused by the gated sum
  }
  const float mean = psum * inv_hw;
  float scale = 1.f, ga = 0.f, gc = 88.f, amean = mean;
  if (G > 0) {
    const int q = Ci / (2 * G);
    const int half = (c / q) & 1, j = c % q;
    if (half == 0) {
      scale = sigmoidf_exact(fmaf(cw[j], mean, cb[j]));
      amean = mean * scale;
    } else {
      float var = fmaxf(psq * inv_hw - mean * mean, 0.f);
      float rstd = rsqrtf(var + 1e-5f);
      ga = sw[j] * gw[j] * rstd;
      gc = fmaf(sw[j], gb[j] - gw[j] * mean * rstd, sb[j]);
      const T* p = img + (int64_t)plane * HW;
      float s = 0.f, dummy = 0.f;
      if ((HW & 7) == 0 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
        // 128-bit loads; MUFU sigmoid (ex2 + rcp, ~3 ulp): this sum only feeds the ECA channel mean
        for (int i = threadIdx.x * 8; i < HW; i += blockDim.x * 8) {
          float v[8];
          ld8<T>(p + i, v);
#pragma unroll
          for (int e = 0; e < 8; ++e) s = fmaf(v[e], __fdividef(1.0f, 1.0f + __expf(-fmaf(ga, v[e], gc))), s);
        }
      } else {
        for (int i = threadIdx.x; i < HW; i += blockDim.x) {
          float x = ldf<T>(p + i);
          s = fmaf(x, sigmoidf_exact(fmaf(ga, x, gc)), s);
        }
      }
      block_sum2(s, dummy, red);
      amean = s * inv_hw;
    }
  }
  if (threadIdx.x == 0) {
    float4 o = make_float4(scale, ga, gc, amean);
    reinterpret_cast<float4*>(attn)[plane] = o;
  }
}

// ---- ECA over the shuffled concat + final conv prologue table (vr_coc.py:346-350, eca.py:16-22) -----------------------
// logical channel k of z = shuffle_channels(cat[image_attn, radar]) reads concat channel chan_src[k]
// (< Ci: attended image channel, else radar channel).  table[b][k] = {scale*eca, 0, gate_a, gate_c}
__global__ void radar_enh_table_kernel(const float* __restrict__ attn, const float* __restrict__ cs_radar,
                                       const int32_t* __restrict__ chan_src, const float* __restrict__ eca_w, int eca_k, int Ci,
                                       int Cr, int HW, float* __restrict__ table, int concat_order) {
  extern __shared__ float means[];   // [Ci+Cr] logical-channel means
  const int b = blockIdx.x, K = Ci + Cr;
  const float inv_hw = 1.0f / (float)HW;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    int s = chan_src ? chan_src[k] : k;
    means[k] = s < Ci ? attn[((int64_t)b * Ci + s) * 4 + 3] : cs_radar[((int64_t)b * Cr + (s - Ci)) * 2] * inv_hw;
  }
  __syncthreads();
  const int half = (eca_k - 1) / 2;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    float a = 0.f;
    for (int j = 0; j < eca_k; ++j) {
      int kk = k + j - half;
      if (kk >= 0 && kk < K) a = fmaf(eca_w[j], means[kk], a);
    }
    float eca = sigmoidf_exact(a);
    int s = chan_src ? chan_src[k] : k;
    float4 o;
    if (s < Ci) {
      const float* at = attn + ((int64_t)b * Ci + s) * 4;
      o = make_float4(at[0] * eca, 0.f, at[1], at[2]);
    } else {
      o = make_float4(eca, 0.f, 0.f, 88.f);
    }
    // concat_order: row = channel of the virtual concat [image | radar] (the GEMM then runs on the sources in memory order
    // with its weight columns permuted instead, and both sources can come in by TMA)
    reinterpret_cast<float4*>(table)[(int64_t)b * K + (concat_order ? s : k)] = o;
  }
}


// ---- backward elementwise passes of the projections ------------------------------------------------------------------
// dU = dy * gelu'(u),  gelu'(u) = Phi(u) + u*phi(u)     (Mlp.act = nn.GELU(), vr_coc.py:206,219)
template <typename TY, typename TU, typename TO>
__global__ void __launch_bounds__(256)
gelu_bwd_kernel(const TY* __restrict__ dy, const TU* __restrict__ u, TO* __restrict__ out, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    float x = ldf<TU>(u + i);
    float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
    float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
    stf<TO>(out + i, ldf<TY>(dy + i) * fmaf(x, pdf, cdf));
  }
}

// The same pass, 8 elements per thread (128-bit loads / stores; the scalar kernel moves 2-byte words: 166 us for the 402 MB of a
// stage-1 hidden map, 2.4 TB/s).  bf16: Phi and phi from ONE exponential each — Phi(u) = [u >= 0] +- 2^P(|u|) with the degree-6
// exponent polynomial (tools/fit_gelu.py with DEG = 6; the pass is memory-bound, the three extra FMAs are free),
// phi(u) = 2^(-u^2 log2(e) / 2) / sqrt(2 pi); fp32 keeps erff / expf.
__device__ __forceinline__ float gelu_grad_fast(float x) {
  const float a = fminf(fabsf(x), 6.081118f);
  float p = fmaf(3.309320891e-05f, a, -7.692232612e-04f);
  p = fmaf(p, a, 8.080728352e-03f);
  p = fmaf(p, a, -5.341212451e-02f);
  p = fmaf(p, a, -4.587709606e-01f);
  p = fmaf(p, a, -1.151201725e+00f);
  p = fmaf(p, a, -9.999930859e-01f);
  float e, g;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(p));                   // 0.5 erfc(a / sqrt 2)
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(g) : "f"(x * x * -0.72134752044448170368f));
  const float cdf = x >= 0.f ? 1.0f - e : e;
  return fmaf(x * 0.39894228040143267794f, g, cdf);
}
template <typename T>
__global__ void __launch_bounds__(256) gelu_bwd_vec_kernel(const T* __restrict__ dy, const T* __restrict__ u, T* __restrict__ out, int64_t n8) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n8; i += stride) {
    float a[8], x[8], r[8];
    ld8<T>(dy + i * 8, a);
    ld8<T>(u + i * 8, x);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      if (sizeof(T) == 2) {
        r[e] = a[e] * gelu_grad_fast(x[e]);
      } else {
        const float cdf = 0.5f * (1.0f + erff(x[e] * 0.70710678118654752440f));
        const float pdf = 0.39894228040143267794f * expf(-0.5f * x[e] * x[e]);
        r[e] = a[e] * fmaf(x[e], pdf, cdf);
      }
    }
    st8<T>(out + i * 8, r);
  }
}

// GroupNorm(1,C) backward statistics: per-(b,c) {sum_p dz, sum_p dz*x}
template <typename TZ, typename TX>
__global__ void __launch_bounds__(256)
gn_bwd_sums_kernel(const TZ* __restrict__ dz, const TX* __restrict__ x, int HW, float* __restrict__ out) {
  __shared__ float red[64];
  const int plane = blockIdx.x;
  const int64_t base = (int64_t)plane * HW;
  float s = 0.f, sx = 0.f;
  const bool vec = sizeof(TZ) == sizeof(TX) && (HW & 7) == 0 &&
                   ((reinterpret_cast<uintptr_t>(dz + base) | reinterpret_cast<uintptr_t>(x + base)) & (8 * sizeof(TZ) - 1)) == 0;
  if (vec) {                                                         // 128-bit loads (the scalar loop ran at < 1 TB/s)
    for (int i = threadIdx.x * 8; i < HW; i += blockDim.x * 8) {
      float d[8], v[8];
      ld8<TZ>(dz + base + i, d);
      ld8<TX>(x + base + i, v);
#pragma unroll
      for (int e = 0; e < 8; ++e) { s += d[e]; sx = fmaf(d[e], v[e], sx); }
    }
  } else {
    for (int i = threadIdx.x; i < HW; i += blockDim.x) {
      float d = ldf<TZ>(dz + base + i);
      s += d; sx = fmaf(d, ldf<TX>(x + base + i), sx);
    }
  }
  block_sum2(s, sx, red);
  if (threadIdx.x == 0) { out[2 * plane] = s; out[2 * plane + 1] = sx; }
}

// dx = dz*a[b,c] + x*bb[b] + cc[b] (+ extra)
template <typename TZ, typename TX, typename TO>
__global__ void __launch_bounds__(256)
gn_bwd_apply_kernel(const TZ* __restrict__ dz, const TX* __restrict__ x, const TO* __restrict__ extra, TO* __restrict__ out,
                    const float* __restrict__ a, const float* __restrict__ bb, const float* __restrict__ cc, int C, int HW) {
  const int plane = blockIdx.x, b = plane / C;
  const float ka = a[plane], kb = bb[b], kc = cc[b];
  const int64_t base = (int64_t)plane * HW;
  const bool vec = sizeof(TZ) == sizeof(TX) && sizeof(TZ) == sizeof(TO) && (HW & 7) == 0 &&
                   ((reinterpret_cast<uintptr_t>(dz + base) | reinterpret_cast<uintptr_t>(x + base) | reinterpret_cast<uintptr_t>(out + base) |
                     reinterpret_cast<uintptr_t>(extra ? extra + base : out + base)) & (8 * sizeof(TZ) - 1)) == 0;
  if (vec) {
    for (int i = threadIdx.x * 8; i < HW; i += blockDim.x * 8) {
      float d[8], v[8], r[8];
      ld8<TZ>(dz + base + i, d);
      ld8<TX>(x + base + i, v);
#pragma unroll
      for (int e = 0; e < 8; ++e) r[e] = fmaf(d[e], ka, fmaf(v[e], kb, kc));
      if (extra) {
        float ex[8];
        ld8<TO>(extra + base + i, ex);
#pragma unroll
        for (int e = 0; e < 8; ++e) r[e] += ex[e];
      }
      st8<TO>(out + base + i, r);
    }
    return;
  }
  for (int i = threadIdx.x; i < HW; i += blockDim.x) {
    float v = fmaf(ldf<TZ>(dz + base + i), ka, fmaf(ldf<TX>(x + base + i), kb, kc));
    if (extra) v += ldf<TO>(extra + base + i);
    stf<TO>(out + base + i, v);
  }
}


// ---- bilinear upsample, align_corners=True (nn.Upsample in CoCUpsample, reference neck/coc_fpn_dual.py:21) ---------------
// out[y][x] = lerp over the 2x2 source neighbourhood at (y*(H-1)/(Ho-1), x*(W-1)/(Wo-1)); one thread = 8 consecutive
// output columns of one (plane, row): 128-bit stores, source rows served from L1/L2.
template <typename T>
__global__ void __launch_bounds__(256)
upsample_bilinear_kernel(const T* __restrict__ x, T* __restrict__ out, int planes, int H, int W, int Ho, int Wo, float sy, float sx) {
  const int cols8 = (Wo + 7) >> 3;
  const int64_t total = (int64_t)planes * Ho * cols8;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(t % cols8);
    const int64_t rowid = t / cols8;
    const int oy = (int)(rowid % Ho);
    const int64_t plane = rowid / Ho;
    const float fy = oy * sy;
    int y0 = (int)fy;
    if (y0 > H - 1) y0 = H - 1;
    const int y1 = min(y0 + 1, H - 1);
    const float wy = fy - (float)y0;
    const T* r0 = x + (plane * H + y0) * W;
    const T* r1 = x + (plane * H + y1) * W;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int ox = c8 * 8 + j;
      const float fx = ox * sx;
      int x0 = (int)fx;
      if (x0 > W - 1) x0 = W - 1;
      const int x1 = min(x0 + 1, W - 1);
      const float wx = fx - (float)x0;
      const float a = ldf<T>(r0 + x0), b = ldf<T>(r0 + x1), c = ldf<T>(r1 + x0), d = ldf<T>(r1 + x1);
      const float top = fmaf(wx, b - a, a), bot = fmaf(wx, d - c, c);
      v[j] = fmaf(wy, bot - top, top);
    }
    T* o = out + (plane * Ho + oy) * (int64_t)Wo + c8 * 8;
    if (c8 * 8 + 8 <= Wo && (reinterpret_cast<uintptr_t>(o) & 15) == 0) st8<T>(o, v);
    else for (int j = 0; j < 8 && c8 * 8 + j < Wo; ++j) stf<T>(o + j, v[j]);
  }
}

// The same map as two separable passes per (plane, strip of R output rows): the source rows the strip touches are interpolated
// HORIZONTALLY once into shared memory at the output width (fp32: exactly the `top` / `bot` terms above), then every output
// is one vertical lerp of two shared-memory values (128-bit reads, 8 outputs per thread and store).  Identical arithmetic per
// output (bit-identical results), ~5x fewer instructions: ncu showed the one-pass kernel bound by instruction issue
// (70 % issue-active, 64 % SM busy, 0.7 % DRAM on the 128 -> 512 logits map: coordinate math + four 2-byte loads per output).
template <typename T>
__global__ void __launch_bounds__(256)
upsample_rows_kernel(const T* __restrict__ x, T* __restrict__ out, int H, int W, int Ho, int Wo, float sy, float sx, int R, int ns_max) {
  extern __shared__ float hrow[];                                    // [ns_max][Wo]
  const int strips = (Ho + R - 1) / R;
  const int plane = blockIdx.x / strips, strip = blockIdx.x - plane * strips;
  const int oy0 = strip * R, oy1 = min(Ho, oy0 + R);
  const int y_lo = min((int)(oy0 * sy), H - 1);
  const int y_hi = min(min((int)((oy1 - 1) * sy), H - 1) + 1, H - 1);
  const int ns = min(y_hi - y_lo + 1, ns_max);
  const T* xp = x + (int64_t)plane * H * W;
  for (int ox = threadIdx.x; ox < Wo; ox += blockDim.x) {
    const float fx = ox * sx;
    int x0 = (int)fx;
    if (x0 > W - 1) x0 = W - 1;
    const int x1 = min(x0 + 1, W - 1);
    const float wx = fx - (float)x0;
    for (int r = 0; r < ns; ++r) {
      const T* row = xp + (int64_t)(y_lo + r) * W;
      const float a = ldf<T>(row + x0), b = ldf<T>(row + x1);
      hrow[r * Wo + ox] = fmaf(wx, b - a, a);
    }
  }
  __syncthreads();
  const int cols8 = Wo >> 3;
  const int tasks = (oy1 - oy0) * cols8;
  for (int t = threadIdx.x; t < tasks; t += blockDim.x) {
    const int r = t / cols8, c8 = t - r * cols8, oy = oy0 + r;
    const float fy = oy * sy;
    int y0 = (int)fy;
    if (y0 > H - 1) y0 = H - 1;
    const int y1 = min(y0 + 1, H - 1);
    const float wy = fy - (float)y0;
    const float4* tp = reinterpret_cast<const float4*>(hrow + (y0 - y_lo) * Wo + c8 * 8);
    const float4* bp = reinterpret_cast<const float4*>(hrow + (y1 - y_lo) * Wo + c8 * 8);
    const float4 t0 = tp[0], t1 = tp[1], b0 = bp[0], b1 = bp[1];
    float v[8];
    v[0] = fmaf(wy, b0.x - t0.x, t0.x); v[1] = fmaf(wy, b0.y - t0.y, t0.y); v[2] = fmaf(wy, b0.z - t0.z, t0.z); v[3] = fmaf(wy, b0.w - t0.w, t0.w);
    v[4] = fmaf(wy, b1.x - t1.x, t1.x); v[5] = fmaf(wy, b1.y - t1.y, t1.y); v[6] = fmaf(wy, b1.z - t1.z, t1.z); v[7] = fmaf(wy, b1.w - t1.w, t1.w);
    st8<T>(out + ((int64_t)plane * Ho + oy) * Wo + c8 * 8, v);
  }
}

// Serving tail of the segmentation half (deeplab.py:149-167: interpolate the logits to the frame size, arg-max over the classes):
// out[b][oy][ox] = argmax_c round_T( bilinear(x[b][c])[oy][ox] ), lowest class index on ties — the class map a frame consumer
// reads, without the [B][classes][Ho][Wo] logits ever reaching memory (38 MB written and read again per 8-frame batch, and the
// two launches were the serial tail of the forward: 46 + 40 us).  Same two passes as upsample_rows_kernel with all classes of
// the strip resident; the value compared is rounded to the logits' dtype so that the map is bit-identical to the unfused path.
template <typename T>
__global__ void __launch_bounds__(256)
upsample_argmax_kernel(const T* __restrict__ x, uint8_t* __restrict__ out, int C, int H, int W, int Ho, int Wo, float sy, float sx, int R,
                       int ns_max) {
  extern __shared__ float hrow[];                                    // [C][ns_max][Wo]
  const int strips = (Ho + R - 1) / R;
  const int b = blockIdx.x / strips, strip = blockIdx.x - b * strips;
  const int oy0 = strip * R, oy1 = min(Ho, oy0 + R);
  const int y_lo = min((int)(oy0 * sy), H - 1);
  const int y_hi = min(min((int)((oy1 - 1) * sy), H - 1) + 1, H - 1);
  const int ns = min(y_hi - y_lo + 1, ns_max);
  const T* xb = x + (int64_t)b * C * H * W;
  for (int i = threadIdx.x; i < C * Wo; i += blockDim.x) {
    const int c = i / Wo, ox = i - c * Wo;
    const float fx = ox * sx;
    int x0 = (int)fx;
    if (x0 > W - 1) x0 = W - 1;
    const int x1 = min(x0 + 1, W - 1);
    const float wx = fx - (float)x0;
    const T* xp = xb + (int64_t)c * H * W;
    float* hp = hrow + (int64_t)c * ns_max * Wo + ox;
    for (int r = 0; r < ns; ++r) {
      const T* row = xp + (int64_t)(y_lo + r) * W;
      const float a = ldf<T>(row + x0), bb = ldf<T>(row + x1);
      hp[r * Wo] = fmaf(wx, bb - a, a);
    }
  }
  __syncthreads();
  const int cols8 = Wo >> 3;
  const int tasks = (oy1 - oy0) * cols8;
  for (int t = threadIdx.x; t < tasks; t += blockDim.x) {
    const int r = t / cols8, c8 = t - r * cols8, oy = oy0 + r;
    const float fy = oy * sy;
    int y0 = (int)fy;
    if (y0 > H - 1) y0 = H - 1;
    const int y1 = min(y0 + 1, H - 1);
    const float wy = fy - (float)y0;
    float best[8];
    uint32_t arg[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { best[j] = -__int_as_float(0x7f800000); arg[j] = 0u; }
    for (int c = 0; c < C; ++c) {
      const float* base = hrow + (int64_t)c * ns_max * Wo + c8 * 8;
      const float4* tp = reinterpret_cast<const float4*>(base + (y0 - y_lo) * Wo);
      const float4* bp = reinterpret_cast<const float4*>(base + (y1 - y_lo) * Wo);
      const float4 t0 = tp[0], t1 = tp[1], b0 = bp[0], b1 = bp[1];
      float v[8];
      v[0] = fmaf(wy, b0.x - t0.x, t0.x); v[1] = fmaf(wy, b0.y - t0.y, t0.y); v[2] = fmaf(wy, b0.z - t0.z, t0.z); v[3] = fmaf(wy, b0.w - t0.w, t0.w);
      v[4] = fmaf(wy, b1.x - t1.x, t1.x); v[5] = fmaf(wy, b1.y - t1.y, t1.y); v[6] = fmaf(wy, b1.z - t1.z, t1.z); v[7] = fmaf(wy, b1.w - t1.w, t1.w);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float q = v[j];
        if (sizeof(T) == 2) q = __bfloat162float(__float2bfloat16_rn(q));
        if (q > best[j]) { best[j] = q; arg[j] = (uint32_t)c; }
      }
    }
    uint2 w;
    w.x = arg[0] | (arg[1] << 8) | (arg[2] << 16) | (arg[3] << 24);
    w.y = arg[4] | (arg[5] << 8) | (arg[6] << 16) | (arg[7] << 24);
    *reinterpret_cast<uint2*>(out + ((int64_t)b * Ho + oy) * Wo + c8 * 8) = w;
  }
}

// ---- explicit im2col, tap-major: col[b][tap*C + c][q] = x[b][c][oy*s - p + ky*d][ox*s - p + kx*d] (0 outside) ---------------
// Used for the k x k convolutions with many input channels on small maps (fusion radar projections, point reducers, ASPP):
// the gathered matrix is written once (9x the input, a few MB at 16x16 / 32x32) and the GEMM then runs on the TMA-fed
// tcgen05 kernel with no loader instructions, instead of every N-tile CTA re-gathering the same im2col rows.
template <typename T>
__global__ void __launch_bounds__(256)
im2col_kernel(const T* __restrict__ x, T* __restrict__ col, int B, int C, int H, int W, int Ho, int Wo, int kh, int kw, int stride,
              int pad, int dil) {
  const int cols8 = (Wo + 7) >> 3;
  const int taps = kh * kw;
  const int64_t total = (int64_t)B * taps * C * Ho * cols8;
  const int64_t Po = (int64_t)Ho * Wo;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(t % cols8);
    int64_t r = t / cols8;
    const int oy = (int)(r % Ho); r /= Ho;
    const int c = (int)(r % C); r /= C;
    const int tap = (int)(r % taps);
    const int b = (int)(r / taps);
    const int ky = tap / kw, kx = tap - ky * kw;
    const int iy = oy * stride - pad + ky * dil;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
    if (iy >= 0 && iy < H) {
      const T* row = x + (((int64_t)b * C + c) * H + iy) * W;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int ox = c8 * 8 + j;
        const int ix = ox * stride - pad + kx * dil;
        if (ox < Wo && ix >= 0 && ix < W) v[j] = ldf<T>(row + ix);
      }
    }
    T* o = col + (((int64_t)b * taps + tap) * C + c) * Po + (int64_t)oy * Wo + c8 * 8;
    if (c8 * 8 + 8 <= Wo && (reinterpret_cast<uintptr_t>(o) & 15) == 0) st8<T>(o, v);
    else for (int j = 0; j < 8 && c8 * 8 + j < Wo; ++j) stf<T>(o + j, v[j]);
  }
}

// 3x3 / pad 1 / dilation 1 / stride 1 or 2 on bf16 maps (every im2col launch of the backbone): one thread = 8 consecutive
// output columns of one (plane, output row) for ALL nine taps.  The three input rows are fetched once with 128-bit loads
// (+ one or two halo elements), the column shifts are funnel shifts / byte permutes on the packed bf16 words, and every
// tap leaves as one 16-byte store: 9 or 12 loads and 9 stores per thread instead of 72 scalar loads, and one index
// decode per nine outputs (the generic kernel above reached 1.1-1.5 TB/s, profiles/r01_kernel_table.txt).
template <int STRIDE>
__global__ void __launch_bounds__(256)
im2col3_bf16_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ col, uint32_t total, int C, int H, int W, int Ho,
                    int Wo) {
  const uint32_t cols8 = (uint32_t)Wo >> 3;
  const int64_t Po = (int64_t)Ho * Wo;
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const uint32_t c8 = t % cols8, r = t / cols8;
    const int oy = (int)(r % (uint32_t)Ho);
    const uint32_t plane = r / (uint32_t)Ho;                          // b * C + c
    const uint32_t b = plane / (uint32_t)C, c = plane - b * (uint32_t)C;
    const int x0 = (int)c8 * 8 * STRIDE;                               // first input column of the centre tap
    const __nv_bfloat16* xp = x + (int64_t)plane * H * W;
    __nv_bfloat16* op = col + ((int64_t)b * 9 * C + c) * Po + (int64_t)oy * Wo + c8 * 8;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = oy * STRIDE - 1 + ky;
      uint4 o0 = make_uint4(0u, 0u, 0u, 0u), o1 = o0, o2 = o0;
      if (iy >= 0 && iy < H) {
        const __nv_bfloat16* row = xp + (int64_t)iy * W;
        const uint32_t left = x0 > 0 ? (uint32_t)__bfloat16_as_ushort(row[x0 - 1]) : 0u;
        if (STRIDE == 1) {
          const uint4 v = __ldg(reinterpret_cast<const uint4*>(row + x0));
          const uint32_t right = x0 + 8 < W ? (uint32_t)__bfloat16_as_ushort(row[x0 + 8]) : 0u;
          o1 = v;
          o0 = make_uint4((v.x << 16) | left, __funnelshift_l(v.x, v.y, 16), __funnelshift_l(v.y, v.z, 16), __funnelshift_l(v.z, v.w, 16));
          o2 = make_uint4(__funnelshift_r(v.x, v.y, 16), __funnelshift_r(v.y, v.z, 16), __funnelshift_r(v.z, v.w, 16), (v.w >> 16) | (right << 16));
        } else {
          const uint4 v = __ldg(reinterpret_cast<const uint4*>(row + x0)), u = __ldg(reinterpret_cast<const uint4*>(row + x0 + 8));
          // centre tap: even elements; right tap: odd elements; left tap: odd elements one word earlier
          o1 = make_uint4(__byte_perm(v.x, v.y, 0x5410), __byte_perm(v.z, v.w, 0x5410), __byte_perm(u.x, u.y, 0x5410), __byte_perm(u.z, u.w, 0x5410));
          o2 = make_uint4(__byte_perm(v.x, v.y, 0x7632), __byte_perm(v.z, v.w, 0x7632), __byte_perm(u.x, u.y, 0x7632), __byte_perm(u.z, u.w, 0x7632));
          o0 = make_uint4(left | (v.x & 0xffff0000u), __byte_perm(v.y, v.z, 0x7632), __byte_perm(v.w, u.x, 0x7632), __byte_perm(u.y, u.z, 0x7632));
        }
      }
      __nv_bfloat16* ot = op + (int64_t)(3 * ky) * C * Po;
      *reinterpret_cast<uint4*>(ot) = o0;
      *reinterpret_cast<uint4*>(ot + (int64_t)C * Po) = o1;
      *reinterpret_cast<uint4*>(ot + 2 * (int64_t)C * Po) = o2;
    }
  }
}

// Horizontal taps only (row-tap mode of the convolution engine, vrcoc.h k_order 2): cols[b][kx*C + c][y][x] = x[b][c][y][x + (kx-1)*dil].
// The vertical taps are TMA boxes of THIS tensor shifted by rows, so a 3x3 convolution writes 3x its input instead of 9x.
// dil == 1, bf16, W % 8 == 0: one thread = 8 columns, funnel shifts as above; otherwise a scalar kernel (small dilated maps).
__global__ void __launch_bounds__(256)
im2col_rows3_bf16_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ cols, uint32_t total, int C, int H, int W) {
  const uint32_t cols8 = (uint32_t)W >> 3;
  const int64_t P = (int64_t)H * W;
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const uint32_t c8 = t % cols8, r = t / cols8;
    const uint32_t y = r % (uint32_t)H, plane = r / (uint32_t)H;       // plane = b * C + c
    const uint32_t b = plane / (uint32_t)C, c = plane - b * (uint32_t)C;
    const int x0 = (int)c8 * 8;
    const __nv_bfloat16* row = x + (int64_t)plane * P + (int64_t)y * W;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(row + x0));
    const uint32_t left = x0 > 0 ? (uint32_t)__bfloat16_as_ushort(row[x0 - 1]) : 0u;
    const uint32_t right = x0 + 8 < W ? (uint32_t)__bfloat16_as_ushort(row[x0 + 8]) : 0u;
    __nv_bfloat16* op = cols + ((int64_t)b * 3 * C + c) * P + (int64_t)y * W + x0;
    *reinterpret_cast<uint4*>(op) =
        make_uint4((v.x << 16) | left, __funnelshift_l(v.x, v.y, 16), __funnelshift_l(v.y, v.z, 16), __funnelshift_l(v.z, v.w, 16));
    *reinterpret_cast<uint4*>(op + (int64_t)C * P) = v;
    *reinterpret_cast<uint4*>(op + 2 * (int64_t)C * P) =
        make_uint4(__funnelshift_r(v.x, v.y, 16), __funnelshift_r(v.y, v.z, 16), __funnelshift_r(v.z, v.w, 16), (v.w >> 16) | (right << 16));
  }
}
template <typename T>
__global__ void __launch_bounds__(256)
im2col_rows_kernel(const T* __restrict__ x, T* __restrict__ cols, int64_t total, int C, int H, int W, int kw, int dil) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int xx = (int)(t % W);
    const int64_t r = t / W;
    const int y = (int)(r % H);
    const int64_t q = r / H;                                            // (b * kw + kx) * C + c
    const int c = (int)(q % C);
    const int64_t bk = q / C;
    const int kx = (int)(bk % kw);
    const int64_t b = bk / kw;
    const int ix = xx + (kx - kw / 2) * dil;
    cols[t] = (ix >= 0 && ix < W) ? x[((b * C + c) * H + y) * (int64_t)W + ix] : T(0.f);
  }
}

// ---- depthwise k x k convolution (DWConv.dconv of the decoupled head, reference normal_conv.py:26-27) -------------------
// one thread = 8 consecutive output columns of one (plane, row); weights of the plane in registers; bandwidth-bound.
template <typename T, int K>
__global__ void __launch_bounds__(256)
dwconv_kernel(const T* __restrict__ x, const T* __restrict__ w, const float* __restrict__ bias, T* __restrict__ out, int planes, int C,
              int H, int W, int Ho, int Wo, int stride, int pad) {
  const int cols8 = (Wo + 7) >> 3;
  const int64_t total = (int64_t)planes * Ho * cols8;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(t % cols8);
    const int64_t rowid = t / cols8;
    const int oy = (int)(rowid % Ho);
    const int64_t plane = rowid / Ho;
    const int c = (int)(plane % C);
    float wk[K * K];
#pragma unroll
    for (int i = 0; i < K * K; ++i) wk[i] = ldf<T>(w + (int64_t)c * K * K + i);
    const float b0 = bias ? bias[c] : 0.f;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = b0;
#pragma unroll
    for (int ky = 0; ky < K; ++ky) {
      const int iy = oy * stride - pad + ky;
      if (iy < 0 || iy >= H) continue;
      const T* row = x + (plane * H + iy) * W;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int ox = c8 * 8 + j;
#pragma unroll
        for (int kx = 0; kx < K; ++kx) {
          const int ix = ox * stride - pad + kx;
          if (ox < Wo && ix >= 0 && ix < W) v[j] = fmaf(wk[ky * K + kx], ldf<T>(row + ix), v[j]);
        }
      }
    }
    T* o = out + (plane * Ho + oy) * (int64_t)Wo + c8 * 8;
    if (c8 * 8 + 8 <= Wo && (reinterpret_cast<uintptr_t>(o) & 15) == 0) st8<T>(o, v);
    else for (int j = 0; j < 8 && c8 * 8 + j < Wo; ++j) stf<T>(o + j, v[j]);
  }
}

// 3x3 / stride 1 / pad 1 fast path (every depthwise conv of the decoupled head): one thread = 8 columns x R output rows.
// Each input row is fetched once per thread as one 16-byte vector plus its two halo elements and feeds up to three output
// rows; R + 2 row fetches per R output rows instead of 3R, no per-tap predicates.  Requires W % 8 == 0 and 16-byte aligned
// planes (host-checked).
template <typename T, int R>
__global__ void __launch_bounds__(256)
dwconv3_s1_kernel(const T* __restrict__ x, const T* __restrict__ w, const float* __restrict__ bias, T* __restrict__ out, int planes, int C,
                  int H, int W) {
  const int cols8 = W >> 3;
  const int rgroups = (H + R - 1) / R;
  const int64_t total = (int64_t)planes * rgroups * cols8;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(t % cols8);
    const int64_t rest = t / cols8;
    const int rg = (int)(rest % rgroups);
    const int64_t plane = rest / rgroups;
    const int c = (int)(plane % C);
    float wk[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) wk[i] = ldf<T>(w + (int64_t)c * 9 + i);
    const float b0 = bias ? bias[c] : 0.f;
    float acc[R][8];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[r][j] = b0;
    const int oy0 = rg * R, x0 = c8 * 8;
    const T* px = x + plane * (int64_t)H * W;
#pragma unroll
    for (int r = -1; r <= R; ++r) {
      const int iy = oy0 + r;
      if (iy < 0 || iy >= H) continue;
      const T* row = px + (int64_t)iy * W;
      float v[10], mid[8];
      ld8<T>(row + x0, mid);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j + 1] = mid[j];
      v[0] = x0 > 0 ? ldf<T>(row + x0 - 1) : 0.f;
      v[9] = x0 + 8 < W ? ldf<T>(row + x0 + 8) : 0.f;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int rel = r - ky + 1;                 // output row (relative) this input row feeds through tap row ky
        if (rel < 0 || rel >= R) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          acc[rel][j] = fmaf(wk[ky * 3 + 2], v[j + 2], fmaf(wk[ky * 3 + 1], v[j + 1], fmaf(wk[ky * 3], v[j], acc[rel][j])));
      }
    }
    T* po = out + plane * (int64_t)H * W + x0;
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (oy0 + r < H) st8<T>(po + (int64_t)(oy0 + r) * W, acc[r]);
  }
}

template <typename F>
static int by_dtype(int dt, F&& f) {
  if (dt == VRCOC_F32) return f((float*)nullptr);
  if (dt == VRCOC_BF16) return f((__nv_bfloat16*)nullptr);
  return fail(VRCOC_EINVAL, "unknown dtype %d", dt);
}

}  // namespace vrcoc

using namespace vrcoc;

extern "C" const char* vrcoc_version(void) { return "vrcoc-b200 0.1 (sm_100a)"; }
extern "C" const char* vrcoc_last_error(void) { return err_buf(); }
extern "C" int vrcoc_device_ok(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}

extern "C" int vrcoc_channel_sums(const void* x, int dtype, int B, int C, int HW, float* chan_sums, double* sample_sums,
                                  void* stream) {
  VRCOC_REQUIRE(x && (chan_sums || sample_sums), "channel_sums: null pointer");
  VRCOC_REQUIRE(B > 0 && C > 0 && HW > 0, "channel_sums: non-positive dimension");
  cudaStream_t st = (cudaStream_t)stream;
  return by_dtype(dtype, [&](auto* t) {
    using T = typename std::remove_pointer<decltype(t)>::type;
    const int chunks = (int)cdiv(HW, PLANE_CHUNK);
    if (chunks > 1 && chan_sums) cudaMemsetAsync(chan_sums, 0, sizeof(float) * 2 * B * C, st);
    channel_sums_kernel<T><<<dim3(chunks, B * C), 256, 0, st>>>((const T*)x, HW, chan_sums, sample_sums, C);
    return check_launch("channel_sums");
  });
}

extern "C" int vrcoc_chan_affine(const void* x, int x_dtype, const void* res, int res_dtype, void* out, int out_dtype,
                                 const float* s1, const float* t1, int act, const float* s2, const float* t2, int B, int C,
                                 int HW, float* out_chan_sums, uint32_t* out_minmax, void* stream) {
  VRCOC_REQUIRE(x && out, "chan_affine: null pointer");
  VRCOC_REQUIRE(B > 0 && C > 0 && HW > 0, "chan_affine: non-positive dimension");
  VRCOC_REQUIRE(x_dtype == out_dtype && (!res || res_dtype == x_dtype), "chan_affine: mixed dtypes unsupported");
  cudaStream_t st = (cudaStream_t)stream;
  return by_dtype(x_dtype, [&](auto* t) {
    using T = typename std::remove_pointer<decltype(t)>::type;
    const int chunks = (int)cdiv(HW, PLANE_CHUNK);
    if (chunks > 1 && out_chan_sums) cudaMemsetAsync(out_chan_sums, 0, sizeof(float) * 2 * B * C, st);
    chan_affine_kernel<T><<<dim3(chunks, B * C), 256, 0, st>>>((const T*)x, (const T*)res, (T*)out, s1, t1, act, s2, t2, C, HW,
                                                               out_chan_sums, out_minmax);
    return check_launch("chan_affine");
  });
}

extern "C" int vrcoc_img_enh_finish(const void* k, int k_dtype, const void* image, int image_dtype, void* out, int out_dtype,
                                    const uint32_t* minmax, const float* s, const float* t, int B, int C, int HW,
                                    float* out_chan_sums, void* stream) {
  VRCOC_REQUIRE(k && image && out && minmax, "img_enh_finish: null pointer");
  VRCOC_REQUIRE(B > 0 && C > 0 && HW > 0, "img_enh_finish: non-positive dimension");
  VRCOC_REQUIRE(k_dtype == image_dtype && k_dtype == out_dtype, "img_enh_finish: mixed dtypes unsupported");
  cudaStream_t st = (cudaStream_t)stream;
  return by_dtype(k_dtype, [&](auto* tt) {
    using T = typename std::remove_pointer<decltype(tt)>::type;
    const int chunks = (int)cdiv(HW, PLANE_CHUNK);
    if (chunks > 1 && out_chan_sums) cudaMemsetAsync(out_chan_sums, 0, sizeof(float) * 2 * B * C, st);
    img_enh_finish_kernel<T><<<dim3(chunks, B * C), 256, 0, st>>>((const T*)k, (const T*)image, (T*)out, minmax, s, t, C, HW,
                                                                  out_chan_sums);
    return check_launch("img_enh_finish");
  });
}

static int launch_sa_gate_sums(const void* image, const void* radar, int dtype, int B, int Ci, int Cr, int HW, int G,
                               const float* chan_sums_img, const float* cweight, const float* cbias, const float* sweight,
                               const float* sbias, const float* gn_weight, const float* gn_bias, float* attn, float* cs_radar,
                               void* stream, const char* what) {
  VRCOC_REQUIRE(image && attn, "%s: null pointer", what);
  VRCOC_REQUIRE(B > 0 && Ci > 0 && HW > 0 && Cr >= 0, "%s: bad dimension", what);
  VRCOC_REQUIRE(Cr == 0 || (radar && cs_radar), "%s: null radar pointer", what);
  if (G > 0) {
    VRCOC_REQUIRE(Ci % (2 * G) == 0 && Ci >= 2 * G, "%s: channels %d not divisible by 2*G=%d", what, Ci, 2 * G);
    VRCOC_REQUIRE(cweight && cbias && sweight && sbias && gn_weight && gn_bias, "%s: null attention parameter", what);
  }
  cudaStream_t st = (cudaStream_t)stream;
  return by_dtype(dtype, [&](auto* t) {
    using T = typename std::remove_pointer<decltype(t)>::type;
    sa_gate_sums_kernel<T><<<B * (Ci + Cr), 256, 0, st>>>((const T*)image, Ci, HW, G, chan_sums_img, cweight, cbias, sweight, sbias,
                                                          gn_weight, gn_bias, attn, (const T*)radar, Cr, cs_radar, B * Ci);
    return check_launch(what);
  });
}

extern "C" int vrcoc_sa_gate_sums(const void* image, int dtype, int B, int Ci, int HW, int G, const float* chan_sums_img,
                                  const float* cweight, const float* cbias, const float* sweight, const float* sbias,
                                  const float* gn_weight, const float* gn_bias, float* attn, void* stream) {
  return launch_sa_gate_sums(image, nullptr, dtype, B, Ci, 0, HW, G, chan_sums_img, cweight, cbias, sweight, sbias, gn_weight, gn_bias,
                             attn, nullptr, stream, "sa_gate_sums");
}

extern "C" int vrcoc_fusion_stats(const void* image, const void* radar, int dtype, int B, int Ci, int Cr, int HW, int G,
                                  const float* cweight, const float* cbias, const float* sweight, const float* sbias,
                                  const float* gn_weight, const float* gn_bias, float* attn, float* chan_sums_radar, void* stream) {
  return launch_sa_gate_sums(image, radar, dtype, B, Ci, Cr, HW, G, nullptr, cweight, cbias, sweight, sbias, gn_weight, gn_bias, attn,
                             chan_sums_radar, stream, "fusion_stats");
}

extern "C" int vrcoc_radar_enh_table(const float* attn, const float* chan_sums_radar, const int32_t* chan_src,
                                     const float* eca_weight, int eca_k, int B, int Ci, int Cr, int HW, float* table,
                                     void* stream) {
  VRCOC_REQUIRE(attn && chan_sums_radar && eca_weight && table, "radar_enh_table: null pointer");
  VRCOC_REQUIRE(B > 0 && Ci > 0 && Cr > 0 && HW > 0 && eca_k > 0 && (eca_k & 1), "radar_enh_table: bad dimension");
  int K = Ci + Cr;
  VRCOC_REQUIRE(K * 4 <= 48 * 1024, "radar_enh_table: too many channels (%d)", K);
  radar_enh_table_kernel<<<B, 256, K * sizeof(float), (cudaStream_t)stream>>>(attn, chan_sums_radar, chan_src, eca_weight,
                                                                              eca_k, Ci, Cr, HW, table, 0);
  return check_launch("radar_enh_table");
}

extern "C" int vrcoc_radar_enh_table_concat_order(const float* attn, const float* chan_sums_radar, const int32_t* chan_src,
                                                  const float* eca_weight, int eca_k, int B, int Ci, int Cr, int HW, float* table,
                                                  void* stream) {
  VRCOC_REQUIRE(attn && chan_sums_radar && eca_weight && table && chan_src, "radar_enh_table_concat_order: null pointer");
  VRCOC_REQUIRE(B > 0 && Ci > 0 && Cr > 0 && HW > 0 && eca_k > 0 && (eca_k & 1), "radar_enh_table_concat_order: bad dimension");
  int K = Ci + Cr;
  VRCOC_REQUIRE(K * 4 <= 48 * 1024, "radar_enh_table_concat_order: too many channels (%d)", K);
  radar_enh_table_kernel<<<B, 256, K * sizeof(float), (cudaStream_t)stream>>>(attn, chan_sums_radar, chan_src, eca_weight,
                                                                              eca_k, Ci, Cr, HW, table, 1);
  return check_launch("radar_enh_table_concat_order");
}

extern "C" int vrcoc_gelu_bwd(const void* dy, const void* u, void* out, int dtype, int64_t n, void* stream) {
  VRCOC_REQUIRE(dy && u && out && n > 0, "gelu_bwd: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  int blocks = (int)(cdiv(n, 256) < 148 * 16 ? cdiv(n, 256) : 148 * 16);
  return by_dtype(dtype, [&](auto* t) {
    using T = typename std::remove_pointer<decltype(t)>::type;
    const int align = 8 * (int)sizeof(T) - 1;
    if ((n & 7) == 0 && ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(u) | reinterpret_cast<uintptr_t>(out)) & align) == 0) {
      const int64_t n8 = n >> 3;
      const int vb = (int)(cdiv(n8, 256) < 148 * 16 ? cdiv(n8, 256) : 148 * 16);
      gelu_bwd_vec_kernel<T><<<vb, 256, 0, st>>>((const T*)dy, (const T*)u, (T*)out, n8);
      return check_launch("gelu_bwd");
    }
    gelu_bwd_kernel<T, T, T><<<blocks, 256, 0, st>>>((const T*)dy, (const T*)u, (T*)out, n);
    return check_launch("gelu_bwd");
  });
}

extern "C" int vrcoc_gn_bwd_sums(const void* dz, const void* x, int dtype, int B, int C, int HW, float* out, void* stream) {
  VRCOC_REQUIRE(dz && x && out && B > 0 && C > 0 && HW > 0, "gn_bwd_sums: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  return by_dtype(dtype, [&](auto* t) {
    using T = typename std::remove_pointer<decltype(t)>::type;
    gn_bwd_sums_kernel<T, T><<<B * C, 256, 0, st>>>((const T*)dz, (const T*)x, HW, out);
    return check_launch("gn_bwd_sums");
  });
}

extern "C" int vrcoc_gn_bwd_apply(const void* dz, const void* x, const void* extra, void* out, int dtype, const float* a,
                                  const float* bb, const float* cc, int B, int C, int HW, void* stream) {
  VRCOC_REQUIRE(dz && x && out && a && bb && cc && B > 0 && C > 0 && HW > 0, "gn_bwd_apply: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  return by_dtype(dtype, [&](auto* t) {
    using T = typename std::remove_pointer<decltype(t)>::type;
    gn_bwd_apply_kernel<T, T, T><<<B * C, 256, 0, st>>>((const T*)dz, (const T*)x, (const T*)extra, (T*)out, a, bb, cc, C, HW);
    return check_launch("gn_bwd_apply");
  });
}

// strip height and resident source rows of the two-pass upsample kernels
static void upsample_plan(int H, int Ho, int Wo, float sy, int& R, int& ns_max, int& threads, int r_pref = 16) {
  R = Ho < r_pref ? Ho : r_pref;
  ns_max = (int)((R - 1) * sy) + 3;
  if (ns_max > H) ns_max = H;
  const int tasks = R * (Wo >> 3);
  threads = tasks >= 256 ? 256 : (tasks <= 64 ? 64 : ((tasks + 31) / 32) * 32);
}

extern "C" int vrcoc_upsample_bilinear(const void* x, void* out, int dtype, int planes, int H, int W, int Ho, int Wo, void* stream) {
  VRCOC_REQUIRE(x && out && planes > 0 && H > 0 && W > 0 && Ho > 0 && Wo > 0, "upsample_bilinear: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const float sy = Ho > 1 ? (float)(H - 1) / (float)(Ho - 1) : 0.f;
  const float sx = Wo > 1 ? (float)(W - 1) / (float)(Wo - 1) : 0.f;
  int R, ns_max, threads;
  upsample_plan(H, Ho, Wo, sy, R, ns_max, threads);
  const size_t smem = (size_t)ns_max * Wo * sizeof(float);
  const int64_t nblk = (int64_t)planes * ((Ho + R - 1) / R);
  const bool aligned = ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  if ((Wo & 7) == 0 && Ho >= H && Wo >= W && smem <= 48 * 1024 && nblk < (1ll << 31) && aligned) {
    return by_dtype(dtype, [&](auto* t) {
      using T = typename std::remove_pointer<decltype(t)>::type;
      upsample_rows_kernel<T><<<(unsigned)nblk, threads, smem, st>>>((const T*)x, (T*)out, H, W, Ho, Wo, sy, sx, R, ns_max);
      return check_launch("upsample_bilinear");
    });
  }
  const int64_t total = (int64_t)planes * Ho * ((Wo + 7) / 8);
  int blocks = (int)(cdiv(total, 256) < 148 * 16 ? cdiv(total, 256) : 148 * 16);
  return by_dtype(dtype, [&](auto* t) {
    using T = typename std::remove_pointer<decltype(t)>::type;
    upsample_bilinear_kernel<T><<<blocks, 256, 0, st>>>((const T*)x, (T*)out, planes, H, W, Ho, Wo, sy, sx);
    return check_launch("upsample_bilinear");
  });
}

extern "C" int vrcoc_upsample_argmax_supported(int C, int H, int W, int Ho, int Wo) {
  if (C <= 0 || C > 255 || H <= 0 || W <= 0 || Ho < H || Wo < W || (Wo & 7) != 0) return 0;
  const float sy = Ho > 1 ? (float)(H - 1) / (float)(Ho - 1) : 0.f;
  for (int r_pref = 16; r_pref >= 4; r_pref >>= 1) {                  // shorter strips when the classes of a 16-row strip do not fit
    int R, ns_max, threads;
    upsample_plan(H, Ho, Wo, sy, R, ns_max, threads, r_pref);
    if ((size_t)C * ns_max * Wo * sizeof(float) <= 200 * 1024) return r_pref;
  }
  return 0;
}

extern "C" int vrcoc_upsample_argmax(const void* x, uint8_t* out, int dtype, int B, int C, int H, int W, int Ho, int Wo, void* stream) {
  VRCOC_REQUIRE(x && out && B > 0, "upsample_argmax: bad argument");
  const int r_pref = vrcoc_upsample_argmax_supported(C, H, W, Ho, Wo);
  VRCOC_REQUIRE(r_pref > 0, "upsample_argmax: unsupported shape (classes <= 255, Wo %% 8 == 0, classes x strip rows x Wo floats within "
                "shared memory)");
  VRCOC_REQUIRE((reinterpret_cast<uintptr_t>(out) & 7) == 0, "upsample_argmax: out must be 8-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const float sy = Ho > 1 ? (float)(H - 1) / (float)(Ho - 1) : 0.f;
  const float sx = Wo > 1 ? (float)(W - 1) / (float)(Wo - 1) : 0.f;
  int R, ns_max, threads;
  upsample_plan(H, Ho, Wo, sy, R, ns_max, threads, r_pref);
  const size_t smem = (size_t)C * ns_max * Wo * sizeof(float);
  const unsigned nblk = (unsigned)(B * ((Ho + R - 1) / R));
  return by_dtype(dtype, [&](auto* t) {
    using T = typename std::remove_pointer<decltype(t)>::type;
    auto kern = upsample_argmax_kernel<T>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<nblk, 256, smem, st>>>((const T*)x, out, C, H, W, Ho, Wo, sy, sx, R, ns_max);
    return check_launch("upsample_argmax");
  });
}

extern "C" int vrcoc_im2col(const void* x, void* col, int dtype, int B, int C, int H, int W, int kh, int kw, int stride, int pad,
                            int dil, void* stream) {
  VRCOC_REQUIRE(x && col && B > 0 && C > 0 && H > 0 && W > 0 && kh > 0 && kw > 0 && stride > 0 && pad >= 0 && dil > 0, "im2col: bad argument");
  const int Ho = (H + 2 * pad - dil * (kh - 1) - 1) / stride + 1, Wo = (W + 2 * pad - dil * (kw - 1) - 1) / stride + 1;
  VRCOC_REQUIRE(Ho > 0 && Wo > 0, "im2col: empty output");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t total = (int64_t)B * kh * kw * C * Ho * ((Wo + 7) / 8);
  int blocks = (int)(cdiv(total, 256) < 148 * 32 ? cdiv(total, 256) : 148 * 32);
  const int64_t fast_total = (int64_t)B * C * Ho * (Wo / 8);
  if (dtype == VRCOC_BF16 && kh == 3 && kw == 3 && pad == 1 && dil == 1 && (stride == 1 || stride == 2) && (Wo % 8) == 0 &&
      (W % (8 * stride)) == 0 && Wo * stride == W && fast_total < (1ll << 31) &&
      ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(col)) & 15) == 0) {
    const int nb = (int)cdiv(fast_total, 256);
    if (stride == 1)
      im2col3_bf16_kernel<1><<<nb, 256, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)col, (uint32_t)fast_total, C, H, W, Ho, Wo);
    else
      im2col3_bf16_kernel<2><<<nb, 256, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)col, (uint32_t)fast_total, C, H, W, Ho, Wo);
    return check_launch("im2col3");
  }
  return by_dtype(dtype, [&](auto* t) {
    using T = typename std::remove_pointer<decltype(t)>::type;
    im2col_kernel<T><<<blocks, 256, 0, st>>>((const T*)x, (T*)col, B, C, H, W, Ho, Wo, kh, kw, stride, pad, dil);
    return check_launch("im2col");
  });
}

extern "C" int vrcoc_im2col_rows(const void* x, void* cols, int dtype, int B, int C, int H, int W, int kw, int dil, void* stream) {
  VRCOC_REQUIRE(x && cols && B > 0 && C > 0 && H > 0 && W > 0 && kw > 0 && (kw & 1) == 1 && dil > 0, "im2col_rows: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t fast_total = (int64_t)B * C * H * (W / 8);
  if (dtype == VRCOC_BF16 && kw == 3 && dil == 1 && (W % 8) == 0 && fast_total < (1ll << 31) &&
      ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(cols)) & 15) == 0) {
    im2col_rows3_bf16_kernel<<<(unsigned)cdiv(fast_total, 256), 256, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)cols,
                                                                           (uint32_t)fast_total, C, H, W);
    return check_launch("im2col_rows3");
  }
  const int64_t total = (int64_t)B * kw * C * H * W;
  const int blocks = (int)(cdiv(total, 256) < 148 * 32 ? cdiv(total, 256) : 148 * 32);
  return by_dtype(dtype, [&](auto* t) {
    using T = typename std::remove_pointer<decltype(t)>::type;
    im2col_rows_kernel<T><<<blocks, 256, 0, st>>>((const T*)x, (T*)cols, total, C, H, W, kw, dil);
    return check_launch("im2col_rows");
  });
}

extern "C" int vrcoc_dwconv(const void* x, const void* weight, const float* bias, void* out, int dtype, int B, int C, int H, int W,
                            int k, int stride, int pad, void* stream) {
  VRCOC_REQUIRE(x && weight && out && B > 0 && C > 0 && H > 0 && W > 0 && stride > 0 && pad >= 0, "dwconv: bad argument");
  VRCOC_REQUIRE(k == 3 || k == 5, "dwconv: kernel size %d unsupported (3 or 5)", k);
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t total = (int64_t)B * C * Ho * ((Wo + 7) / 8);
  int blocks = (int)(cdiv(total, 256) < 148 * 32 ? cdiv(total, 256) : 148 * 32);
  return by_dtype(dtype, [&](auto* t) {
    using T = typename std::remove_pointer<decltype(t)>::type;
    const int es = (int)sizeof(T);
    if (k == 3 && stride == 1 && pad == 1 && W % 8 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0 &&
        ((int64_t)H * W * es) % 16 == 0) {
      constexpr int R = 4;
      const int64_t tot = (int64_t)B * C * ((H + R - 1) / R) * (W / 8);
      const int nb = (int)(cdiv(tot, 256) < 148 * 32 ? cdiv(tot, 256) : 148 * 32);
      dwconv3_s1_kernel<T, R><<<nb, 256, 0, st>>>((const T*)x, (const T*)weight, bias, (T*)out, B * C, C, H, W);
      return check_launch("dwconv3");
    }
    if (k == 3) dwconv_kernel<T, 3><<<blocks, 256, 0, st>>>((const T*)x, (const T*)weight, bias, (T*)out, B * C, C, H, W, Ho, Wo, stride, pad);
    else dwconv_kernel<T, 5><<<blocks, 256, 0, st>>>((const T*)x, (const T*)weight, bias, (T*)out, B * C, C, H, W, Ho, Wo, stride, pad);
    return check_launch("dwconv");
  });
}
