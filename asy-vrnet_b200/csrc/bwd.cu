// Small backward kernels of the training path: everything that used to be O(B*C)- or O(O*K)-sized "coefficient algebra" done
// with dozens of tiny library launches per autograd node (torch profiler, tools/prof_train.py: 11 K launches per training step,
// host-bound) plus the adjoint passes of the k x k convolutions and the per-channel normalisations.
//
//   vrcoc_gn_bwd_coef       GroupNorm(1,C) backward coefficients from the per-(b,c) sums        (vr_coc.py:105-111, :265,270)
//   vrcoc_proj_res_bwd_coef layer-scale / bias / weight gradients of  out = res + ls * (W h + b) (vr_coc.py:191,222,266-271)
//   vrcoc_col2im            adjoint of vrcoc_im2col (tap-major): input gradient of a k x k convolution (vr_coc.py:99-102,313)
//   vrcoc_chan_bwd_sums / vrcoc_chan_bwd_apply
//                           per-channel affine (+ activation) backward = BatchNorm2d backward in train and eval mode
//                           (normal_conv.py:45-49, vr_coc.py:315,341,356)
#include "common.cuh"

namespace vrcoc {
namespace {

__device__ __forceinline__ void block_sum2d(double& a, double& b, double* red) {
  a = warp_sum(a); b = warp_sum(b);
  const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) { red[w] = a; red[32 + w] = b; }
  __syncthreads();
  double x = 0.0, y = 0.0;
  for (int i = 0; i < nw; ++i) { x += red[i]; y += red[32 + i]; }
  a = x; b = y;
}

// one CTA per sample: m1, m2 -> a[b][c], bb[b], cc[b]; sxh[b][c] kept for the channel reduction
__global__ void __launch_bounds__(256) gn_bwd_coef_sample_kernel(const float* __restrict__ s, const double* __restrict__ gn_sums,
                                                                 const float* __restrict__ gamma, float eps, int C, int HW,
                                                                 float* __restrict__ a, float* __restrict__ bb, float* __restrict__ cc,
                                                                 float* __restrict__ sxh) {
  __shared__ double red[64];
  const int b = blockIdx.x;
  const double cnt = (double)C * (double)HW;
  double t0 = 0.0, t1 = 0.0;
  if (threadIdx.x < VRCOC_STAT_SLOTS) {
    t0 = gn_sums[((int64_t)b * VRCOC_STAT_SLOTS + threadIdx.x) * 2];
    t1 = gn_sums[((int64_t)b * VRCOC_STAT_SLOTS + threadIdx.x) * 2 + 1];
  }
  block_sum2d(t0, t1, red);
  const double mean = t0 / cnt;
  double var = t1 / cnt - mean * mean;
  var = var > 0.0 ? var : 0.0;
  const double rstd = 1.0 / sqrt(var + (double)eps);
  double m1 = 0.0, m2 = 0.0;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const double s1 = s[((int64_t)b * C + c) * 2], s2 = s[((int64_t)b * C + c) * 2 + 1];
    const double xh = (s2 - mean * s1) * rstd;                    // sum dz * xhat
    sxh[(int64_t)b * C + c] = (float)xh;
    const double g = gamma[c];
    m1 += s1 * g; m2 += xh * g;
    a[(int64_t)b * C + c] = (float)(rstd * g);
  }
  block_sum2d(m1, m2, red);
  m1 /= cnt; m2 /= cnt;
  if (threadIdx.x == 0) {
    bb[b] = (float)(-(rstd * rstd * m2));
    cc[b] = (float)(rstd * (mean * rstd * m2 - m1));
  }
}

// dgamma[c] = sum_b sxh[b][c], dbeta[c] = sum_b s1[b][c]   (fixed order: deterministic)
__global__ void gn_bwd_coef_chan_kernel(const float* __restrict__ s, const float* __restrict__ sxh, int B, int C,
                                        float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double g = 0.0, h = 0.0;
  for (int b = 0; b < B; ++b) { g += sxh[(int64_t)b * C + c]; h += s[((int64_t)b * C + c) * 2]; }
  dgamma[c] = (float)g; dbeta[c] = (float)h;
}

// one CTA per output row o of the projection
template <typename TW>
__global__ void __launch_bounds__(128) proj_res_bwd_coef_kernel(const float* __restrict__ G, const float* __restrict__ sdy, const TW* __restrict__ W,
                                                                const float* __restrict__ bias, const float* __restrict__ ls, int O, int K,
                                                                float* __restrict__ dW, float* __restrict__ db, float* __restrict__ dls,
                                                                TW* __restrict__ wt) {
  __shared__ float red[4];
  const int o = blockIdx.x;
  const float l = ls ? ls[o] : 1.f;
  float acc = 0.f;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const float g = G[(int64_t)o * K + k], w = ldf<TW>(W + (int64_t)o * K + k);
    dW[(int64_t)o * K + k] = l * g;
    acc = fmaf(w, g, acc);
    stf<TW>(wt + (int64_t)k * O + o, w * l);                        // (W * ls)^T: the dgrad weight
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    const float t = red[0] + red[1] + red[2] + red[3];
    const float sy = sdy[o];
    if (db) db[o] = l * sy;
    if (dls) dls[o] = t + (bias ? bias[o] * sy : 0.f);
  }
}

// dx[b][c][y][x] = sum over taps (ky,kx) and output positions (oy,ox) with oy*s - p + ky*d == y, ox*s - p + kx*d == x of
// dcol[b][(ky*kw + kx)*C + c][oy][ox]   (gather form: one thread per input element, no atomics)
template <typename T>
__global__ void __launch_bounds__(256) col2im_kernel(const T* __restrict__ dcol, T* __restrict__ dx, int C, int H, int W, int kh, int kw,
                                                     int stride, int pad, int dil, int Ho, int Wo, int64_t total) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int x = (int)(i % W), y = (int)((i / W) % H), c = (int)((i / ((int64_t)W * H)) % C);
  const int64_t b = i / ((int64_t)W * H * C);
  const int64_t plane = (int64_t)Ho * Wo;
  const T* base = dcol + b * (int64_t)kh * kw * C * plane;
  float acc = 0.f;
  for (int ky = 0; ky < kh; ++ky) {
    const int ty = y + pad - ky * dil;
    if (ty < 0 || ty % stride) continue;
    const int oy = ty / stride;
    if (oy >= Ho) continue;
    for (int kx = 0; kx < kw; ++kx) {
      const int tx = x + pad - kx * dil;
      if (tx < 0 || tx % stride) continue;
      const int ox = tx / stride;
      if (ox >= Wo) continue;
      acc += ldf<T>(base + ((int64_t)(ky * kw + kx) * C + c) * plane + (int64_t)oy * Wo + ox);
    }
  }
  stf<T>(dx + i, acc);
}

// 3x3 / stride 1 / pad 1 / bf16 (ImageEnhanceByRadar's convolution and the other full-resolution 3x3 ones: their dcol is nine times
// the map, 302 MB at stage 1 with batch 16, and the scalar gather read it in 2-byte words: 0.31 ms per launch in the training
// profile): one thread = 8 consecutive input columns of one row; per tap one aligned 16-byte load of the output row it reads plus
// the one halo element the +-1 column shift needs.
__global__ void __launch_bounds__(256) col2im3_s1_bf16_kernel(const __nv_bfloat16* __restrict__ dcol, __nv_bfloat16* __restrict__ dx, int C, int H, int W,
                                                              int64_t total8) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total8) return;
  const int w8 = W >> 3;
  const int x0 = (int)(t % w8) * 8, y = (int)((t / w8) % H), c = (int)((t / ((int64_t)w8 * H)) % C);
  const int64_t b = t / ((int64_t)w8 * H * C);
  const int64_t plane = (int64_t)H * W;
  const __nv_bfloat16* base = dcol + b * 9 * (int64_t)C * plane;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int oy = y + 1 - ky;
    if (oy < 0 || oy >= H) continue;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const __nv_bfloat16* row = base + ((int64_t)(ky * 3 + kx) * C + c) * plane + (int64_t)oy * W;
      float v[8];
      ld8<__nv_bfloat16>(row + x0, v);
      if (kx == 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += v[j];
      } else if (kx == 0) {                                           // output column x + 1
        const float edge = x0 + 8 < W ? __bfloat162float(row[x0 + 8]) : 0.f;
#pragma unroll
        for (int j = 0; j < 7; ++j) acc[j] += v[j + 1];
        acc[7] += edge;
      } else {                                                        // output column x - 1
        const float edge = x0 > 0 ? __bfloat162float(row[x0 - 1]) : 0.f;
        acc[0] += edge;
#pragma unroll
        for (int j = 1; j < 8; ++j) acc[j] += v[j - 1];
      }
    }
  }
  st8<__nv_bfloat16>(dx + ((b * C + c) * (int64_t)H + y) * W + x0, acc);
}

// 3x3 / stride 2 / pad 1 / bf16 (the point reducers between the stages, vr_coc.py:83-102): input column x takes tap kx = 1 from output
// column x/2 when x is even, taps kx = 0 and kx = 2 from output columns (x+1)/2 and (x-1)/2 when it is odd, and the same for the rows;
// one thread = 8 consecutive input columns of one row = four output columns per tap plane (one aligned 8-byte load each, plus one
// halo element for kx = 0), instead of a 9-tap scalar gather with a division and a modulo per tap and element.
__global__ void __launch_bounds__(256) col2im3_s2_bf16_kernel(const __nv_bfloat16* __restrict__ dcol, __nv_bfloat16* __restrict__ dx, int C, int H, int W,
                                                              int Ho, int Wo, int64_t total8) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total8) return;
  const int w8 = W >> 3;
  const int x0 = (int)(t % w8) * 8, y = (int)((t / w8) % H), c = (int)((t / ((int64_t)w8 * H)) % C);
  const int64_t b = t / ((int64_t)w8 * H * C);
  const int64_t plane = (int64_t)Ho * Wo;
  const __nv_bfloat16* base = dcol + b * 9 * (int64_t)C * plane;
  const int ox0 = x0 >> 1;                                            // multiple of 4: the 8-byte loads below are aligned
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  auto load4 = [&](const __nv_bfloat16* p, float (&v)[4]) {
    const uint2 raw = __ldg(reinterpret_cast<const uint2*>(p));
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.x));
    const float2 bb = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.y));
    v[0] = a.x; v[1] = a.y; v[2] = bb.x; v[3] = bb.y;
  };
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int ty = y + 1 - ky;
    if (ty < 0 || (ty & 1)) continue;
    const int oy = ty >> 1;
    if (oy >= Ho) continue;
    const __nv_bfloat16* r0 = base + ((int64_t)(ky * 3 + 0) * C + c) * plane + (int64_t)oy * Wo;
    const __nv_bfloat16* r1 = r0 + (int64_t)C * plane;
    const __nv_bfloat16* r2 = r1 + (int64_t)C * plane;
    float v[4];
    load4(r1 + ox0, v);                                               // kx = 1: even columns x0 + 2i <- output column ox0 + i
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[2 * i] += v[i];
    load4(r2 + ox0, v);                                               // kx = 2: odd columns x0 + 2i + 1 <- output column ox0 + i
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[2 * i + 1] += v[i];
    load4(r0 + ox0, v);                                               // kx = 0: odd columns x0 + 2i + 1 <- output column ox0 + i + 1
    const float edge = ox0 + 4 < Wo ? __bfloat162float(r0[ox0 + 4]) : 0.f;
    acc[1] += v[1]; acc[3] += v[2]; acc[5] += v[3]; acc[7] += edge;
  }
  st8<__nv_bfloat16>(dx + ((b * C + c) * (int64_t)H + y) * W + x0, acc);
}

// k x k / stride k / pad 0 / bf16, k = 4 (the patch embedding, vr_coc.py:582-586): a permutation — input pixel (y, x) is tap (y % 4, x % 4) of
// output pixel (y / 4, x / 4); one thread = 8 consecutive input columns = two output columns of each of the four kx planes of its row's ky.
__global__ void __launch_bounds__(256) col2im4_s4_bf16_kernel(const __nv_bfloat16* __restrict__ dcol, __nv_bfloat16* __restrict__ dx, int C, int H, int W,
                                                              int Ho, int Wo, int64_t total8) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total8) return;
  const int w8 = W >> 3;
  const int x0 = (int)(t % w8) * 8, y = (int)((t / w8) % H), c = (int)((t / ((int64_t)w8 * H)) % C);
  const int64_t b = t / ((int64_t)w8 * H * C);
  const int64_t plane = (int64_t)Ho * Wo;
  const int ky = y & 3, oy = y >> 2, ox0 = x0 >> 2;                   // ox0 even: the 4-byte loads are aligned
  const __nv_bfloat16* base = dcol + (b * 16 * (int64_t)C + (int64_t)(ky * 4) * C + c) * plane + (int64_t)oy * Wo + ox0;
  float acc[8];
#pragma unroll
  for (int kx = 0; kx < 4; ++kx) {
    const uint32_t raw = __ldg(reinterpret_cast<const uint32_t*>(base + (int64_t)kx * C * plane));
    const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw));
    acc[kx] = v.x;
    acc[4 + kx] = v.y;
  }
  st8<__nv_bfloat16>(dx + ((b * C + c) * (int64_t)H + y) * W + x0, acc);
}

// act'(.) evaluated from the forward OUTPUT y of the activation (relu / lrelu: sign of y; none: 1); SiLU needs the
// pre-activation z = u * zs[c] + zt[c] (the normalised convolution output), recomputed from u
__device__ __forceinline__ float act_grad_from_out(float y, int act) {
  if (act == VRCOC_ACT_RELU) return y > 0.f ? 1.f : 0.f;
  if (act == VRCOC_ACT_LRELU) return y > 0.f ? 1.f : 0.1f;
  return 1.f;
}
__device__ __forceinline__ float silu_grad(float z) {
  const float sg = 1.0f / (1.0f + expf(-z));
  return sg * (1.0f + z * (1.0f - sg));
}

// 128-bit path of the plane kernels below: every pointer (at its plane base) aligned to 8 elements and HW a multiple of 8.  The
// scalar loops moved 2-byte words and ran at ~1 TB/s (torch profiler of a training step: chan_bwd_sums 2.6 ms, table_bwd_sums 1.3 ms).
template <typename T>
__device__ __forceinline__ bool vec8_ok(int HW, const T* a, const T* b = nullptr, const T* c = nullptr, const T* d = nullptr, const T* e = nullptr) {
  uintptr_t m = reinterpret_cast<uintptr_t>(a);
  if (b) m |= reinterpret_cast<uintptr_t>(b);
  if (c) m |= reinterpret_cast<uintptr_t>(c);
  if (d) m |= reinterpret_cast<uintptr_t>(d);
  if (e) m |= reinterpret_cast<uintptr_t>(e);
  return (HW & 7) == 0 && (m & (8 * sizeof(T) - 1)) == 0;
}

// per (b, c): { sum_p g, sum_p g * u }  with g = dy * act'(y)
template <typename T>
__global__ void __launch_bounds__(256) chan_bwd_sums_kernel(const T* __restrict__ dy, const T* __restrict__ yact, const T* __restrict__ u, int act, int C, int HW,
                                                            const float* __restrict__ zs, const float* __restrict__ zt, float* __restrict__ out) {
  __shared__ float red[64];
  const int64_t base = (int64_t)blockIdx.x * HW;
  const int c = blockIdx.x % C;
  const float z_s = zs ? zs[c] : 1.f, z_t = zt ? zt[c] : 0.f;
  float s = 0.f, su = 0.f;
  auto one = [&](float g, float uu, float ya) {
    if (act == VRCOC_ACT_SILU) g *= silu_grad(fmaf(uu, z_s, z_t));
    else if (yact) g *= act_grad_from_out(ya, act);
    else if (zs) g *= act_grad_from_out(fmaf(uu, z_s, z_t), act);          // sign of the recomputed pre-activation
    s += g; su = fmaf(g, uu, su);
  };
  if (vec8_ok<T>(HW, dy + base, u + base, yact ? yact + base : nullptr)) {
    for (int i = threadIdx.x * 8; i < HW; i += blockDim.x * 8) {
      float g[8], uu[8], ya[8];
      ld8<T>(dy + base + i, g);
      ld8<T>(u + base + i, uu);
      if (yact) ld8<T>(yact + base + i, ya);
#pragma unroll
      for (int e = 0; e < 8; ++e) one(g[e], uu[e], yact ? ya[e] : 0.f);
    }
  } else {
    for (int i = threadIdx.x; i < HW; i += blockDim.x)
      one(ldf<T>(dy + base + i), ldf<T>(u + base + i), yact ? ldf<T>(yact + base + i) : 0.f);
  }
  s = warp_sum(s); su = warp_sum(su);
  if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = s; red[32 + (threadIdx.x >> 5)] = su; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { a += red[i]; b += red[32 + i]; }
    out[2 * blockIdx.x] = a; out[2 * blockIdx.x + 1] = b;
  }
}

// out = (dy * act'(y)) * ca[c] + u * cb[c] + cd[c]  (+ extra)
template <typename T>
__global__ void __launch_bounds__(256) chan_bwd_apply_kernel(const T* __restrict__ dy, const T* __restrict__ yact, const T* __restrict__ u, const T* __restrict__ extra,
                                                             int act, const float* __restrict__ ca, const float* __restrict__ cb, const float* __restrict__ cd,
                                                             const float* __restrict__ zs, const float* __restrict__ zt, int C, int HW, T* __restrict__ out) {
  const int plane = blockIdx.y, c = plane % C;
  const float a = ca[c], b = cb ? cb[c] : 0.f, d0 = cd ? cd[c] : 0.f;
  const float z_s = zs ? zs[c] : 1.f, z_t = zt ? zt[c] : 0.f;
  const int64_t base = (int64_t)plane * HW;
  auto one = [&](float g, float uu, float ya, float ex) {
    if (act == VRCOC_ACT_SILU) g *= silu_grad(fmaf(uu, z_s, z_t));
    else if (yact) g *= act_grad_from_out(ya, act);
    else if (zs) g *= act_grad_from_out(fmaf(uu, z_s, z_t), act);
    float r = fmaf(g, a, d0);
    if (cb) r = fmaf(uu, b, r);
    return r + ex;
  };
  if (vec8_ok<T>(HW, dy + base, out + base, u ? u + base : nullptr, yact ? yact + base : nullptr, extra ? extra + base : nullptr)) {
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) * 8; i < HW; i += gridDim.x * blockDim.x * 8) {
      float g[8], uu[8], ya[8], ex[8], r[8];
      ld8<T>(dy + base + i, g);
      if (u) ld8<T>(u + base + i, uu);
      if (yact) ld8<T>(yact + base + i, ya);
      if (extra) ld8<T>(extra + base + i, ex);
#pragma unroll
      for (int e = 0; e < 8; ++e) r[e] = one(g[e], u ? uu[e] : 0.f, yact ? ya[e] : 0.f, extra ? ex[e] : 0.f);
      st8<T>(out + base + i, r);
    }
    return;
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x)
    stf<T>(out + base + i, one(ldf<T>(dy + base + i), u ? ldf<T>(u + base + i) : 0.f, yact ? ldf<T>(yact + base + i) : 0.f,
                               extra ? ldf<T>(extra + base + i) : 0.f));
}

__device__ __forceinline__ void decode_minmax(const uint32_t* mm, float& mn, float& mx) {
  mx = __uint_as_float(mm[0]);
  mn = __uint_as_float(~mm[1]);
}

// ImageEnhanceByRadar tail backward (vr_coc.py:314: y = (1 + (k - mn)/(mx - mn)) * image), given dyv = dL/dy:
//   dimage = dyv * (1 + kn),  dk = dyv * image / (mx - mn)  (the direct path),
//   part[plane] = { sum dkn, sum dkn * k, #(k == mn), #(k == mx) }  with dkn = dyv * image   (for the min / max paths)
template <typename T>
__global__ void __launch_bounds__(256) img_enh_bwd_kernel(const T* __restrict__ dyv, const T* __restrict__ image, const T* __restrict__ k,
                                                          const uint32_t* __restrict__ minmax, int HW, T* __restrict__ dimage, T* __restrict__ dk,
                                                          float* __restrict__ part) {
  __shared__ float red[4][8];
  float mn, mx;
  decode_minmax(minmax, mn, mx);
  const float r = 1.0f / (mx - mn);
  const int64_t base = (int64_t)blockIdx.x * HW;
  float s0 = 0.f, s1 = 0.f, c0 = 0.f, c1 = 0.f;
  auto one = [&](float g, float im, float kk, float& di, float& dkv) {
    const float kn = (kk - mn) * r, dkn = g * im;
    di = g * (1.0f + kn);
    dkv = dkn * r;
    s0 += dkn; s1 = fmaf(dkn, kk, s1);
    c0 += (kk == mn) ? 1.f : 0.f; c1 += (kk == mx) ? 1.f : 0.f;
  };
  if (vec8_ok<T>(HW, dyv + base, image + base, k + base, dimage + base, dk + base)) {
    for (int i = threadIdx.x * 8; i < HW; i += blockDim.x * 8) {
      float g[8], im[8], kk[8], di[8], dkv[8];
      ld8<T>(dyv + base + i, g);
      ld8<T>(image + base + i, im);
      ld8<T>(k + base + i, kk);
#pragma unroll
      for (int e = 0; e < 8; ++e) one(g[e], im[e], kk[e], di[e], dkv[e]);
      st8<T>(dimage + base + i, di);
      st8<T>(dk + base + i, dkv);
    }
  } else {
    for (int i = threadIdx.x; i < HW; i += blockDim.x) {
      float di, dkv;
      one(ldf<T>(dyv + base + i), ldf<T>(image + base + i), ldf<T>(k + base + i), di, dkv);
      stf<T>(dimage + base + i, di);
      stf<T>(dk + base + i, dkv);
    }
  }
  s0 = warp_sum(s0); s1 = warp_sum(s1); c0 = warp_sum(c0); c1 = warp_sum(c1);
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { red[0][w] = s0; red[1][w] = s1; red[2][w] = c0; red[3][w] = c1; }
  __syncthreads();
  if (threadIdx.x < 4) {
    float t = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[threadIdx.x][i];
    part[4 * blockIdx.x + threadIdx.x] = t;
  }
}

// dk += (k == mn) * coef[0] + (k == mx) * coef[1]: the gradient that reaches k through the global min / max (evenly over ties)
template <typename T>
__global__ void __launch_bounds__(256) minmax_scatter_kernel(const T* __restrict__ k, T* __restrict__ dk, const uint32_t* __restrict__ minmax,
                                                             const float* __restrict__ coef, int64_t n) {
  float mn, mx;
  decode_minmax(minmax, mn, mx);
  const float a = coef[0], b = coef[1];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float kk = ldf<T>(k + i);
    if (kk == mn || kk == mx) stf<T>(dk + i, ldf<T>(dk + i) + (kk == mn ? a : 0.f) + (kk == mx ? b : 0.f));
  }
}

// Adjoint of the bilinear upsample with align_corners=True (CoCUpsample, coc_fpn_dual.py:19-22): gather form, one thread per INPUT
// pixel; the output rows / columns that touch it are found with the forward kernel's own index arithmetic (stats.cu), so the
// weights are bit-identical to the forward ones.
template <typename T>
__global__ void __launch_bounds__(256) upsample_bilinear_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, int H, int W, int Ho, int Wo,
                                                                    float sy, float sx, int64_t total) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int x = (int)(i % W), y = (int)((i / W) % H);
  const int64_t plane = i / ((int64_t)W * H);
  // candidate output rows: fy = oy*sy in (y-1, y+1)
  int oy0 = sy > 0.f ? (int)floorf((float)(y - 1) / sy) : 0, oy1 = sy > 0.f ? (int)ceilf((float)(y + 1) / sy) : Ho - 1;
  int ox0 = sx > 0.f ? (int)floorf((float)(x - 1) / sx) : 0, ox1 = sx > 0.f ? (int)ceilf((float)(x + 1) / sx) : Wo - 1;
  oy0 = max(oy0, 0); oy1 = min(oy1, Ho - 1); ox0 = max(ox0, 0); ox1 = min(ox1, Wo - 1);
  const T* base = dy + plane * (int64_t)Ho * Wo;
  float acc = 0.f;
  for (int oy = oy0; oy <= oy1; ++oy) {
    const float fy = oy * sy;
    int y0 = (int)fy;
    if (y0 > H - 1) y0 = H - 1;
    const int y1 = min(y0 + 1, H - 1);
    const float wy = fy - (float)y0;
    const float cy = (y0 == y ? 1.f - wy : 0.f) + (y1 == y ? wy : 0.f);
    if (cy == 0.f) continue;
    float row = 0.f;
    for (int ox = ox0; ox <= ox1; ++ox) {
      const float fx = ox * sx;
      int x0 = (int)fx;
      if (x0 > W - 1) x0 = W - 1;
      const int x1 = min(x0 + 1, W - 1);
      const float wx = fx - (float)x0;
      const float cx = (x0 == x ? 1.f - wx : 0.f) + (x1 == x ? wx : 0.f);
      if (cx != 0.f) row = fmaf(cx, ldf<T>(base + (int64_t)oy * Wo + ox), row);
    }
    acc = fmaf(cy, row, acc);
  }
  stf<T>(dx + i, acc);
}

// YOLOX box decode of the three detection maps (utils/utils_bbox.py:32-84) in one launch: out[b][n][c], n over levels then cells
// (row-major), c over the 5 + num_classes channels:  xy = (xy + cell) * stride / input,  wh = exp(wh) * stride / input,  rest = sigmoid.
struct DecodeArgs { const void* maps[3]; int H[3], W[3], off[3]; float sy[3], sx[3]; };
template <typename T>
__global__ void __launch_bounds__(256) decode_outputs_kernel(DecodeArgs a, int B, int CH, int N, float inv_h, float inv_w, float* __restrict__ out) {
  const int64_t total = (int64_t)B * N;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(i % N);
    const int64_t b = i / N;
    const int l = n >= a.off[2] ? 2 : (n >= a.off[1] ? 1 : 0);
    const int cell = n - a.off[l];
    const int gy = cell / a.W[l], gx = cell % a.W[l];
    const int64_t plane = (int64_t)a.H[l] * a.W[l];
    const T* src = reinterpret_cast<const T*>(a.maps[l]) + b * CH * plane + cell;
    float* o = out + i * CH;
    o[0] = (ldf<T>(src) + (float)gx) * a.sx[l] * inv_w;
    o[1] = (ldf<T>(src + plane) + (float)gy) * a.sy[l] * inv_h;
    o[2] = expf(ldf<T>(src + 2 * plane)) * a.sx[l] * inv_w;
    o[3] = expf(ldf<T>(src + 3 * plane)) * a.sy[l] * inv_h;
    for (int c = 4; c < CH; ++c) o[c] = 1.0f / (1.0f + expf(-ldf<T>(src + c * plane)));
  }
}

// ---- backward of the table-driven prologue  z = x * s * h(x) * e,  h(x) = sigmoid(ga*x + gc)  (ShuffleAttention gates + ECA scale of
//      RadarEnhanceByImage, vr_coc.py:344-350; shuffle_attention.py:48-72; eca.py:16-22).  Source-channel order; dz is read at the logical
//      channel kidx[c] of source channel c.  The statistics chain (channel means / variances -> gates -> ECA) is O(B*C) algebra on the
//      six sums below; its results come back as the four coefficients of the apply pass. ------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) table_bwd_sums_kernel(const T* __restrict__ dz, const T* __restrict__ x, const int32_t* __restrict__ kidx,
                                                             const float* __restrict__ gate, int C, int K, int HW, float* __restrict__ out) {
  __shared__ float red[6][8];
  const int b = blockIdx.x / C, c = blockIdx.x % C;
  const int k = kidx ? kidx[c] : c;
  const float ga = gate ? gate[2 * blockIdx.x] : 0.f, gc = gate ? gate[2 * blockIdx.x + 1] : 0.f;
  const T* dzp = dz + ((int64_t)b * K + k) * HW;
  const T* xp = x + (int64_t)blockIdx.x * HW;
  float a1 = 0.f, a2 = 0.f, a3 = 0.f, j1 = 0.f, j2 = 0.f, j3 = 0.f;
  auto one = [&](float g, float xv) {
    const float h = gate ? 1.0f / (1.0f + expf(-fmaf(ga, xv, gc))) : 1.0f;
    const float xh = xv * h, xd = xv * h * (1.0f - h);
    a1 = fmaf(g, xh, a1); a2 = fmaf(g * xv, xd, a2); a3 = fmaf(g, xd, a3);
    j1 += xh; j2 = fmaf(xv, xd, j2); j3 += xd;
  };
  if (vec8_ok<T>(HW, dzp, xp)) {
    for (int i = threadIdx.x * 8; i < HW; i += blockDim.x * 8) {
      float g[8], xv[8];
      ld8<T>(dzp + i, g);
      ld8<T>(xp + i, xv);
#pragma unroll
      for (int e = 0; e < 8; ++e) one(g[e], xv[e]);
    }
  } else {
    for (int i = threadIdx.x; i < HW; i += blockDim.x) one(ldf<T>(dzp + i), ldf<T>(xp + i));
  }
  float v[6] = {a1, a2, a3, j1, j2, j3};
  const int w = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < 6; ++q) {
    v[q] = warp_sum(v[q]);
    if ((threadIdx.x & 31) == 0) red[q][w] = v[q];
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    float t = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[threadIdx.x][i];
    out[6 * (int64_t)blockIdx.x + threadIdx.x] = t;
  }
}

// dx = (dz*cd + cj) * (h + x*ga*h*(1-h)) + c1 + 2*x*c2 (+ extra),  coef[b][c] = {cd, cj, c1, c2}
template <typename T>
__global__ void __launch_bounds__(256) table_bwd_apply_kernel(const T* __restrict__ dz, const T* __restrict__ x, const T* __restrict__ extra,
                                                              const int32_t* __restrict__ kidx, const float* __restrict__ gate,
                                                              const float* __restrict__ coef, int C, int K, int HW, T* __restrict__ out) {
  const int plane = blockIdx.y, b = plane / C, c = plane % C;
  const int k = kidx ? kidx[c] : c;
  const float ga = gate ? gate[2 * plane] : 0.f, gc = gate ? gate[2 * plane + 1] : 0.f;
  const float cd = coef[4 * plane], cj = coef[4 * plane + 1], c1 = coef[4 * plane + 2], c2 = coef[4 * plane + 3];
  const T* dzp = dz + ((int64_t)b * K + k) * HW;
  const int64_t base = (int64_t)plane * HW;
  auto one = [&](float g, float xv, float ex) {
    float t = 1.0f;
    if (gate) {
      const float h = 1.0f / (1.0f + expf(-fmaf(ga, xv, gc)));
      t = h + xv * ga * h * (1.0f - h);
    }
    return fmaf(fmaf(g, cd, cj), t, fmaf(2.0f * xv, c2, c1)) + ex;
  };
  if (vec8_ok<T>(HW, dzp, x + base, out + base, extra ? extra + base : nullptr)) {
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) * 8; i < HW; i += gridDim.x * blockDim.x * 8) {
      float g[8], xv[8], ex[8], r[8];
      ld8<T>(dzp + i, g);
      ld8<T>(x + base + i, xv);
      if (extra) ld8<T>(extra + base + i, ex);
#pragma unroll
      for (int e = 0; e < 8; ++e) r[e] = one(g[e], xv[e], extra ? ex[e] : 0.f);
      st8<T>(out + base + i, r);
    }
    return;
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x)
    stf<T>(out + base + i, one(ldf<T>(dzp + i), ldf<T>(x + base + i), extra ? ldf<T>(extra + base + i) : 0.f));
}

// ---- depthwise 3x3 / stride 1 / pad 1 weight gradient (the head's DWConv towers, normal_conv.py:23-33, decouplehead.py:24-37):
//      dW[c][ky][kx] = sum_{b,y,x} dy[b][c][y][x] * x[b][c][y+ky-1][x+kx-1],  db[c] = sum dy.  One block per channel, a thread = 8
//      columns of one row per step (dy vector, three x vectors + halos), fixed-order block reduction (deterministic).
template <typename T>
__global__ void __launch_bounds__(256) dwconv3_wgrad_kernel(const T* __restrict__ x, const T* __restrict__ dy, int B, int C, int H, int W,
                                                            float* __restrict__ dW, float* __restrict__ db) {
  __shared__ float red[10][8];
  const int c = blockIdx.x;
  const int w8 = W >> 3;
  const int per = H * w8;                                             // 8-column row segments per (b, c) plane
  float acc[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) acc[i] = 0.f;
  for (int t = threadIdx.x; t < B * per; t += blockDim.x) {
    const int b = t / per, r = t - b * per, y = r / w8, x0 = (r - y * w8) * 8;
    const int64_t pl = ((int64_t)b * C + c) * H * W;
    float g[8];
    ld8<T>(dy + pl + (int64_t)y * W + x0, g);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[9] += g[j];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = y + ky - 1;
      if (iy < 0 || iy >= H) continue;
      const T* row = x + pl + (int64_t)iy * W;
      float v[10], mid[8];
      ld8<T>(row + x0, mid);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j + 1] = mid[j];
      v[0] = x0 > 0 ? ldf<T>(row + x0 - 1) : 0.f;
      v[9] = x0 + 8 < W ? ldf<T>(row + x0 + 8) : 0.f;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[ky * 3 + kx] = fmaf(g[j], v[j + kx], acc[ky * 3 + kx]);
    }
  }
  const int w = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const float v = warp_sum(acc[i]);
    if ((threadIdx.x & 31) == 0) red[i][w] = v;
  }
  __syncthreads();
  if (threadIdx.x < 10) {
    float t = 0.f;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += red[threadIdx.x][k];
    if (threadIdx.x < 9) dW[c * 9 + threadIdx.x] = t;
    else if (db) db[c] = t;
  }
}

// ---- BatchNorm2d bookkeeping of the training path as two tiny kernels (normal_conv.py:45-49, vr_coc.py:315,341,356).  The host
//      side did this algebra with torch ops on [C]-sized tensors: ~22 launches per BatchNorm forward (batch statistics, running-stat
//      update, folded scale / shift) and ~30 per backward, ~65 BatchNorms per step = ~3000 of the 6800 launches of a training step.
// batch statistics from the per-(b, c) sums -> mean / biased variance (fp64), running-stat update of nn.BatchNorm2d (unbiased
// variance, momentum), folded affine BN(x) = x * scale + shift
__global__ void __launch_bounds__(128) bn_stats_kernel(const float* __restrict__ cs, int B, int C, double count, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, float eps, float momentum, float* __restrict__ rmean,
                                                       float* __restrict__ rvar, long long* __restrict__ nbt, float* __restrict__ scale,
                                                       float* __restrict__ shift, double* __restrict__ mean_out, double* __restrict__ var_out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && nbt) *nbt += 1;
  if (c >= C) return;
  double s = 0.0, s2 = 0.0;
  for (int b = 0; b < B; ++b) {
    s += (double)cs[((int64_t)b * C + c) * 2];
    s2 += (double)cs[((int64_t)b * C + c) * 2 + 1];
  }
  const double mean = s / count;
  double var = s2 / count - mean * mean;
  var = var > 0.0 ? var : 0.0;
  if (rmean && momentum >= 0.f) {
    rmean[c] = rmean[c] * (1.0f - momentum) + momentum * (float)mean;
    const double unb = var * (count / (count > 1.0 ? count - 1.0 : 1.0));
    rvar[c] = rvar[c] * (1.0f - momentum) + momentum * (float)unb;
  }
  if (mean_out) { mean_out[c] = mean; var_out[c] = var; }
  if (scale) {
    float sc = 1.0f / sqrtf((float)var + eps);
    if (gamma) sc *= gamma[c];
    scale[c] = sc;
    shift[c] = -(float)mean * sc + (beta ? beta[c] : 0.f);
  }
}

// backward coefficients of y = act(BN(u)) from S[b][c] = {sum g, sum g*u}: du = g*ca + u*cb + cd; dgamma, dbeta.  With sums == NULL
// only the recomputation affine of the pre-activation, z = u * zs + zt, is produced (needed BEFORE the sums pass).
__global__ void __launch_bounds__(128) bn_bwd_coef_kernel(const float* __restrict__ sums, int B, int C, const double* __restrict__ mean,
                                                          const double* __restrict__ var, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          float eps, double N, int training, float* __restrict__ zs, float* __restrict__ zt,
                                                          float* __restrict__ ca, float* __restrict__ cb, float* __restrict__ cd,
                                                          float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double m = mean[c], rstd = 1.0 / sqrt(var[c] + (double)eps), g = gamma ? (double)gamma[c] : 1.0;
  if (zs) {
    zs[c] = (float)(g * rstd);
    zt[c] = (float)((beta ? (double)beta[c] : 0.0) - m * g * rstd);
  }
  if (!sums) return;
  double s1 = 0.0, s2 = 0.0;
  for (int b = 0; b < B; ++b) {
    s1 += (double)sums[((int64_t)b * C + c) * 2];
    s2 += (double)sums[((int64_t)b * C + c) * 2 + 1];
  }
  const double sgx = (s2 - m * s1) * rstd;
  ca[c] = (float)(g * rstd);
  if (training) {
    const double b_ = -g * rstd * rstd * sgx / N;
    cb[c] = (float)b_;
    cd[c] = (float)(-g * rstd * s1 / N - b_ * m);
  }
  dgamma[c] = (float)sgx;
  dbeta[c] = (float)s1;
}

}  // namespace
}  // namespace vrcoc

using namespace vrcoc;

extern "C" int vrcoc_dwconv3_wgrad(const void* x, const void* dy, int dtype, int B, int C, int H, int W, float* dW, float* db, void* stream) {
  VRCOC_REQUIRE(x && dy && dW && B > 0 && C > 0 && H > 0 && W > 0, "dwconv3_wgrad: bad argument");
  VRCOC_REQUIRE((W & 7) == 0, "dwconv3_wgrad: the map width (%d) must be a multiple of 8", W);
  const int es = dtype == VRCOC_F32 ? 4 : 2;
  VRCOC_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy)) & (8 * es - 1)) == 0, "dwconv3_wgrad: tensors must be aligned to 8 elements");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == VRCOC_BF16) dwconv3_wgrad_kernel<__nv_bfloat16><<<C, 256, 0, st>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, B, C, H, W, dW, db);
  else if (dtype == VRCOC_F32) dwconv3_wgrad_kernel<float><<<C, 256, 0, st>>>((const float*)x, (const float*)dy, B, C, H, W, dW, db);
  else return fail(VRCOC_EINVAL, "dwconv3_wgrad: unsupported dtype %d", dtype);
  return check_launch("dwconv3_wgrad");
}

extern "C" int vrcoc_bn_stats(const float* chan_sums, int B, int C, double count, const float* gamma, const float* beta, float eps, float momentum,
                              float* running_mean, float* running_var, long long* num_batches_tracked, float* scale, float* shift,
                              double* mean_out, double* var_out, void* stream) {
  VRCOC_REQUIRE(chan_sums && B > 0 && C > 0 && count > 0, "bn_stats: bad argument");
  VRCOC_REQUIRE((scale == nullptr) == (shift == nullptr) && (mean_out == nullptr) == (var_out == nullptr) &&
                (running_mean == nullptr) == (running_var == nullptr), "bn_stats: outputs come in pairs");
  bn_stats_kernel<<<(unsigned)cdiv(C, 128), 128, 0, (cudaStream_t)stream>>>(chan_sums, B, C, count, gamma, beta, eps, momentum, running_mean,
                                                                            running_var, num_batches_tracked, scale, shift, mean_out, var_out);
  return check_launch("bn_stats");
}

extern "C" int vrcoc_bn_bwd_coef(const float* sums, int B, int C, const double* mean, const double* var, const float* gamma, const float* beta,
                                 float eps, double N, int training, float* zs, float* zt, float* ca, float* cb, float* cd, float* dgamma,
                                 float* dbeta, void* stream) {
  VRCOC_REQUIRE(mean && var && C > 0 && N > 0, "bn_bwd_coef: bad argument");
  VRCOC_REQUIRE((zs == nullptr) == (zt == nullptr), "bn_bwd_coef: zs / zt come as a pair");
  VRCOC_REQUIRE(!sums || (B > 0 && ca && dgamma && dbeta && (!training || (cb && cd))), "bn_bwd_coef: null coefficient output");
  bn_bwd_coef_kernel<<<(unsigned)cdiv(C, 128), 128, 0, (cudaStream_t)stream>>>(sums, B, C, mean, var, gamma, beta, eps, N, training, zs, zt, ca, cb,
                                                                               cd, dgamma, dbeta);
  return check_launch("bn_bwd_coef");
}

extern "C" int vrcoc_decode_outputs(const void* p3, const void* p4, const void* p5, int dtype, int B, int channels, int h3, int w3, int h4,
                                    int w4, int h5, int w5, int input_h, int input_w, float* out, void* stream) {
  VRCOC_REQUIRE(p3 && p4 && p5 && out && B > 0 && channels >= 5 && h3 > 0 && w3 > 0 && h4 > 0 && w4 > 0 && h5 > 0 && w5 > 0 && input_h > 0 && input_w > 0,
                "decode_outputs: bad argument");
  DecodeArgs a;
  a.maps[0] = p3; a.maps[1] = p4; a.maps[2] = p5;
  a.H[0] = h3; a.W[0] = w3; a.H[1] = h4; a.W[1] = w4; a.H[2] = h5; a.W[2] = w5;
  a.off[0] = 0; a.off[1] = h3 * w3; a.off[2] = h3 * w3 + h4 * w4;
  // the reference uses ONE stride per level, input_shape[0] / h, for both axes (utils_bbox.py:63)
  for (int l = 0; l < 3; ++l) { a.sy[l] = (float)input_h / (float)a.H[l]; a.sx[l] = a.sy[l]; }
  const int N = h3 * w3 + h4 * w4 + h5 * w5;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned blocks = (unsigned)(cdiv((int64_t)B * N, 256) < 1184 ? cdiv((int64_t)B * N, 256) : 1184);
  if (dtype == VRCOC_BF16) decode_outputs_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(a, B, channels, N, 1.0f / input_h, 1.0f / input_w, out);
  else decode_outputs_kernel<float><<<blocks, 256, 0, st>>>(a, B, channels, N, 1.0f / input_h, 1.0f / input_w, out);
  return check_launch("decode_outputs");
}

extern "C" int vrcoc_upsample_bilinear_bwd(const void* dy, void* dx, int dtype, int planes, int H, int W, int Ho, int Wo, void* stream) {
  VRCOC_REQUIRE(dy && dx && planes > 0 && H > 0 && W > 0 && Ho > 0 && Wo > 0, "upsample_bilinear_bwd: bad argument");
  const float sy = Ho > 1 ? (float)(H - 1) / (float)(Ho - 1) : 0.f, sx = Wo > 1 ? (float)(W - 1) / (float)(Wo - 1) : 0.f;
  const int64_t total = (int64_t)planes * H * W;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned blocks = (unsigned)cdiv(total, 256);
  if (dtype == VRCOC_BF16)
    upsample_bilinear_bwd_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)dy, (__nv_bfloat16*)dx, H, W, Ho, Wo, sy, sx, total);
  else
    upsample_bilinear_bwd_kernel<float><<<blocks, 256, 0, st>>>((const float*)dy, (float*)dx, H, W, Ho, Wo, sy, sx, total);
  return check_launch("upsample_bilinear_bwd");
}

extern "C" int vrcoc_table_bwd_sums(const void* dz, const void* x, int dtype, const int32_t* kidx, const float* gate, int B, int C, int K,
                                    int HW, float* out, void* stream) {
  VRCOC_REQUIRE(dz && x && out && B > 0 && C > 0 && K >= C && HW > 0, "table_bwd_sums: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == VRCOC_BF16)
    table_bwd_sums_kernel<__nv_bfloat16><<<B * C, 256, 0, st>>>((const __nv_bfloat16*)dz, (const __nv_bfloat16*)x, kidx, gate, C, K, HW, out);
  else
    table_bwd_sums_kernel<float><<<B * C, 256, 0, st>>>((const float*)dz, (const float*)x, kidx, gate, C, K, HW, out);
  return check_launch("table_bwd_sums");
}

extern "C" int vrcoc_table_bwd_apply(const void* dz, const void* x, const void* extra, void* out, int dtype, const int32_t* kidx,
                                     const float* gate, const float* coef, int B, int C, int K, int HW, void* stream) {
  VRCOC_REQUIRE(dz && x && out && coef && B > 0 && C > 0 && K >= C && HW > 0, "table_bwd_apply: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)(HW >= 16384 ? cdiv(HW, 4096) : 1), (unsigned)(B * C));
  if (dtype == VRCOC_BF16)
    table_bwd_apply_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)dz, (const __nv_bfloat16*)x, (const __nv_bfloat16*)extra, kidx, gate,
                                                                 coef, C, K, HW, (__nv_bfloat16*)out);
  else
    table_bwd_apply_kernel<float><<<grid, 256, 0, st>>>((const float*)dz, (const float*)x, (const float*)extra, kidx, gate, coef, C, K, HW, (float*)out);
  return check_launch("table_bwd_apply");
}

extern "C" int vrcoc_img_enh_bwd(const void* dyv, const void* image, const void* k, int dtype, const uint32_t* minmax, int B, int C, int HW,
                                 void* dimage, void* dk, float* part, void* stream) {
  VRCOC_REQUIRE(dyv && image && k && minmax && dimage && dk && part && B > 0 && C > 0 && HW > 0, "img_enh_bwd: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == VRCOC_BF16)
    img_enh_bwd_kernel<__nv_bfloat16><<<B * C, 256, 0, st>>>((const __nv_bfloat16*)dyv, (const __nv_bfloat16*)image, (const __nv_bfloat16*)k, minmax, HW,
                                                             (__nv_bfloat16*)dimage, (__nv_bfloat16*)dk, part);
  else
    img_enh_bwd_kernel<float><<<B * C, 256, 0, st>>>((const float*)dyv, (const float*)image, (const float*)k, minmax, HW, (float*)dimage, (float*)dk, part);
  return check_launch("img_enh_bwd");
}

extern "C" int vrcoc_minmax_scatter(const void* k, void* dk, int dtype, const uint32_t* minmax, const float* coef, int64_t n, void* stream) {
  VRCOC_REQUIRE(k && dk && minmax && coef && n > 0, "minmax_scatter: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned blocks = (unsigned)(cdiv(n, 256) < 1184 ? cdiv(n, 256) : 1184);
  if (dtype == VRCOC_BF16)
    minmax_scatter_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)k, (__nv_bfloat16*)dk, minmax, coef, n);
  else
    minmax_scatter_kernel<float><<<blocks, 256, 0, st>>>((const float*)k, (float*)dk, minmax, coef, n);
  return check_launch("minmax_scatter");
}

extern "C" int vrcoc_gn_bwd_coef(const float* s, const double* gn_sums, const float* gamma, float eps, int B, int C, int HW, float* a,
                                 float* bb, float* cc, float* dgamma, float* dbeta, float* workspace, void* stream) {
  VRCOC_REQUIRE(s && gn_sums && gamma && a && bb && cc && dgamma && dbeta && workspace && B > 0 && C > 0 && HW > 0, "gn_bwd_coef: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  gn_bwd_coef_sample_kernel<<<B, 256, 0, st>>>(s, gn_sums, gamma, eps, C, HW, a, bb, cc, workspace);
  gn_bwd_coef_chan_kernel<<<(C + 127) / 128, 128, 0, st>>>(s, workspace, B, C, dgamma, dbeta);
  return check_launch("gn_bwd_coef");
}

extern "C" int vrcoc_proj_res_bwd_coef(const float* G, const float* sum_dy, const void* W, int w_dtype, const float* bias, const float* ls,
                                       int O, int K, float* dW, float* db, float* dls, void* wt, void* stream) {
  VRCOC_REQUIRE(G && sum_dy && W && dW && wt && O > 0 && K > 0, "proj_res_bwd_coef: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (w_dtype == VRCOC_BF16)
    proj_res_bwd_coef_kernel<__nv_bfloat16><<<O, 128, 0, st>>>(G, sum_dy, (const __nv_bfloat16*)W, bias, ls, O, K, dW, db, dls, (__nv_bfloat16*)wt);
  else
    proj_res_bwd_coef_kernel<float><<<O, 128, 0, st>>>(G, sum_dy, (const float*)W, bias, ls, O, K, dW, db, dls, (float*)wt);
  return check_launch("proj_res_bwd_coef");
}

extern "C" int vrcoc_col2im(const void* dcol, void* dx, int dtype, int B, int C, int H, int W, int kh, int kw, int stride, int pad, int dil,
                            void* stream) {
  VRCOC_REQUIRE(dcol && dx && B > 0 && C > 0 && H > 0 && W > 0 && kh > 0 && kw > 0 && stride > 0 && pad >= 0, "col2im: bad argument");
  if (dil <= 0) dil = 1;
  const int Ho = (H + 2 * pad - dil * (kh - 1) - 1) / stride + 1, Wo = (W + 2 * pad - dil * (kw - 1) - 1) / stride + 1;
  VRCOC_REQUIRE(Ho > 0 && Wo > 0, "col2im: empty output");
  const int64_t total = (int64_t)B * C * H * W;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned blocks = (unsigned)cdiv(total, 256);
  if (dtype == VRCOC_BF16 && kh == 3 && kw == 3 && stride == 1 && pad == 1 && dil == 1 && (W & 7) == 0 &&
      ((reinterpret_cast<uintptr_t>(dcol) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0) {
    const int64_t total8 = total >> 3;
    col2im3_s1_bf16_kernel<<<(unsigned)cdiv(total8, 256), 256, 0, st>>>((const __nv_bfloat16*)dcol, (__nv_bfloat16*)dx, C, H, W, total8);
    return check_launch("col2im3");
  }
  if (dtype == VRCOC_BF16 && kh == 4 && kw == 4 && stride == 4 && pad == 0 && dil == 1 && (W & 7) == 0 && (H & 3) == 0 && Wo * 4 == W && Ho * 4 == H &&
      ((reinterpret_cast<uintptr_t>(dcol) & 3) | (reinterpret_cast<uintptr_t>(dx) & 15)) == 0) {
    const int64_t total8 = total >> 3;
    col2im4_s4_bf16_kernel<<<(unsigned)cdiv(total8, 256), 256, 0, st>>>((const __nv_bfloat16*)dcol, (__nv_bfloat16*)dx, C, H, W, Ho, Wo, total8);
    return check_launch("col2im4s4");
  }
  if (dtype == VRCOC_BF16 && kh == 3 && kw == 3 && stride == 2 && pad == 1 && dil == 1 && (W & 7) == 0 && (H & 1) == 0 && Wo * 2 == W && Ho * 2 == H &&
      ((reinterpret_cast<uintptr_t>(dcol) & 7) | (reinterpret_cast<uintptr_t>(dx) & 15)) == 0) {
    const int64_t total8 = total >> 3;
    col2im3_s2_bf16_kernel<<<(unsigned)cdiv(total8, 256), 256, 0, st>>>((const __nv_bfloat16*)dcol, (__nv_bfloat16*)dx, C, H, W, Ho, Wo, total8);
    return check_launch("col2im3s2");
  }
  if (dtype == VRCOC_BF16)
    col2im_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)dcol, (__nv_bfloat16*)dx, C, H, W, kh, kw, stride, pad, dil, Ho, Wo, total);
  else
    col2im_kernel<float><<<blocks, 256, 0, st>>>((const float*)dcol, (float*)dx, C, H, W, kh, kw, stride, pad, dil, Ho, Wo, total);
  return check_launch("col2im");
}

extern "C" int vrcoc_chan_bwd_sums(const void* dy, const void* y_act, const void* u, int dtype, int act, const float* z_scale,
                                   const float* z_shift, int B, int C, int HW, float* out, void* stream) {
  VRCOC_REQUIRE(dy && u && out && B > 0 && C > 0 && HW > 0, "chan_bwd_sums: bad argument");
  VRCOC_REQUIRE(act != VRCOC_ACT_GELU, "chan_bwd_sums: GELU is handled by vrcoc_gelu_bwd");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == VRCOC_BF16)
    chan_bwd_sums_kernel<__nv_bfloat16><<<B * C, 256, 0, st>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)y_act, (const __nv_bfloat16*)u, act, C, HW,
                                                               z_scale, z_shift, out);
  else
    chan_bwd_sums_kernel<float><<<B * C, 256, 0, st>>>((const float*)dy, (const float*)y_act, (const float*)u, act, C, HW, z_scale, z_shift, out);
  return check_launch("chan_bwd_sums");
}

extern "C" int vrcoc_chan_bwd_apply(const void* dy, const void* y_act, const void* u, const void* extra, void* out, int dtype, int act,
                                    const float* ca, const float* cb, const float* cd, const float* z_scale, const float* z_shift, int B,
                                    int C, int HW, void* stream) {
  VRCOC_REQUIRE(dy && out && ca && B > 0 && C > 0 && HW > 0 && (!cb || u) && (act != VRCOC_ACT_SILU || u), "chan_bwd_apply: bad argument");
  VRCOC_REQUIRE(act != VRCOC_ACT_GELU, "chan_bwd_apply: GELU is handled by vrcoc_gelu_bwd");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)(HW >= 16384 ? cdiv(HW, 4096) : 1), (unsigned)(B * C));
  if (dtype == VRCOC_BF16)
    chan_bwd_apply_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)y_act, (const __nv_bfloat16*)u,
                                                                (const __nv_bfloat16*)extra, act, ca, cb, cd, z_scale, z_shift, C, HW, (__nv_bfloat16*)out);
  else
    chan_bwd_apply_kernel<float><<<grid, 256, 0, st>>>((const float*)dy, (const float*)y_act, (const float*)u, (const float*)extra, act, ca, cb, cd, z_scale,
                                                       z_shift, C, HW, (float*)out);
  return check_launch("chan_bwd_apply");
}
