// Pieces shared by the CUDA-core and the tcgen05 variants of the convolution-as-GEMM engine.
#pragma once
#include "common.cuh"

namespace vrcoc {

// Kernel-side copy of vrcoc_conv_desc plus derived quantities (passed by value, < 4 KB of parameter space).
struct ConvArgs {
  int B, H_in, W_in, H_out, W_out, C0, C1, Cin, O, kh, kw, stride, pad, dil;
  int k_order;  // 0: k = c*kh*kw + tap (PyTorch conv weight layout); 1: k = tap*Cin + c (tap-major, weight permuted by caller)
  int K;        // Cin * kh * kw
  int P_in, P_out;
  const void* src0; int src0_dtype; int64_t src0_bstride;
  const void* src1; int src1_dtype; int64_t src1_bstride;
  const int32_t* chan_src;
  const double* gn_sums; const float* gn_gamma; const float* gn_beta; float gn_eps;
  const float* gn_fold_k1;   // GroupNorm folded into the weights ([O][2K] = hi|lo), statistics applied in the epilogue (vrcoc.h)
  const float* table; int has_gate;
  const void* weight; int weight_dtype;
  const float* e_scale; const float* e_shift; int act; const float* post_scale;
  const void* res; int res_dtype;
  const float* f_scale; const float* f_shift;
  void* out; int out_dtype; void* out2; int out2_dtype; int O_split;
  double* out_sample_sums; uint32_t* out_minmax;
  int fast1x1;  // 1x1, stride 1, no pad, P % 8 == 0, 16-byte aligned sources: vectorised slab loads
  int vec_out;  // P_out % 8 == 0 and aligned outputs: vectorised stores
  // row-tap mode (vrcoc.h, k_order 2): src0 is the horizontal-tap expansion [B][rt_C][H][W] of a k x k convolution's input
  // (vrcoc_im2col_rows); the GEMM runs as a 1x1 projection over K = rt_taps * rt_C virtual channels whose slab (tap ky, c0) is the
  // TMA box of src0 shifted by (ky - rt_taps/2) * rt_dil rows (zero fill outside the map).  0 taps = off.
  int rt_taps, rt_C, rt_dil, rt_W;
};

// GroupNorm(1, C) statistics of sample b from the slot-wise partial sums: every warp reduces the 32 slots with shuffles
// (lane = slot), so all threads get the same (mean, rstd) without shared memory.
__device__ __forceinline__ void gn_mean_rstd(const double* gn_sums, int b, double cnt, float eps, float& mu, float& rstd) {
  const int lane = threadIdx.x & 31;
  double s = gn_sums[((int64_t)b * VRCOC_STAT_SLOTS + lane) * 2];
  double s2 = gn_sums[((int64_t)b * VRCOC_STAT_SLOTS + lane) * 2 + 1];
  s = warp_sum(s);
  s2 = warp_sum(s2);
  const double mean = s / cnt;
  double var = s2 / cnt - mean * mean;
  var = var > 0.0 ? var : 0.0;
  rstd = (float)(1.0 / sqrt(var + (double)eps));
  mu = (float)mean;
}

// Per-CTA prologue table tab[c] = {scale, shift, gate_a, gate_c} for the CTA's sample b.
__device__ __forceinline__ void build_prologue_table(const ConvArgs& a, int b, float4* tab) {
  if (a.gn_sums) {
    // GroupNorm(1, C): per-sample mean / rstd (vr_coc.py:105-111), folded with the per-channel affine so the slab
    // loader does one FMA per element.
    float mu, rstd;
    gn_mean_rstd(a.gn_sums, b, (double)a.C0 * (double)a.P_in, a.gn_eps, mu, rstd);
    for (int c = threadIdx.x; c < a.Cin; c += blockDim.x) {
      float sc = rstd * a.gn_gamma[c];
      tab[c] = make_float4(sc, fmaf(-mu, sc, a.gn_beta[c]), 0.f, 88.f);
    }
  } else if (a.table) {
    const float4* t = reinterpret_cast<const float4*>(a.table) + (int64_t)b * a.Cin;
    for (int c = threadIdx.x; c < a.Cin; c += blockDim.x) tab[c] = t[c];
  } else {
    for (int c = threadIdx.x; c < a.Cin; c += blockDim.x) tab[c] = make_float4(1.f, 0.f, 0.f, 88.f);
  }
}

struct EpiCoef { float es, eh, ps, fs, fh; };

__device__ __forceinline__ EpiCoef load_epi(const ConvArgs& a, int o) {
  EpiCoef e;
  e.es = a.e_scale ? a.e_scale[o] : 1.f;
  e.eh = a.e_shift ? a.e_shift[o] : 0.f;
  e.ps = a.post_scale ? a.post_scale[o] : 1.f;
  e.fs = a.f_scale ? a.f_scale[o] : 1.f;
  e.fh = a.f_shift ? a.f_shift[o] : 0.f;
  return e;
}

// y = act(acc*es + eh) * ps + res;  y = y*fs + fh
__device__ __forceinline__ float epilogue_value(float acc, const EpiCoef& e, int act, float res) {
  float y = apply_act(fmaf(acc, e.es, e.eh), act);
  y = fmaf(y, e.ps, res);
  return fmaf(y, e.fs, e.fh);
}

// block-reduce the per-thread side statistics, then ONE fp64 atomic pair per CTA into the CTA's slot
__device__ __forceinline__ void emit_side_stats(const ConvArgs& a, int b, float ssum, float ssq, float vmax, float vmin) {
  if (a.out_sample_sums) {
    __shared__ float red_stats[64];
    ssum = warp_sum(ssum);
    ssq = warp_sum(ssq);
    const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if ((threadIdx.x & 31) == 0) { red_stats[w] = ssum; red_stats[32 + w] = ssq; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0, s2 = 0.0;
      for (int i = 0; i < nw; ++i) { s += (double)red_stats[i]; s2 += (double)red_stats[32 + i]; }
      const int slot = (blockIdx.x + 7 * blockIdx.y) & (VRCOC_STAT_SLOTS - 1);
      double* dst = a.out_sample_sums + ((int64_t)b * VRCOC_STAT_SLOTS + slot) * 2;
      atomicAdd(dst, s);
      atomicAdd(dst + 1, s2);
    }
  }
  if (a.out_minmax) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
      vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
    }
    if ((threadIdx.x & 31) == 0) {
      atomicMax(&a.out_minmax[0], __float_as_uint(vmax));
      atomicMax(&a.out_minmax[1], ~__float_as_uint(vmin));
    }
  }
}

int launch_conv_simt(const ConvArgs& a, cudaStream_t st);
int launch_conv_small(const ConvArgs& a, cudaStream_t st);    // few-channel streaming kernel (conv_simt.cu)
bool conv_small_supported(const ConvArgs& a);
int launch_conv_tc(const ConvArgs& a, cudaStream_t st);       // tcgen05 bf16 path (conv_tc.cu)
bool conv_tc_supported(const ConvArgs& a);

}  // namespace vrcoc
