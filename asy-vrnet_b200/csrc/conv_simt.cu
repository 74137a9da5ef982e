// Convolution-as-GEMM engine, CUDA-core fp32 path (exact arithmetic for the fp32 parity gate: TF32 tensor-core
// products flip assignments and miss the 1e-4 gate).  See include/vrcoc.h for the operator contract.
//
// Tiling: one CTA = 128 output points (M) x 64 output channels (N) of one sample, K swept in slabs of 16 through
// double-buffered shared memory with register prefetch.  Points are the contiguous axis of NCHW, so A-slab rows
// (one logical input channel x 128 points) are 128-bit coalesced loads and each thread's 8x4 micro-tile writes
// 8 consecutive points per output channel.  The prologue (GroupNorm apply / attention gate / ECA scale) is applied
// while the slab is loaded; bias, activation, layer-scale, residual, BatchNorm affine and the side statistics are
// applied on the accumulators, so no intermediate ever reaches HBM.
#include "conv_common.cuh"

namespace vrcoc {

constexpr int BM = 128, BN = 64, BK = 16, BNP = 68;
constexpr int SIMT_THREADS = 256;

__device__ __forceinline__ void load8_any(const void* base, int64_t idx, int dtype, float (&v)[8]) {
  if (dtype == VRCOC_F32) ld8<float>(reinterpret_cast<const float*>(base) + idx, v);
  else ld8<__nv_bfloat16>(reinterpret_cast<const __nv_bfloat16*>(base) + idx, v);
}
__device__ __forceinline__ void store8_any(void* base, int64_t idx, int dtype, const float (&v)[8]) {
  if (dtype == VRCOC_F32) st8<float>(reinterpret_cast<float*>(base) + idx, v);
  else st8<__nv_bfloat16>(reinterpret_cast<__nv_bfloat16*>(base) + idx, v);
}

__global__ void __launch_bounds__(SIMT_THREADS) conv_simt_kernel(ConvArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  float* As = reinterpret_cast<float*>(smem);                 // [2][BK][BM]
  float* Bs = As + 2 * BK * BM;                               // [2][BK][BNP]
  float4* tab = reinterpret_cast<float4*>(Bs + 2 * BK * BNP); // [Cin]

  const int tid = threadIdx.x;
  const int b = blockIdx.z;
  const int p0 = blockIdx.x * BM;
  const int o0 = blockIdx.y * BN;
  const int P = a.P_out;
  const int taps = a.kh * a.kw;

  build_prologue_table(a, b, tab);
  __syncthreads();

  // ---- slab loaders -------------------------------------------------------------------------------------------
  const int a_k = tid >> 4;          // slab row (0..15)
  const int a_pg = tid & 15;         // point group: 8 consecutive points
  const int b_o = tid >> 2;          // weight row inside the tile (0..63)
  const int b_kq = (tid & 3) * 4;    // 4 consecutive k
  float areg[8], breg[4];

  auto load_a = [&](int k0) {
    const int kk = k0 + a_k;
#pragma unroll
    for (int i = 0; i < 8; ++i) areg[i] = 0.f;
    if (kk >= a.K) return;
    int c, tap;
    if (a.k_order) { tap = kk / a.Cin; c = kk - tap * a.Cin; }        // tap-major: k = tap*Cin + c
    else { c = kk / taps; tap = kk - c * taps; }                      // channel-major (PyTorch weight layout)
    const int s = a.chan_src ? a.chan_src[c] : c;
    const void* src; int dt; int64_t base;
    if (s < a.C0) { src = a.src0; dt = a.src0_dtype; base = (int64_t)b * a.src0_bstride + (int64_t)s * a.P_in; }
    else          { src = a.src1; dt = a.src1_dtype; base = (int64_t)b * a.src1_bstride + (int64_t)(s - a.C0) * a.P_in; }
    const float4 t = tab[c];
    const int q0 = p0 + a_pg * 8;
    if (a.fast1x1) {
      if (q0 < P) {
        load8_any(src, base + q0, dt, areg);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float x = areg[i];
          float y = fmaf(x, t.x, t.y);
          if (a.has_gate) y *= sigmoidf_exact(fmaf(t.z, x, t.w));
          areg[i] = y;
        }
      }
    } else {
      const int ky = tap / a.kw, kx = tap - ky * a.kw;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        int q = q0 + i;
        if (q < P) {
          int oy = q / a.W_out, ox = q - oy * a.W_out;
          int iy = oy * a.stride - a.pad + ky * a.dil, ix = ox * a.stride - a.pad + kx * a.dil;
          if (iy >= 0 && iy < a.H_in && ix >= 0 && ix < a.W_in) {
            float x = ld_any(src, base + (int64_t)iy * a.W_in + ix, dt);
            float y = fmaf(x, t.x, t.y);
            if (a.has_gate) y *= sigmoidf_exact(fmaf(t.z, x, t.w));
            areg[i] = y;
          }
        }
      }
    }
  };
  auto load_b = [&](int k0) {
    const int o = o0 + b_o;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int kk = k0 + b_kq + j;
      breg[j] = (o < a.O && kk < a.K) ? ld_any(a.weight, (int64_t)o * a.K + kk, a.weight_dtype) : 0.f;
    }
  };
  auto store_slab = [&](int buf) {
    float* ad = As + (buf * BK + a_k) * BM + a_pg * 8;
    *reinterpret_cast<float4*>(ad) = make_float4(areg[0], areg[1], areg[2], areg[3]);
    *reinterpret_cast<float4*>(ad + 4) = make_float4(areg[4], areg[5], areg[6], areg[7]);
#pragma unroll
    for (int j = 0; j < 4; ++j) Bs[(buf * BK + b_kq + j) * BNP + b_o] = breg[j];
  };

  // ---- main loop ------------------------------------------------------------------------------------------------
  const int pg = tid & 15;   // micro-tile: points pg*8..+8, outs og*4..+4
  const int og = tid >> 4;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int nk = (a.K + BK - 1) / BK;
  load_a(0); load_b(0); store_slab(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) { load_a((kt + 1) * BK); load_b((kt + 1) * BK); }
    const float* ap = As + buf * BK * BM + pg * 8;
    const float* bp = Bs + buf * BK * BNP + og * 4;
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(ap + k * BM);
      float4 a1 = *reinterpret_cast<const float4*>(ap + k * BM + 4);
      float4 bv = *reinterpret_cast<const float4*>(bp + k * BNP);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bw[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bw[j], acc[i][j]);
    }
    if (kt + 1 < nk) store_slab(buf ^ 1);
    __syncthreads();
  }

  // ---- epilogue ---------------------------------------------------------------------------------------------------
  const int q0 = p0 + pg * 8;
  float ssum = 0.f, ssq = 0.f, vmax = 0.f, vmin = __int_as_float(0x7f800000);
  if (q0 < P) {
    const bool full8 = a.vec_out && (q0 + 8 <= P);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int o = o0 + og * 4 + j;
      if (o >= a.O) continue;
      const EpiCoef ec = load_epi(a, o);
      void* dst; int ddt; int64_t dbase;
      if (o < a.O_split) { dst = a.out; ddt = a.out_dtype; dbase = ((int64_t)b * a.O_split + o) * P; }
      else { dst = a.out2; ddt = a.out2_dtype; dbase = ((int64_t)b * (a.O - a.O_split) + (o - a.O_split)) * P; }
      const int64_t rbase = ((int64_t)b * a.O + o) * P;
      float y[8];
      if (full8) {
        float r[8];
        if (a.res) load8_any(a.res, rbase + q0, a.res_dtype, r);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          y[i] = epilogue_value(acc[i][j], ec, a.act, a.res ? r[i] : 0.f);
          ssum += y[i]; ssq = fmaf(y[i], y[i], ssq);
          vmax = fmaxf(vmax, y[i]); vmin = fminf(vmin, y[i]);
        }
        store8_any(dst, dbase + q0, ddt, y);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (q0 + i < P) {
            float r = a.res ? ld_any(a.res, rbase + q0 + i, a.res_dtype) : 0.f;
            float v = epilogue_value(acc[i][j], ec, a.act, r);
            ssum += v; ssq = fmaf(v, v, ssq);
            vmax = fmaxf(vmax, v); vmin = fminf(vmin, v);
            st_any(dst, dbase + q0 + i, ddt, v);
          }
        }
      }
    }
  }
  emit_side_stats(a, b, ssum, ssq, vmax, vmin);
}

// ---- few-channel convolutions (the 512x512 ingest: 1x1 3->3 / 4->4 / 7->4 and the 3x3 4->3 radar projection) ---------------
// O <= 8 outputs and K <= 64 taps: bandwidth-bound streaming work, one thread = one output pixel x all output channels;
// the weights and the prologue table live in shared memory, the K gathered inputs of neighbouring pixels share L1 lines.
// (On the GEMM tiles these layers waste >90 % of every 128x32x64 MMA slab and 16 K CTAs of fixed setup cost.)
constexpr int SMALL_MAX_O = 8, SMALL_MAX_K = 64, SMALL_THREADS = 256;

__global__ void __launch_bounds__(SMALL_THREADS) conv_small_kernel(ConvArgs a) {
  __shared__ float wsm[SMALL_MAX_O * SMALL_MAX_K];
  __shared__ float4 tab[SMALL_MAX_K];
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < a.O * a.K; i += blockDim.x) wsm[i] = ld_any(a.weight, i, a.weight_dtype);
  build_prologue_table(a, b, tab);
  __syncthreads();
  const int taps = a.kh * a.kw;
  const int P = a.P_out;
  float ssum = 0.f, ssq = 0.f, vmax = 0.f, vmin = __int_as_float(0x7f800000);
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < P; q += gridDim.x * blockDim.x) {
    const int oy = q / a.W_out, ox = q - oy * a.W_out;
    float acc[SMALL_MAX_O];
#pragma unroll
    for (int o = 0; o < SMALL_MAX_O; ++o) acc[o] = 0.f;
    int kk = 0;
    for (int c = 0; c < a.Cin; ++c) {
      const int s = a.chan_src ? a.chan_src[c] : c;
      const void* src; int dt; int64_t base;
      if (s < a.C0) { src = a.src0; dt = a.src0_dtype; base = (int64_t)b * a.src0_bstride + (int64_t)s * a.P_in; }
      else          { src = a.src1; dt = a.src1_dtype; base = (int64_t)b * a.src1_bstride + (int64_t)(s - a.C0) * a.P_in; }
      const float4 t = tab[c];
      for (int ky = 0; ky < a.kh; ++ky) {
        const int iy = oy * a.stride - a.pad + ky * a.dil;
        for (int kx = 0; kx < a.kw; ++kx, ++kk) {
          const int ix = ox * a.stride - a.pad + kx * a.dil;
          float z = 0.f;
          if (iy >= 0 && iy < a.H_in && ix >= 0 && ix < a.W_in) {
            const float x = ld_any(src, base + (int64_t)iy * a.W_in + ix, dt);
            z = fmaf(x, t.x, t.y);
            if (a.has_gate) z *= sigmoidf_exact(fmaf(t.z, x, t.w));
          }
#pragma unroll
          for (int o = 0; o < SMALL_MAX_O; ++o)
            if (o < a.O) acc[o] = fmaf(wsm[o * a.K + kk], z, acc[o]);
        }
      }
    }
#pragma unroll
    for (int o = 0; o < SMALL_MAX_O; ++o) {
      if (o < a.O) {
        const EpiCoef ec = load_epi(a, o);
        const float r = a.res ? ld_any(a.res, ((int64_t)b * a.O + o) * P + q, a.res_dtype) : 0.f;
        const float y = epilogue_value(acc[o], ec, a.act, r);
        ssum += y; ssq = fmaf(y, y, ssq);
        vmax = fmaxf(vmax, y); vmin = fminf(vmin, y);
        if (o < a.O_split) st_any(a.out, ((int64_t)b * a.O_split + o) * P + q, a.out_dtype, y);
        else st_any(a.out2, ((int64_t)b * (a.O - a.O_split) + (o - a.O_split)) * P + q, a.out2_dtype, y);
      }
    }
  }
  emit_side_stats(a, b, ssum, ssq, vmax, vmin);
}

// Vectorised variant (stride 1, W_out % 8 == 0, one source dtype): one thread = 8 consecutive output pixels of a row x all
// output channels; per (channel, ky) it loads the 8 + kw - 1 input window once (one 128-bit load + edge scalars) and reuses it
// for the kw taps; outputs leave as 128-bit stores.
// "Same" k x k convolution (k = KW in {1, 3}, stride 1, pad (k-1)/2) on rows that are a multiple of 8 wide and 16-byte
// aligned — every ingest convolution of the model (512 x 512 frames, 3..7 channels).  Compile-time taps: the 8 centre pixels
// of a row are ONE 128-bit load, the halo two scalars; no per-tap predicates.  Bandwidth-bound by construction.
template <typename TS, int MAXO, int KW>
__global__ void __launch_bounds__(SMALL_THREADS) conv_small_same_kernel(ConvArgs a) {
  constexpr int PAD = (KW - 1) / 2;
  __shared__ float wsm[SMALL_MAX_O * SMALL_MAX_K];
  __shared__ float4 tab[SMALL_MAX_K];
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < a.O * a.K; i += blockDim.x) wsm[i] = ld_any(a.weight, i, a.weight_dtype);
  build_prologue_table(a, b, tab);
  __syncthreads();
  const int P = a.P_out, W = a.W_in;
  const int groups = P >> 3;
  float ssum = 0.f, ssq = 0.f, vmax = 0.f, vmin = __int_as_float(0x7f800000);
  for (int gi = blockIdx.x * blockDim.x + threadIdx.x; gi < groups; gi += gridDim.x * blockDim.x) {
    const int q0 = gi << 3;
    const int oy = q0 / W, ox0 = q0 - oy * W;
    float acc[MAXO][8];
#pragma unroll
    for (int o = 0; o < MAXO; ++o)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[o][j] = 0.f;
    for (int c = 0; c < a.Cin; ++c) {
      const int s = a.chan_src ? a.chan_src[c] : c;
      const TS* plane = (s < a.C0) ? reinterpret_cast<const TS*>(a.src0) + (int64_t)b * a.src0_bstride + (int64_t)s * a.P_in
                                   : reinterpret_cast<const TS*>(a.src1) + (int64_t)b * a.src1_bstride + (int64_t)(s - a.C0) * a.P_in;
      const float4 t = tab[c];
#pragma unroll
      for (int ky = 0; ky < KW; ++ky) {
        const int iy = oy - PAD + ky;
        if (iy < 0 || iy >= a.H_in) continue;
        const TS* row = plane + (int64_t)iy * W;
        float win[8 + 2 * PAD];
        {
          float ctr[8];
          ld8<TS>(row + ox0, ctr);
#pragma unroll
          for (int j = 0; j < 8; ++j) win[PAD + j] = ctr[j];
          if (PAD) {
            win[0] = ox0 > 0 ? (float)row[ox0 - 1] : 0.f;
            win[8 + 2 * PAD - 1] = ox0 + 8 < W ? (float)row[ox0 + 8] : 0.f;
          }
        }
#pragma unroll
        for (int j = 0; j < 8 + 2 * PAD; ++j) {
          const float x = win[j];
          float z = fmaf(x, t.x, t.y);
          if (a.has_gate) z *= sigmoidf_exact(fmaf(t.z, x, t.w));
          win[j] = z;
        }
        if (PAD) {                                            // zero padding is applied AFTER the prologue
          if (ox0 == 0) win[0] = 0.f;
          if (ox0 + 8 >= W) win[8 + 2 * PAD - 1] = 0.f;
        }
        const float* wrow = wsm + c * KW * KW + ky * KW;
#pragma unroll
        for (int kx = 0; kx < KW; ++kx) {
#pragma unroll
          for (int o = 0; o < MAXO; ++o) {
            if (o < a.O) {
              const float wv = wrow[o * a.K + kx];
#pragma unroll
              for (int j = 0; j < 8; ++j) acc[o][j] = fmaf(wv, win[j + kx], acc[o][j]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int o = 0; o < MAXO; ++o) {
      if (o < a.O) {
        const EpiCoef ec = load_epi(a, o);
        float r[8], y[8];
        if (a.res) load8_any(a.res, ((int64_t)b * a.O + o) * P + q0, a.res_dtype, r);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          y[j] = epilogue_value(acc[o][j], ec, a.act, a.res ? r[j] : 0.f);
          ssum += y[j]; ssq = fmaf(y[j], y[j], ssq);
          vmax = fmaxf(vmax, y[j]); vmin = fminf(vmin, y[j]);
        }
        if (o < a.O_split) store8_any(a.out, ((int64_t)b * a.O_split + o) * P + q0, a.out_dtype, y);
        else store8_any(a.out2, ((int64_t)b * (a.O - a.O_split) + (o - a.O_split)) * P + q0, a.out2_dtype, y);
      }
    }
  }
  emit_side_stats(a, b, ssum, ssq, vmax, vmin);
}

// The four ingest convolutions of the model at compile-time sizes (bf16 sources, no gate): same contract as
// conv_small_same_kernel, but a thread first issues EVERY row load it needs (clamped addresses, no branches: CIN x KW 128-bit
// loads + halos in flight at once) and only then computes.  The run-time kernel above keeps one (channel, row) round trip to
// HBM in flight per thread and ran at ~10 % of the HBM roofline on the 512 x 512 frames (55 us for 29 MB).
// ACT (none / ReLU), STATS (side statistics wanted) and bf16 outputs are compile-time too: ncu showed the first version
// issue-bound at 835 warp instructions for 8 pixels of a 3 -> 3 channel 1x1 (run-time activation switch, statistics and
// dtype dispatch per output element).
template <int KW, int CIN, int OC, int ACT, bool STATS>
__global__ void __launch_bounds__(SMALL_THREADS) conv_ingest_kernel(ConvArgs a) {
  constexpr int PAD = (KW - 1) / 2, K = CIN * KW * KW;
  __shared__ float wsm[OC * K];
  __shared__ float4 tab[CIN];
  __shared__ float epi[5 * OC];
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < OC * K; i += blockDim.x) wsm[i] = ld_any(a.weight, i, a.weight_dtype);
  if (threadIdx.x < OC) {
    const EpiCoef ec = load_epi(a, threadIdx.x);
    epi[threadIdx.x] = ec.es; epi[OC + threadIdx.x] = ec.eh; epi[2 * OC + threadIdx.x] = ec.ps;
    epi[3 * OC + threadIdx.x] = ec.fs; epi[4 * OC + threadIdx.x] = ec.fh;
  }
  build_prologue_table(a, b, tab);
  __syncthreads();
  const int P = a.P_out, W = a.W_in, H = a.H_in;
  const int groups = P >> 3;
  const __nv_bfloat16* planes[CIN];
#pragma unroll
  for (int c = 0; c < CIN; ++c) {
    const int s = a.chan_src ? a.chan_src[c] : c;
    planes[c] = (s < a.C0) ? reinterpret_cast<const __nv_bfloat16*>(a.src0) + (int64_t)b * a.src0_bstride + (int64_t)s * a.P_in
                           : reinterpret_cast<const __nv_bfloat16*>(a.src1) + (int64_t)b * a.src1_bstride + (int64_t)(s - a.C0) * a.P_in;
  }
  float ssum = 0.f, ssq = 0.f, vmax = 0.f, vmin = __int_as_float(0x7f800000);
  for (int gi = blockIdx.x * blockDim.x + threadIdx.x; gi < groups; gi += gridDim.x * blockDim.x) {
    const int q0 = gi << 3;
    const int oy = q0 / W, ox0 = q0 - oy * W;
    // ---- phase 1: every load, unconditionally (rows / halo columns outside the frame are clamped and zeroed later) ------------
    uint4 ctr[CIN][KW];
    unsigned short hl[CIN][KW], hr[CIN][KW];
    const int xl = ox0 > 0 ? ox0 - 1 : 0, xr = ox0 + 8 < W ? ox0 + 8 : W - 1;
#pragma unroll
    for (int c = 0; c < CIN; ++c) {
#pragma unroll
      for (int ky = 0; ky < KW; ++ky) {
        int iy = oy - PAD + ky;
        iy = iy < 0 ? 0 : (iy >= H ? H - 1 : iy);
        const __nv_bfloat16* row = planes[c] + (int64_t)iy * W;
        ctr[c][ky] = __ldg(reinterpret_cast<const uint4*>(row + ox0));
        if (PAD) {
          hl[c][ky] = __ldg(reinterpret_cast<const unsigned short*>(row + xl));
          hr[c][ky] = __ldg(reinterpret_cast<const unsigned short*>(row + xr));
        }
      }
    }
    // ---- phase 2: prologue, zero padding (applied AFTER the prologue), taps -------------------------------------------------------
    float acc[OC][8];
#pragma unroll
    for (int o = 0; o < OC; ++o)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[o][j] = 0.f;
#pragma unroll
    for (int c = 0; c < CIN; ++c) {
      const float4 t = tab[c];
#pragma unroll
      for (int ky = 0; ky < KW; ++ky) {
        const int iy = oy - PAD + ky;
        const bool row_ok = iy >= 0 && iy < H;
        float win[8 + 2 * PAD];
        const uint32_t wds[4] = {ctr[c][ky].x, ctr[c][ky].y, ctr[c][ky].z, ctr[c][ky].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          win[PAD + 2 * i] = __uint_as_float(wds[i] << 16);
          win[PAD + 2 * i + 1] = __uint_as_float(wds[i] & 0xffff0000u);
        }
        if (PAD) {
          win[0] = __uint_as_float((uint32_t)hl[c][ky] << 16);
          win[8 + 2 * PAD - 1] = __uint_as_float((uint32_t)hr[c][ky] << 16);
        }
#pragma unroll
        for (int j = 0; j < 8 + 2 * PAD; ++j) win[j] = row_ok ? fmaf(win[j], t.x, t.y) : 0.f;
        if (PAD) {
          if (ox0 == 0) win[0] = 0.f;
          if (ox0 + 8 >= W) win[8 + 2 * PAD - 1] = 0.f;
        }
        const float* wrow = wsm + c * KW * KW + ky * KW;
#pragma unroll
        for (int kx = 0; kx < KW; ++kx) {
#pragma unroll
          for (int o = 0; o < OC; ++o) {
            const float wv = wrow[o * K + kx];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[o][j] = fmaf(wv, win[j + kx], acc[o][j]);
          }
        }
      }
    }
    __nv_bfloat16* outp = reinterpret_cast<__nv_bfloat16*>(a.out) + (int64_t)b * OC * P + q0;
#pragma unroll
    for (int o = 0; o < OC; ++o) {
      const float es = epi[o], eh = epi[OC + o], ps = epi[2 * OC + o], fs = epi[3 * OC + o], fh = epi[4 * OC + o];
      float r[8], y[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = 0.f;
      if (a.res) ld8<__nv_bfloat16>(reinterpret_cast<const __nv_bfloat16*>(a.res) + ((int64_t)b * OC + o) * P + q0, r);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float v = fmaf(acc[o][j], es, eh);                       // same order as epilogue_value (conv_common.cuh)
        if (ACT == VRCOC_ACT_RELU) v = fmaxf(v, 0.f);
        v = fmaf(v, ps, r[j]);
        v = fmaf(v, fs, fh);
        y[j] = v;
        if (STATS) {
          ssum += v; ssq = fmaf(v, v, ssq);
          vmax = fmaxf(vmax, v); vmin = fminf(vmin, v);
        }
      }
      st8<__nv_bfloat16>(outp + (int64_t)o * P, y);
    }
  }
  if (STATS) emit_side_stats(a, b, ssum, ssq, vmax, vmin);
}

template <typename TS, int MAXO>
__global__ void __launch_bounds__(SMALL_THREADS) conv_small_vec_kernel(ConvArgs a) {
  __shared__ float wsm[SMALL_MAX_O * SMALL_MAX_K];
  __shared__ float4 tab[SMALL_MAX_K];
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < a.O * a.K; i += blockDim.x) wsm[i] = ld_any(a.weight, i, a.weight_dtype);
  build_prologue_table(a, b, tab);
  __syncthreads();
  const int P = a.P_out;
  const int groups = P >> 3;
  const int taps = a.kh * a.kw;
  float ssum = 0.f, ssq = 0.f, vmax = 0.f, vmin = __int_as_float(0x7f800000);
  for (int gi = blockIdx.x * blockDim.x + threadIdx.x; gi < groups; gi += gridDim.x * blockDim.x) {
    const int q0 = gi << 3;
    const int oy = q0 / a.W_out, ox0 = q0 - oy * a.W_out;
    float acc[MAXO][8];
#pragma unroll
    for (int o = 0; o < MAXO; ++o)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[o][j] = 0.f;
    for (int c = 0; c < a.Cin; ++c) {
      const int s = a.chan_src ? a.chan_src[c] : c;
      const TS* plane = (s < a.C0) ? reinterpret_cast<const TS*>(a.src0) + (int64_t)b * a.src0_bstride + (int64_t)s * a.P_in
                                   : reinterpret_cast<const TS*>(a.src1) + (int64_t)b * a.src1_bstride + (int64_t)(s - a.C0) * a.P_in;
      const float4 t = tab[c];
      for (int ky = 0; ky < a.kh; ++ky) {
        const int iy = oy - a.pad + ky;
        if (iy < 0 || iy >= a.H_in) continue;
        const TS* row = plane + (int64_t)iy * a.W_in;
        const int ix0 = ox0 - a.pad;
        float win[8 + 6];                                    // kw <= 7
#pragma unroll
        for (int j = 0; j < 8 + 6; ++j) {
          win[j] = 0.f;
          if (j < 8 + a.kw - 1) {
            const int ix = ix0 + j;
            if (ix >= 0 && ix < a.W_in) {
              const float x = (float)row[ix];
              float z = fmaf(x, t.x, t.y);
              if (a.has_gate) z *= sigmoidf_exact(fmaf(t.z, x, t.w));
              win[j] = z;
            }
          }
        }
        const float* wrow = wsm + c * taps + ky * a.kw;
#pragma unroll
        for (int kx = 0; kx < 7; ++kx) {
          if (kx < a.kw) {
#pragma unroll
            for (int o = 0; o < MAXO; ++o) {
              if (o < a.O) {
                const float wv = wrow[o * a.K + kx];
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[o][j] = fmaf(wv, win[j + kx], acc[o][j]);
              }
            }
          }
        }
      }
    }
#pragma unroll
    for (int o = 0; o < MAXO; ++o) {
      if (o < a.O) {
        const EpiCoef ec = load_epi(a, o);
        float r[8], y[8];
        if (a.res) load8_any(a.res, ((int64_t)b * a.O + o) * P + q0, a.res_dtype, r);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          y[j] = epilogue_value(acc[o][j], ec, a.act, a.res ? r[j] : 0.f);
          ssum += y[j]; ssq = fmaf(y[j], y[j], ssq);
          vmax = fmaxf(vmax, y[j]); vmin = fminf(vmin, y[j]);
        }
        if (o < a.O_split) store8_any(a.out, ((int64_t)b * a.O_split + o) * P + q0, a.out_dtype, y);
        else store8_any(a.out2, ((int64_t)b * (a.O - a.O_split) + (o - a.O_split)) * P + q0, a.out2_dtype, y);
      }
    }
  }
  emit_side_stats(a, b, ssum, ssq, vmax, vmin);
}

bool conv_small_supported(const ConvArgs& a) { return a.O <= SMALL_MAX_O && a.K <= SMALL_MAX_K && a.k_order == 0; }

int launch_conv_small(const ConvArgs& a, cudaStream_t st) {
  const bool same_dt = a.C1 == 0 || a.src1_dtype == a.src0_dtype;
  const bool vec = a.stride == 1 && a.dil == 1 && a.W_out % 8 == 0 && a.kw <= 7 && same_dt && a.vec_out;
  if (vec) {
    int bx = (int)cdiv(a.P_out / 8, SMALL_THREADS);
    if (bx > 2048) bx = 2048;
    dim3 grid(bx, a.B);
    const bool f32 = a.src0_dtype == VRCOC_F32;
    const int es = f32 ? 4 : 2;
    auto al = [&](const void* p, int64_t bstride) {
      return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && (bstride * es) % 16 == 0 && ((int64_t)a.P_in * es) % 16 == 0;
    };
    const bool same = a.kh == a.kw && (a.kw == 1 || a.kw == 3) && a.pad == (a.kw - 1) / 2 && a.W_in == a.W_out && a.H_in == a.H_out &&
                      a.W_in % 8 == 0 && al(a.src0, a.src0_bstride) && (a.C1 == 0 || al(a.src1, a.src1_bstride));
    if (same && !f32 && !a.has_gate && (a.C1 == 0 || a.src1_dtype == VRCOC_BF16) && a.out_dtype == VRCOC_BF16 && a.O_split == a.O &&
        (a.act == VRCOC_ACT_NONE || a.act == VRCOC_ACT_RELU) && (!a.res || a.res_dtype == VRCOC_BF16) && a.vec_out) {
      // the ingest convolutions of the model: compile-time channel counts, all loads of a thread in flight at once
      const bool relu = a.act == VRCOC_ACT_RELU, stats = a.out_sample_sums || a.out_minmax;
#define INGEST(KWV, CI, OCV)                                                                                           \
  do {                                                                                                                 \
    if (relu && stats) conv_ingest_kernel<KWV, CI, OCV, VRCOC_ACT_RELU, true><<<grid, SMALL_THREADS, 0, st>>>(a);      \
    else if (relu) conv_ingest_kernel<KWV, CI, OCV, VRCOC_ACT_RELU, false><<<grid, SMALL_THREADS, 0, st>>>(a);         \
    else if (stats) conv_ingest_kernel<KWV, CI, OCV, VRCOC_ACT_NONE, true><<<grid, SMALL_THREADS, 0, st>>>(a);         \
    else conv_ingest_kernel<KWV, CI, OCV, VRCOC_ACT_NONE, false><<<grid, SMALL_THREADS, 0, st>>>(a);                   \
    return check_launch("conv_ingest");                                                                                \
  } while (0)
      if (a.kw == 3 && a.Cin == 4 && a.O == 3) INGEST(3, 4, 3);
      if (a.kw == 1 && a.Cin == 7 && a.O == 4) INGEST(1, 7, 4);
      if (a.kw == 1 && a.Cin == 3 && a.O == 3) INGEST(1, 3, 3);
      if (a.kw == 1 && a.Cin == 4 && a.O == 4) INGEST(1, 4, 4);
#undef INGEST
    }
    if (same) {
#define SAME(TS, MO, KWV) conv_small_same_kernel<TS, MO, KWV><<<grid, SMALL_THREADS, 0, st>>>(a)
      if (a.O <= 4) {
        if (a.kw == 1) { if (f32) SAME(float, 4, 1); else SAME(__nv_bfloat16, 4, 1); }
        else           { if (f32) SAME(float, 4, 3); else SAME(__nv_bfloat16, 4, 3); }
      } else {
        if (a.kw == 1) { if (f32) SAME(float, 8, 1); else SAME(__nv_bfloat16, 8, 1); }
        else           { if (f32) SAME(float, 8, 3); else SAME(__nv_bfloat16, 8, 3); }
      }
#undef SAME
      return check_launch("conv_small_same");
    }
    if (a.O <= 4) {
      if (f32) conv_small_vec_kernel<float, 4><<<grid, SMALL_THREADS, 0, st>>>(a);
      else conv_small_vec_kernel<__nv_bfloat16, 4><<<grid, SMALL_THREADS, 0, st>>>(a);
    } else {
      if (f32) conv_small_vec_kernel<float, 8><<<grid, SMALL_THREADS, 0, st>>>(a);
      else conv_small_vec_kernel<__nv_bfloat16, 8><<<grid, SMALL_THREADS, 0, st>>>(a);
    }
    return check_launch("conv_small_vec");
  }
  int bx = (int)cdiv(a.P_out, SMALL_THREADS);
  if (bx > 1024) bx = 1024;
  conv_small_kernel<<<dim3(bx, a.B), SMALL_THREADS, 0, st>>>(a);
  return check_launch("conv_small");
}

int launch_conv_simt(const ConvArgs& a, cudaStream_t st) {
  size_t smem = (size_t)(2 * BK * BM + 2 * BK * BNP) * sizeof(float) + (size_t)a.Cin * sizeof(float4);
  VRCOC_REQUIRE(smem <= 200 * 1024, "conv: too many input channels (%d) for the prologue table", a.Cin);
  static thread_local size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaFuncSetAttribute(conv_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured = smem;
  }
  dim3 grid((unsigned)cdiv(a.P_out, BM), (unsigned)cdiv(a.O, BN), (unsigned)a.B);
  conv_simt_kernel<<<grid, SIMT_THREADS, smem, st>>>(a);
  return check_launch("conv_simt");
}

}  // namespace vrcoc
