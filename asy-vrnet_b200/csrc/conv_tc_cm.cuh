// tcgen05 path, kernel 4: 1x1 projection with a CHANNEL-MAJOR accumulator (included by conv_tc.cu).
//
//   D[128 outs x 128 points] (TMEM, fp32)  +=  W[128 outs x 64 k] (smem, K-major)  *  X[64 k x 128 points] (smem, MN-major)
//
// The other tcgen05 kernels put the points on the TMEM lanes, so an epilogue thread owns ONE point and 16 output channels:
// in NCHW that is sixteen 2-byte stores, sixteen shared-memory coefficient reads and ~16 SASS instructions per element
// (profiles/: the epilogue is 2/3 of a CTA's life, issue-bound).  Here the roles of the operands are swapped — the weights
// are the A operand (M = output channels), the activations the B operand (N = points) — so an epilogue thread owns ONE
// output channel and 16 CONSECUTIVE points: its bias / BN / layer-scale are five registers, the residual is two 16-byte
// loads, the result two 16-byte stores, ~2 instructions per element before the activation.  The shared-memory images of both
// operands are the same as in the other kernels (weights: K-major SW128 by TMA; activations: 64-point x 64-channel SW128
// blocks by TMA or by the transform-on-load builders), only the descriptors trade places.
//
// One CTA owns one 128-point tile and a range of 128-channel output tiles:
//   warps 0-7  (XMODE 2: build the normalised / gated bf16 X operand once, all K slabs resident, then) drain the accumulator;
//   warp 8     lane 0 runs the TMA ring (weights, and X when it comes by TMA);
//   warp 9     lane 0 issues tcgen05.mma into TMEM buffer j&1 while the epilogue warps drain buffer (j-1)&1.
// XMODE 0: X streamed with the weights (K too large to keep);  1: X resident, fetched by TMA with the first output tile;
//       2: X resident, built by the CTA's threads through the prologue table (GroupNorm / gate / plain fp32 sources);
//       3: X resident, bf16 source with a prologue: the RAW slabs arrive by TMA already in the operand layout (all of them in
//          flight at once) and warps 0-7 normalise / gate them IN PLACE — the prologue is per channel = per k-row, so the
//          layout does not change; slab by slab, so the first MMAs start while later slabs are still being transformed.
#pragma once

namespace vrcoc {

constexpr int TQ_THREADS = 320;
constexpr int TQ_NP = 128;                          // points per tile  (MMA N)
constexpr int TQ_MT = 128;                          // outputs per tile (MMA M)
constexpr int TQ_W_BYTES = TQ_MT * TC_BK * 2;       // 16 KB
constexpr int TQ_X_BYTES = TQ_NP * TC_BK * 2;       // 16 KB
constexpr int TQ_MAX_STAGES = 8;
constexpr int TQ_MAX_SLABS3 = 9;                    // XMODE 3: per-slab barriers

struct TqLayout {
  int stages;        // ring depth
  int nslabs;        // K slabs
  int nx;            // X slabs (== nslabs, except with the folded GroupNorm: the weights are [hi | lo], 2 * nx slabs per feat tile)
  int tiles;         // output tiles per CTA
  int plain;         // epilogue is act(acc*es + eh) only
  int stage_bufs;    // 1 or 2 epilogue staging buffers of 8 x 4 KB
  int off_x, off_w, off_stage, off_tab, off_bar, total;
};

// instruction descriptor: D=f32, A=B=bf16, A K-major, B MN-major, N = points, M = 128
__host__ __device__ constexpr uint32_t make_idesc_cm(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (0u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TQ_MT >> 4) << 24);
}

struct CmCoef { float es, eh, ps, fs, fh; };

// split tcgen05.ld: the destination registers are only defined after tmem_ld_wait, which takes them as in/out operands so
// that the compiler orders every use after the wait
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,"
      "%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                 "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}

// ---- packed fp32x2 arithmetic (FFMA2: two lanes' worth of FMA per issue slot; ncu showed the GELU epilogue bound by the
// FMA pipe, not by issue) ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t pk2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// gelu_fast (conv_tc.cu) on two values
__device__ __forceinline__ uint64_t gelu2(uint64_t x) {
  float x0, x1;
  upk2(x, x0, x1);
  const float a0 = fminf(fabsf(x0), VRCOC_GELU_AMAX), a1 = fminf(fabsf(x1), VRCOC_GELU_AMAX);
  const uint64_t a = pk2(a0, a1);
  uint64_t p = fma2(pk2(VRCOC_GELU_C3, VRCOC_GELU_C3), a, pk2(VRCOC_GELU_C2, VRCOC_GELU_C2));
  p = fma2(p, a, pk2(VRCOC_GELU_C1, VRCOC_GELU_C1));
  p = fma2(p, a, pk2(VRCOC_GELU_C0, VRCOC_GELU_C0));
  float p0, p1;
  upk2(p, p0, p1);
  return fma2(pk2(-a0, -a1), pk2(ex2_approx(p0), ex2_approx(p1)), pk2(fmaxf(x0, 0.f), fmaxf(x1, 0.f)));
}

template <int ACT>
__device__ __forceinline__ uint64_t act2(uint64_t v) {
  if (ACT == VRCOC_ACT_NONE) return v;
  if (ACT == VRCOC_ACT_GELU) return gelu2(v);
  float a, b;
  upk2(v, a, b);
  return pk2(act_tc<ACT>(a), act_tc<ACT>(b));
}

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ void unpack8_bf16(const uint4& u, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}

// Drains the accumulator tiles of one CTA.  Warp w owns TMEM lanes 32*(w%4).. (= 32 output channels) and the point half
// w/4 (64 points) of every tile.  Its 4 KB staging region is one TMA box [32 channels][128 B], 128-byte swizzled: the bf16
// residual lands there by TMA, is replaced in place by the result, and leaves by TMA store — global memory only ever sees
// whole 128-byte rows, and the clipping of the boxes to the tensors replaces every edge predicate on the stores.
// fp32 outputs (the similarity half of fc1|fc_v) go through the same region in two 32-point passes.
// RES_MODE: 0 = the residual box is requested on entry; 1 = after the accumulator wait (region busy until then);
//           2 = the caller has already requested it on `rbar` (first phase);
//           3 = LINEAR only: the boxes of all tiles are requested on entry (one barrier phase; every tile has its own region)
// LINEAR: the `tiles` accumulators are all complete when acc_full[0] fires and sit side by side in TMEM (tile j at column
//         j * 128); tile j is staged in region0 + j * buf_stride (the fused MLP's output epilogue)
template <int ACT, bool PLAIN, int RES_MODE = 0, bool LINEAR = false>
__device__ __forceinline__ void cm_epilogue(const ConvArgs& a, uint32_t tmem_base, uint64_t* acc_full, uint64_t* acc_empty, uint64_t* rbar,
                                            unsigned char* region0, int buf_stride, const CUtensorMap* tmO1, const CUtensorMap* tmO2,
                                            const CUtensorMap* tmR, int b, int p0, int o_begin, int tiles) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int lq = warp & 3, ch = warp >> 2;
  const int P = a.P_out;
  const int q_base = p0 + ch * (TQ_NP / 2);
  const bool has_res = !PLAIN && a.res != nullptr;
  const bool want_sums = !PLAIN && a.out_sample_sums != nullptr;
  const bool want_mm = !PLAIN && a.out_minmax != nullptr;
  const int sw = lane & 7;
  uint32_t res_phase = 0;
  unsigned long long t_wait = 0, t_comp = 0;       // debug trace: warp 0's time waiting for accumulators / draining them
  float ssum = 0.f, ssq = 0.f, vmax = 0.f, vmin = __int_as_float(0x7f800000);
  // folded GroupNorm (vrcoc.h, gn_fold_k1): y = rstd_b * acc + (k0[o] - rstd_b * mean_b * k1[o])
  float fold_mu = 0.f, fold_rstd = 1.f;
  if (a.gn_fold_k1) gn_mean_rstd(a.gn_sums, b, (double)a.C0 * (double)a.P_in, a.gn_eps, fold_mu, fold_rstd);
  if (RES_MODE == 3 && has_res && lane == 0) {
    int live = 0;
    for (int j = 0; j < tiles; ++j) live += (o_begin + j * TQ_MT + lq * 32 < a.O) ? 1 : 0;
    if (live) mbar_expect_tx(rbar, 4096u * (uint32_t)live);
    for (int j = 0; j < live; ++j) tma_load_3d(region0 + j * buf_stride, tmR, q_base, o_begin + j * TQ_MT + lq * 32, b, rbar);
  }
  for (int j = 0; j < tiles; ++j) {
    const int buf = LINEAR ? 0 : (j & 1);
    const int o_row0 = o_begin + j * TQ_MT + lq * 32;
    const int o = o_row0 + lane;
    const bool ok = o < a.O;
    const bool row_live = o_row0 < a.O;                              // warp-uniform
    // staging buffer j&1 (when there are two): its previous store — tile j-2's — must have finished reading it
    unsigned char* region = region0 + (LINEAR ? j : (j & 1)) * buf_stride;
    const uint32_t rbase = smem_u32(region) + lane * 128;
    if (j > 0 && !LINEAR) {
      if (lane == 0) {
        if (buf_stride) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        else tma_store_wait_read();
      }
      __syncwarp();
    }
    if (RES_MODE == 0 && has_res && row_live && lane == 0) {
      mbar_expect_tx(rbar, 4096u);
      tma_load_3d(region, tmR, q_base, o_row0, b, rbar);
    }
    CmCoef k;
    k.es = (ok && a.e_scale) ? __ldg(a.e_scale + o) : 1.f;
    k.eh = (ok && a.e_shift) ? __ldg(a.e_shift + o) : 0.f;
    if (a.gn_fold_k1) {
      k.es = fold_rstd;
      k.eh = ok ? fmaf(-fold_rstd * fold_mu, __ldg(a.gn_fold_k1 + o), k.eh) : 0.f;
    }
    k.ps = (!PLAIN && ok && a.post_scale) ? __ldg(a.post_scale + o) : 1.f;
    k.fs = (!PLAIN && ok && a.f_scale) ? __ldg(a.f_scale + o) : 1.f;
    k.fh = (!PLAIN && ok && a.f_shift) ? __ldg(a.f_shift + o) : 0.f;
    const bool second = o_row0 >= a.O_split;                         // warp-uniform: O_split % 32 == 0 (host-checked)
    const int odt = second ? a.out2_dtype : a.out_dtype;
    const CUtensorMap* tmO = second ? tmO2 : tmO1;
    const int ochan0 = second ? o_row0 - a.O_split : o_row0;
    const bool tr = g_tc_trace != nullptr && threadIdx.x == 0;
    unsigned long long t0 = 0, t1 = 0;
    if (tr) t0 = gtime();
    mbar_wait(&acc_full[buf], LINEAR ? 0u : ((uint32_t)(j >> 1) & 1));
    tc_fence_after();
    if (threadIdx.x == 0 && j == 0) trace(3);
    if (tr) { t1 = gtime(); t_wait += t1 - t0; }
    if (RES_MODE == 1 && has_res && row_live && lane == 0) {         // the region only becomes free with the accumulator
      mbar_expect_tx(rbar, 4096u);
      tma_load_3d(region, tmR, q_base, o_row0, b, rbar);
    }
    if (has_res && row_live) {
      if (RES_MODE == 3) mbar_wait(rbar, 0);
      else { mbar_wait(rbar, res_phase); res_phase ^= 1u; }
    }
    const uint32_t tbase = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)((LINEAR ? j : buf) * TQ_NP + ch * (TQ_NP / 2));
    // the TMEM read of group c+1 is in flight while group c is processed
    // (not unrolled: a CTA often drains a single tile, i.e. runs this code once, from a cold instruction cache — ncu showed
    // instruction fetch, not issue, bounding the fully unrolled version)
    uint32_t rn[16];
    tmem_ld16_issue(tbase, rn);
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t r[16];
      tmem_ld_wait(rn);
#pragma unroll
      for (int i = 0; i < 16; ++i) r[i] = rn[i];
      if (c < 3) tmem_ld16_issue(tbase + (uint32_t)(16 * (c + 1)), rn);
      if (row_live) {
        const int q0 = q_base + 16 * c;
        float y[16];
        {
          const uint64_t es2 = pk2(k.es, k.es), eh2 = pk2(k.eh, k.eh);
          uint64_t v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i)
            v[i] = act2<ACT>(fma2(pk2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), es2, eh2));
          if (!PLAIN) {
            const uint64_t ps2 = pk2(k.ps, k.ps), fs2 = pk2(k.fs, k.fs), fh2 = pk2(k.fh, k.fh);
            if (has_res) {
              float res[16];
              unpack8_bf16(lds128(rbase + (uint32_t)(((2 * c) ^ sw) << 4)), res);
              unpack8_bf16(lds128(rbase + (uint32_t)(((2 * c + 1) ^ sw) << 4)), res + 8);
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = fma2(fma2(v[i], ps2, pk2(res[2 * i], res[2 * i + 1])), fs2, fh2);
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = fma2(mul2(v[i], ps2), fs2, fh2);
            }
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) upk2(v[i], y[2 * i], y[2 * i + 1]);
        }
        if (!PLAIN && (want_sums || want_mm)) {
          const bool whole = ok && q0 + 16 <= P;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            if (whole || (ok && q0 + i < P)) {
              ssum += y[i]; ssq = fmaf(y[i], y[i], ssq);
              vmax = fmaxf(vmax, y[i]); vmin = fminf(vmin, y[i]);
            }
          }
        }
        if (odt == VRCOC_BF16) {
          float lo[8], hi[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) { lo[i] = y[i]; hi[i] = y[8 + i]; }
          sts128(rbase + (uint32_t)(((2 * c) ^ sw) << 4), pack8_bf16(lo));
          sts128(rbase + (uint32_t)(((2 * c + 1) ^ sw) << 4), pack8_bf16(hi));
        } else {
          // fp32: a 128-byte row is 32 points; chunks 0,1 fill the region, leave, then chunks 2,3
          if (c == 2) {
            if (lane == 0) tma_store_wait_read();
            __syncwarp();
          }
#pragma unroll
          for (int i = 0; i < 4; ++i)
            sts128(rbase + (uint32_t)(((((c & 1) << 2) + i) ^ sw) << 4),
                   make_uint4(__float_as_uint(y[4 * i]), __float_as_uint(y[4 * i + 1]), __float_as_uint(y[4 * i + 2]), __float_as_uint(y[4 * i + 3])));
          if (c & 1) {
            fence_async_smem();
            __syncwarp();
            if (lane == 0) { tma_store_3d(tmO, region, q_base + 32 * (c >> 1), ochan0, b); tma_store_commit(); }
          }
        }
      }
    }
    if (row_live && odt == VRCOC_BF16) {
      fence_async_smem();
      __syncwarp();
      if (lane == 0) { tma_store_3d(tmO, region, q_base, ochan0, b); tma_store_commit(); }
    }
    // this warp is done reading TMEM buffer `buf`
    tc_fence_before();
    __syncwarp();
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&acc_empty[buf])) : "memory");
    if (tr) t_comp += gtime() - t1;
  }
  if (g_tc_trace != nullptr && threadIdx.x == 0) {
    unsigned long long* trp = g_tc_trace + ((blockIdx.z * gridDim.y + blockIdx.y) * (size_t)gridDim.x + blockIdx.x) * 8;
    trp[6] = t_wait;
    trp[7] = t_comp;
  }
  if (lane == 0) tma_store_wait_read();                            // shared memory must outlive the bulk stores' reads
  if (want_sums) {
    ssum = warp_sum(ssum);
    ssq = warp_sum(ssq);
    if (lane == 0) {
      const int slot = (blockIdx.x * 8 + warp + 7 * blockIdx.y) & (VRCOC_STAT_SLOTS - 1);
      double* dst = a.out_sample_sums + ((int64_t)b * VRCOC_STAT_SLOTS + slot) * 2;
      atomicAdd(dst, (double)ssum);
      atomicAdd(dst + 1, (double)ssq);
    }
  }
  if (want_mm) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
      vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, s));
      vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, s));
    }
    if (lane == 0) {
      atomicMax(&a.out_minmax[0], __float_as_uint(vmax));
      atomicMax(&a.out_minmax[1], ~__float_as_uint(vmin));
    }
  }
}

template <typename TS, int XMODE>
__global__ void __launch_bounds__(TQ_THREADS, 2) conv_tc_cm_kernel(ConvArgs a, TqLayout L, const __grid_constant__ CUtensorMap tmapW,
                                                                const __grid_constant__ CUtensorMap tmapX,
                                                                const __grid_constant__ CUtensorMap tmapO1,
                                                                const __grid_constant__ CUtensorMap tmapO2,
                                                                const __grid_constant__ CUtensorMap tmapR) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  unsigned char* smem = smem_raw + pad;
  unsigned char* sX = smem + L.off_x;                                // XMODE 0: [stages][16 KB]; else [nslabs][16 KB]
  unsigned char* sW = smem + L.off_w;                                // [stages][16 KB]
  unsigned char* stage = smem + L.off_stage;                         // [8 warps][4 KB] epilogue staging
  float4* tab = reinterpret_cast<float4*>(smem + L.off_tab);
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + L.off_bar);
  uint64_t* bar_free = bar_full + TQ_MAX_STAGES;
  uint64_t* acc_full = bar_free + TQ_MAX_STAGES;                     // [2]
  uint64_t* acc_empty = acc_full + 2;                                // [2]
  uint64_t* res_bar = acc_empty + 2;                                 // [8] one per epilogue warp
  uint64_t* x_full = res_bar + 8;                                    // [9] XMODE 3: raw slab landed
  uint64_t* x_ready = x_full + TQ_MAX_SLABS3;                        // [9] XMODE 3: slab transformed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(x_ready + TQ_MAX_SLABS3);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.z, p0 = blockIdx.x * TQ_NP;
  const int o_begin = blockIdx.y * L.tiles * TQ_MT;
  const int P = a.P_out;
  int tiles = (a.O - o_begin + TQ_MT - 1) / TQ_MT;
  if (tiles > L.tiles) tiles = L.tiles;
  const int nk = L.nslabs;
  const int nkx = L.nx;
  const int ST = L.stages;
  // folded GroupNorm: the weight rows are [hi | lo]; a feat tile (rows < O_split) contracts both halves against the same nkx
  // raw X slabs, a value tile only the hi half
  const bool fold = XMODE <= 1 && a.gn_fold_k1 != nullptr;        // X resident (1) or streamed with the weights (0: large C)
  // (when the feat | value boundary is not on a tile boundary every tile contracts both halves: the caller then supplies a lo half
  // for the value rows as well)
  const bool value_hi_only = (a.O_split % TQ_MT) == 0;
  auto slabs_of = [&](int j) { return fold ? ((value_hi_only && o_begin + j * TQ_MT >= a.O_split) ? nkx : 2 * nkx) : nk; };
  int total_steps = 0;
  for (int j = 0; j < tiles; ++j) total_steps += slabs_of(j);

  // one ring step = the weight slab (tile j, k-slab kc) and, when X is streamed / fetched with the first tile, its X slab.
  // Steps are issued in order by ONE thread (tid 256): (pj, pkc) walk the (tile, slab) pairs.
  int pj = 0, pkc = 0;
  auto issue = [&](int it) {
    const int s = it % ST;
    const int j = pj, kc = pkc;
    if (++pkc == slabs_of(pj)) { pkc = 0; ++pj; }
    if (it >= ST) mbar_wait(&bar_free[s], (uint32_t)((it / ST) - 1) & 1);
    const bool need_x = XMODE == 0 || (XMODE == 1 && j == 0 && kc < nkx);
    mbar_expect_tx(&bar_full[s], (uint32_t)(TQ_W_BYTES + (need_x ? TQ_X_BYTES : 0)));
    tma_load_2d(sW + s * TQ_W_BYTES, &tmapW, kc * TC_BK, o_begin + j * TQ_MT, &bar_full[s]);
    if (need_x) {
      const int kx = (fold && kc >= nkx) ? kc - nkx : kc;               // the lo half of the fold contracts the same x slabs again
      unsigned char* dst = sX + (XMODE == 0 ? s : kc) * TQ_X_BYTES;
      tma_load_xbox(a, dst, &tmapX, p0, kx * TC_BK, b, &bar_full[s]);
      tma_load_xbox(a, dst + TC_A_LBO, &tmapX, p0 + 64, kx * TC_BK, b, &bar_full[s]);
    }
  };
  const int early_steps = total_steps < ST ? total_steps : ST;       // need no free-slot wait

  if (tid == 0) trace(0);
  if (warp == 9) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)(2 * TQ_NP)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 256) {
    for (int i = 0; i < ST; ++i) { mbar_init(&bar_full[i], 1); mbar_init(&bar_free[i], 1); }
    mbar_init(&acc_full[0], 1); mbar_init(&acc_full[1], 1);
    mbar_init(&acc_empty[0], 8); mbar_init(&acc_empty[1], 8);      // one arrival per epilogue warp
    for (int i = 0; i < 8; ++i) mbar_init(&res_bar[i], 1);
    if (XMODE == 3)
      for (int i = 0; i < TQ_MAX_SLABS3; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_ready[i], 8); }
    mbar_fence_init();
    fence_async_smem();
    // this thread is the TMA producer: everything that depends on nothing goes out NOW, under the prologue-table build and
    // the TMEM allocation (the per-CTA trace showed 1.1 us of set-up before the first request and the CTA then waiting
    // ~2 us for it): every raw X slab (XMODE 3) and the first ring-depth weight slabs
    if (XMODE == 3) {
      for (int kc = 0; kc < nk; ++kc) {
        // two-source input (virtual concat [src0 | src1], C0 a multiple of 64): the second source's map rides in the slot of
        // the second output, which such launches do not have
        const bool second = a.C1 > 0 && kc * TC_BK >= a.C0;
        const CUtensorMap* tx = second ? &tmapO2 : &tmapX;
        const int ch = kc * TC_BK - (second ? a.C0 : 0);
        mbar_expect_tx(&x_full[kc], (uint32_t)TQ_X_BYTES);
        tma_load_3d(sX + kc * TQ_X_BYTES, tx, p0, ch, b, &x_full[kc]);
        tma_load_3d(sX + kc * TQ_X_BYTES + TC_A_LBO, tx, p0 + 64, ch, b, &x_full[kc]);
      }
    }
    for (int it = 0; it < early_steps; ++it) issue(it);
    tma_prefetch_desc(&tmapO1);
  }
  if (XMODE >= 2) build_prologue_table(a, b, tab);                   // strides by blockDim.x: every thread takes part
  tc_fence_before();
  __syncthreads();                                                     // barriers + TMEM slot + table visible
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    // ---- TMA producer: the rest of the ring (every step from here on waits for a slot the MMAs have released) -------------
    if (XMODE == 2) __syncthreads();                                   // (A)
    if (lane == 0)
      for (int it = early_steps; it < total_steps; ++it) issue(it);
    __syncwarp();
  } else if (warp == 9) {
    if (XMODE == 2) __syncthreads();                                   // (A)
    // ---- MMA issuer -----------------------------------------------------------------------------------------------------
    if (lane == 0) {
      tc_fence_after();
      const uint32_t idesc = make_idesc_cm(TQ_NP);
      int s = 0;                                                       // ring slot and its phase, advanced without divisions:
      uint32_t ph = 0;                                                 // this loop is the critical path of the CTA
      for (int j = 0; j < tiles; ++j) {
        const int buf = j & 1;
        if (j >= 2) { mbar_wait(&acc_empty[buf], (uint32_t)((j >> 1) - 1) & 1); tc_fence_after(); }
        const uint32_t tacc = tmem_base + (uint32_t)(buf * TQ_NP);
        const int nkj = slabs_of(j);
        for (int kc = 0; kc < nkj; ++kc) {
          mbar_wait(&bar_full[s], ph);
          if (XMODE == 3 && j == 0) mbar_wait(&x_ready[kc], 0);
          tc_fence_after();
          const int kx = (fold && kc >= nkx) ? kc - nkx : kc;             // X slab (the lo half of the fold re-reads the slabs)
          const int ksteps = (min(TC_BK, a.K - kx * TC_BK) + 15) >> 4;
          const uint32_t w_addr = smem_u32(sW + s * TQ_W_BYTES);
          const uint32_t x_addr = smem_u32(sX + (XMODE == 0 ? s : kx) * TQ_X_BYTES);
          // A = weights, K-major: 16 k = 32 B inside the 128 B row; B = activations, MN-major: 16 k-rows = two 8-row groups
          tc_issue_slab<32 / 16, 2048 / 16>(tacc, tc_desc_lo(w_addr, 16), tc_desc_lo(x_addr, TC_A_LBO), idesc, kc > 0 ? 1u : 0u, ksteps);
          tc_commit(&bar_free[s]);
          if (++s == ST) { s = 0; ph ^= 1u; }
        }
        tc_commit(&acc_full[buf]);
      }
    }
    __syncwarp();
  } else {
    if (XMODE == 2) {
      // ---- X operand: transform on load, all K slabs resident ----------------------------------------------------------------
      const int a_chunk = tid & 15, a_krow0 = tid >> 4, a_blk = a_chunk >> 3, a_c = a_chunk & 7;
      const int q0 = p0 + a_chunk * 8;
      RawSlab<TS> raw;
      if (tid == 0) trace(1);
      slab_gload<TS, true>(a, b, 0, a_krow0, q0, P, 1, raw);
      for (int kc = 0; kc < nk; ++kc) {
        const int ksteps = (min(TC_BK, a.K - kc * TC_BK) + 15) >> 4;
        RawSlab<TS> cur = raw;
        if (kc + 1 < nk) slab_gload<TS, true>(a, b, kc + 1, a_krow0, q0, P, 1, raw);
        if (a.has_gate) slab_sstore<TS, true, true>(kc, ksteps, a_krow0, a_blk, a_c, 1, 0, a.Cin, tab, cur, sX + kc * TQ_X_BYTES);
        else slab_sstore<TS, true, false>(kc, ksteps, a_krow0, a_blk, a_c, 1, 0, a.Cin, tab, cur, sX + kc * TQ_X_BYTES);
      }
      fence_async_smem();
      if (tid == 0) trace(2);
      __syncthreads();                                                 // (A)
    }
    if (XMODE == 3) {
      // ---- X operand: in-place prologue on the TMA-landed raw slabs -----------------------------------------------------------
      if (tid == 0) trace(1);
      for (int kc = 0; kc < nk; ++kc) {
        mbar_wait(&x_full[kc], 0);
        const uint32_t slab = smem_u32(sX + kc * TQ_X_BYTES);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int u = tid + 256 * i;                                   // 16-byte chunk of the slab; k-row = (u >> 3) & 63
          const int chn = kc * TC_BK + ((u >> 3) & 63);
          if (chn < a.Cin) {                                             // rows past Cin stay zero (TMA fill)
            const float4 t = tab[chn];
            float f[8];
            unpack8_bf16(lds128(slab + (uint32_t)u * 16u), f);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float x = f[e];
              float y = fmaf(x, t.x, t.y);
              if (a.has_gate) y *= sigmoid_fast(fmaf(t.z, x, t.w));
              f[e] = y;
            }
            sts128(slab + (uint32_t)u * 16u, pack8_bf16(f));
          }
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&x_ready[kc])) : "memory");
      }
      if (tid == 0) trace(2);
    }
#define CM_EPI(ACTV, PL)                                                                                                  \
  cm_epilogue<ACTV, PL>(a, tmem_base, acc_full, acc_empty, &res_bar[warp], stage + warp * 4096, L.stage_bufs == 2 ? 8 * 4096 : 0, \
                        &tmapO1, &tmapO2, &tmapR, b, p0, \
                        o_begin, tiles)
    if (L.plain) {
      switch (a.act) {
        case VRCOC_ACT_NONE: CM_EPI(VRCOC_ACT_NONE, true); break;
        case VRCOC_ACT_RELU: CM_EPI(VRCOC_ACT_RELU, true); break;
        case VRCOC_ACT_GELU: CM_EPI(VRCOC_ACT_GELU, true); break;
        case VRCOC_ACT_SILU: CM_EPI(VRCOC_ACT_SILU, true); break;
        default: CM_EPI(VRCOC_ACT_LRELU, true); break;
      }
    } else {
      switch (a.act) {
        case VRCOC_ACT_NONE: CM_EPI(VRCOC_ACT_NONE, false); break;
        case VRCOC_ACT_RELU: CM_EPI(VRCOC_ACT_RELU, false); break;
        case VRCOC_ACT_GELU: CM_EPI(VRCOC_ACT_GELU, false); break;
        case VRCOC_ACT_SILU: CM_EPI(VRCOC_ACT_SILU, false); break;
        default: CM_EPI(VRCOC_ACT_LRELU, false); break;
      }
    }
#undef CM_EPI
    if (tid == 0) trace(4);
  }
  tc_fence_before();
  __syncthreads();
  if (tid == 0) trace(5);
  if (warp == 9) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * TQ_NP)));
  }
}

}  // namespace vrcoc
