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
//       2: X resident, built by the CTA's threads through the prologue table (GroupNorm / gate / plain fp32 sources).
#pragma once

namespace vrcoc {

constexpr int TQ_THREADS = 320;
constexpr int TQ_NP = 128;                          // points per tile  (MMA N)
constexpr int TQ_MT = 128;                          // outputs per tile (MMA M)
constexpr int TQ_W_BYTES = TQ_MT * TC_BK * 2;       // 16 KB
constexpr int TQ_X_BYTES = TQ_NP * TC_BK * 2;       // 16 KB
constexpr int TQ_MAX_STAGES = 4;

struct TqLayout {
  int stages;        // ring depth
  int nslabs;        // K slabs
  int tiles;         // output tiles per CTA
  int plain;         // epilogue is act(acc*es + eh) only
  int off_x, off_w, off_stage, off_tab, off_bar, total;
};

// instruction descriptor: D=f32, A=B=bf16, A K-major, B MN-major, N = points, M = 128
__host__ __device__ constexpr uint32_t make_idesc_cm(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (0u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TQ_MT >> 4) << 24);
}

struct CmCoef { float es, eh, ps, fs, fh; };

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
template <int ACT, bool PLAIN>
__device__ __forceinline__ void cm_epilogue(const ConvArgs& a, uint32_t tmem_base, uint64_t* acc_full, uint64_t* acc_empty, uint64_t* rbar,
                                            unsigned char* region, const CUtensorMap* tmO1, const CUtensorMap* tmO2,
                                            const CUtensorMap* tmR, int b, int p0, int o_begin, int tiles) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int lq = warp & 3, ch = warp >> 2;
  const int P = a.P_out;
  const int q_base = p0 + ch * (TQ_NP / 2);
  const bool has_res = !PLAIN && a.res != nullptr;
  const bool want_sums = !PLAIN && a.out_sample_sums != nullptr;
  const bool want_mm = !PLAIN && a.out_minmax != nullptr;
  const uint32_t rbase = smem_u32(region) + lane * 128;
  const int sw = lane & 7;
  uint32_t res_phase = 0;
  float ssum = 0.f, ssq = 0.f, vmax = 0.f, vmin = __int_as_float(0x7f800000);
  for (int j = 0; j < tiles; ++j) {
    const int buf = j & 1;
    const int o_row0 = o_begin + j * TQ_MT + lq * 32;
    const int o = o_row0 + lane;
    const bool ok = o < a.O;
    const bool row_live = o_row0 < a.O;                              // warp-uniform
    if (j > 0) {                                                     // the previous tile's store has finished reading the region
      if (lane == 0) tma_store_wait_read();
      __syncwarp();
    }
    if (has_res && row_live && lane == 0) {
      mbar_expect_tx(rbar, 4096u);
      tma_load_3d(region, tmR, q_base, o_row0, b, rbar);
    }
    CmCoef k;
    k.es = (ok && a.e_scale) ? __ldg(a.e_scale + o) : 1.f;
    k.eh = (ok && a.e_shift) ? __ldg(a.e_shift + o) : 0.f;
    k.ps = (!PLAIN && ok && a.post_scale) ? __ldg(a.post_scale + o) : 1.f;
    k.fs = (!PLAIN && ok && a.f_scale) ? __ldg(a.f_scale + o) : 1.f;
    k.fh = (!PLAIN && ok && a.f_shift) ? __ldg(a.f_shift + o) : 0.f;
    const bool second = o_row0 >= a.O_split;                         // warp-uniform: O_split % 32 == 0 (host-checked)
    const int odt = second ? a.out2_dtype : a.out_dtype;
    const CUtensorMap* tmO = second ? tmO2 : tmO1;
    const int ochan0 = second ? o_row0 - a.O_split : o_row0;
    mbar_wait(&acc_full[buf], (uint32_t)(j >> 1) & 1);
    tc_fence_after();
    if (threadIdx.x == 0 && j == 0) trace(3);
    if (has_res && row_live) { mbar_wait(rbar, res_phase); res_phase ^= 1u; }
    const uint32_t tbase = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(buf * TQ_NP + ch * (TQ_NP / 2));
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t r[16];
      tmem_ld16(tbase + (uint32_t)(16 * c), r);
      if (row_live) {
        const int q0 = q_base + 16 * c;
        float y[16];
        if (PLAIN) {
#pragma unroll
          for (int i = 0; i < 16; ++i) y[i] = act_tc<ACT>(fmaf(__uint_as_float(r[i]), k.es, k.eh));
        } else {
          float res[16];
          if (has_res) {
            unpack8_bf16(lds128(rbase + (uint32_t)(((2 * c) ^ sw) << 4)), res);
            unpack8_bf16(lds128(rbase + (uint32_t)(((2 * c + 1) ^ sw) << 4)), res + 8);
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) res[i] = 0.f;
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float v = act_tc<ACT>(fmaf(__uint_as_float(r[i]), k.es, k.eh));
            v = fmaf(v, k.ps, res[i]);
            y[i] = fmaf(v, k.fs, k.fh);
          }
          if (want_sums || want_mm) {
            const bool whole = ok && q0 + 16 <= P;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              if (whole || (ok && q0 + i < P)) {
                ssum += y[i]; ssq = fmaf(y[i], y[i], ssq);
                vmax = fmaxf(vmax, y[i]); vmin = fminf(vmin, y[i]);
              }
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
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 8);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.z, p0 = blockIdx.x * TQ_NP;
  const int o_begin = blockIdx.y * L.tiles * TQ_MT;
  const int P = a.P_out;
  int tiles = (a.O - o_begin + TQ_MT - 1) / TQ_MT;
  if (tiles > L.tiles) tiles = L.tiles;
  const int nk = L.nslabs;
  const int ST = L.stages;

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
    mbar_fence_init();
    tma_prefetch_desc(&tmapW);
    if (XMODE != 2) tma_prefetch_desc(&tmapX);
    tma_prefetch_desc(&tmapO1);
  }
  if (XMODE == 2) build_prologue_table(a, b, tab);                   // strides by blockDim.x: every thread takes part
  tc_fence_before();
  __syncthreads();                                                     // barriers + TMEM slot + table visible
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    // ---- TMA producer ---------------------------------------------------------------------------------------------------
    const int total = tiles * nk;
    auto issue = [&](int it) {
      const int s = it % ST;
      const int j = it / nk, kc = it - j * nk;
      if (it >= ST) mbar_wait(&bar_free[s], (uint32_t)((it / ST) - 1) & 1);
      const bool need_x = XMODE == 0 || (XMODE == 1 && j == 0);
      mbar_expect_tx(&bar_full[s], (uint32_t)(TQ_W_BYTES + (need_x ? TQ_X_BYTES : 0)));
      tma_load_2d(sW + s * TQ_W_BYTES, &tmapW, kc * TC_BK, o_begin + j * TQ_MT, &bar_full[s]);
      if (need_x) {
        unsigned char* dst = sX + (XMODE == 0 ? s : kc) * TQ_X_BYTES;
        tma_load_3d(dst, &tmapX, p0, kc * TC_BK, b, &bar_full[s]);
        tma_load_3d(dst + TC_A_LBO, &tmapX, p0 + 64, kc * TC_BK, b, &bar_full[s]);
      }
    };
    if (XMODE == 2) {
      // the first ST slabs need no free-slot wait and go out while the other warps still build X; the rest depends on
      // MMA progress and therefore has to come after barrier (A)
      if (lane == 0)
        for (int it = 0; it < total && it < ST; ++it) issue(it);
      __syncwarp();
      __syncthreads();                                                 // (A)
      if (lane == 0)
        for (int it = ST; it < total; ++it) issue(it);
    } else if (lane == 0) {
      for (int it = 0; it < total; ++it) issue(it);
    }
    __syncwarp();
  } else if (warp == 9) {
    if (XMODE == 2) __syncthreads();                                   // (A)
    // ---- MMA issuer -----------------------------------------------------------------------------------------------------
    if (lane == 0) {
      tc_fence_after();
      const uint32_t idesc = make_idesc_cm(TQ_NP);
      int it = 0;
      for (int j = 0; j < tiles; ++j) {
        const int buf = j & 1;
        if (j >= 2) { mbar_wait(&acc_empty[buf], (uint32_t)((j >> 1) - 1) & 1); tc_fence_after(); }
        const uint32_t tacc = tmem_base + (uint32_t)(buf * TQ_NP);
        for (int kc = 0; kc < nk; ++kc, ++it) {
          const int s = it % ST;
          mbar_wait(&bar_full[s], (uint32_t)(it / ST) & 1);
          tc_fence_after();
          const int ksteps = (min(TC_BK, a.K - kc * TC_BK) + 15) >> 4;
          const uint32_t w_addr = smem_u32(sW + s * TQ_W_BYTES);
          const uint32_t x_addr = smem_u32(sX + (XMODE == 0 ? s : kc) * TQ_X_BYTES);
          for (int t = 0; t < ksteps; ++t) {
            const uint64_t wd = make_desc(w_addr + t * 32, 16, 1024);            // 16 k = 32 B inside the 128 B row
            const uint64_t xd = make_desc(x_addr + t * 2048, TC_A_LBO, 1024);    // 16 k-rows = two 8-row groups
            tc_mma(tacc, wd, xd, idesc, (kc > 0 || t > 0) ? 1u : 0u);
          }
          tc_commit(&bar_free[s]);
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
#define CM_EPI(ACTV, PL)                                                                                                  \
  cm_epilogue<ACTV, PL>(a, tmem_base, acc_full, acc_empty, &res_bar[warp], stage + warp * 4096, &tmapO1, &tmapO2, &tmapR, b, p0, \
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
