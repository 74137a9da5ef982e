// tcgen05 path, kernel 5: the channel MLP of a ClusterBlock in ONE kernel (included by conv_tc.cu)
//
//   out = x + ls * ( W2 . gelu( W1 . GN(x) + b1 ) + b2 )          (reference vr_coc.py:208-228 Mlp, :264-275 ClusterBlock)
//
// As two launches the hidden activation (mlp_ratio 8: 512 / 1024 channels at the two large stages) is written to and read
// back from HBM — 2 x 134 MB per block at stage 1, more than everything else the block moves.  Here it never leaves the SM:
//   * one CTA owns a 128-point tile; X (C <= 128 channels) arrives by TMA and is GroupNorm'ed in place (XMODE 3 of
//     conv_tc_cm.cuh);
//   * the hidden layer is produced 128 channels at a time:  acc1[128 hid x 128 pts] = W1 chunk . X  (tcgen05, TMEM cols 0-127);
//     the 8 epilogue warps read it, add b1, apply the erf GELU and write it as bf16 straight into the MN-major SW128 operand
//     layout (lane = hidden channel = k-row of the second GEMM, 16 consecutive points = two 16-byte chunks);
//   * acc2[C x 128 pts] += W2 chunk . H  (TMEM cols 128-255) accumulates over the chunks;
//   * the last epilogue is cm_epilogue: + b2, layer scale, residual (TMA), GroupNorm statistics of the result for the next
//     block, TMA store.
// Both weight matrices stream through one TMA ring in the order the MMA thread consumes them.
// C <= 128 (stages 1, 2): TMEM 256 columns, one hidden buffer, two CTAs per SM (one's GELU overlaps the other's loads).
// C <= 384 (stage 3):     acc2 is ceil(C/128) tiles side by side (TMEM 512 columns), two hidden buffers so that the second GEMM
//                         of chunk j runs under the GELU of chunk j+1, one CTA per SM.
#pragma once

#ifndef VRCOC_MF_TRACE
#define VRCOC_MF_TRACE 0   // 1: the MMA thread and epilogue warp 0 record their barrier wait cycles (tools/trace_mlpf.py --waits)
#endif
#if VRCOC_MF_TRACE
#define MF_T0() t_ = clock64()
#define MF_T1(acc) acc += clock64() - t_
// per-chunk time stamps (C > 128 variant): slot k of chunk j at g_tc_trace[4096*16 + cta*128 + j*8 + k] (clock64, same SM for all roles)
#define MF_STAMP(j, k)                                                                                                     \
  do {                                                                                                                     \
    if (g_tc_trace && (j) < 16) g_tc_trace[4096 * 16 + (blockIdx.z * (size_t)gridDim.x + blockIdx.x) * 128 + (j) * 8 + (k)] = clock64(); \
  } while (0)
#else
#define MF_T0()
#define MF_T1(acc)
#define MF_STAMP(j, k)
#endif

namespace vrcoc {

constexpr int MF_THREADS = 320;                 // C <= 128: warps 0-7 epilogue, 8 TMA producer, 9 MMA issuer
constexpr int MF_THREADS_WIDE = 384;            // C > 128: + warps 10, 11 = two more weight producers (see the producer block)
constexpr int MF_PRODUCERS_WIDE = 3;
constexpr int MF_MAX_STAGES = 8;
constexpr int MF_MAX_NK1 = 6;

struct MlpLayout {
  int nk1;           // k-slabs of the first GEMM  (C / 64, 1..2)
  int nh;            // hidden chunks of 128
  int stages;        // weight ring depth
  int mt2;           // output tiles of 128 channels (ceil(C / 128), 1..3)
  int h_bufs;        // hidden buffers in shared memory (1 or 2)
  int tmem_cols;     // 256 or 512
  int off_ring, off_h, off_tab, off_bar, total;
};

template <bool SINGLE>      // SINGLE: C <= 128 — one output tile, one hidden buffer (compile-time, the hot configuration)
__global__ void __launch_bounds__(SINGLE ? MF_THREADS : MF_THREADS_WIDE, SINGLE ? 2 : 1)
mlp_fused_kernel(ConvArgs a1, ConvArgs a2, MlpLayout L, const float* __restrict__ b1, const __grid_constant__ CUtensorMap tmapX,
                 const __grid_constant__ CUtensorMap tmapW1, const __grid_constant__ CUtensorMap tmapW2,
                 const __grid_constant__ CUtensorMap tmapO, const __grid_constant__ CUtensorMap tmapR) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  unsigned char* smem = smem_raw + pad;
  unsigned char* sX = smem;                                          // [nk1][16 KB]
  unsigned char* ring = smem + L.off_ring;                           // [stages][16 KB]
  unsigned char* sH = smem + L.off_h;                                // [h_bufs][2][16 KB] hidden chunk(s), bf16 operand layout
  float4* tab = reinterpret_cast<float4*>(smem + L.off_tab);
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + L.off_bar);
  uint64_t* bar_free = bar_full + MF_MAX_STAGES;
  uint64_t* x_full = bar_free + MF_MAX_STAGES;                       // [MF_MAX_NK1]
  uint64_t* x_ready = x_full + MF_MAX_NK1;                           // [MF_MAX_NK1]
  uint64_t* acc1_full = x_ready + MF_MAX_NK1;
  uint64_t* acc1_empty = acc1_full + 1;
  uint64_t* h_full = acc1_empty + 1;                                 // [2]
  uint64_t* h_empty = h_full + 2;                                    // [2]
  uint64_t* acc2_full = h_empty + 2;
  uint64_t* acc2_empty = acc2_full + 1;                              // never waited for (single output tile)
  uint64_t* res_bar = acc2_empty + 1;                                // [8]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 8);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.z, p0 = blockIdx.x * TQ_NP;
  const int nk1 = L.nk1, nh = L.nh, ST = L.stages;
  const int MT2 = SINGLE ? 1 : L.mt2, HB = SINGLE ? 1 : L.h_bufs;

  if (tid == 0) trace(0);
  if (warp == 9) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)L.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 256) {
    for (int i = 0; i < ST; ++i) { mbar_init(&bar_full[i], 1); mbar_init(&bar_free[i], 1); }
    for (int i = 0; i < MF_MAX_NK1; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_ready[i], 8); }
    mbar_init(acc1_full, 1); mbar_init(acc1_empty, 8);
    for (int i = 0; i < 2; ++i) { mbar_init(&h_full[i], 8); mbar_init(&h_empty[i], 1); }
    mbar_init(acc2_full, 1); mbar_init(acc2_empty, 8);
    for (int i = 0; i < 8; ++i) mbar_init(&res_bar[i], 1);
    mbar_fence_init();
    tma_prefetch_desc(&tmapX); tma_prefetch_desc(&tmapW1); tma_prefetch_desc(&tmapW2);
    tma_prefetch_desc(&tmapO); tma_prefetch_desc(&tmapR);
  }
  build_prologue_table(a1, b, tab);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (!SINGLE && (warp == 8 || warp == 10)) {
    // ---- C > 128: TWO weight rings, one per GEMM, each with its own producer and its own MMA-issuing warp ----------------------
    // The barrier-wait trace of the single-issuer form (tools/trace_mlpf.py on a VRCOC_MF_TRACE build, stage 3: C = 320) showed
    // the MMA thread waiting for weights 4 us, for the epilogue 4 us and ISSUING for 27 of the 39 us of its loop: 0.24 us per
    // 16 KB slab (4 MMAs + barrier wait + commit, ~65 dependent SASS instructions on a scheduler it shares with two GELU warps)
    // against 0.13 us of tensor-pipe time.  The two GEMMs of a chunk write different accumulators, so they are issued by two
    // warps on different schedulers: warp 9 = first GEMM (W1 ring, stages [0, SA)), warp 11 = second GEMM (W2 ring, stages
    // [SA, ST)); warp 8 feeds ring A (and X), warp 10 ring B.
    const int SA = ST / 2, SB = ST - SA;
    if (lane == 0) {
      if (warp == 8) {
        for (int kc = 0; kc < nk1; ++kc) {
          mbar_expect_tx(&x_full[kc], (uint32_t)TQ_X_BYTES);
          tma_load_3d(sX + kc * TQ_X_BYTES, &tmapX, p0, kc * TC_BK, b, &x_full[kc]);
          tma_load_3d(sX + kc * TQ_X_BYTES + TC_A_LBO, &tmapX, p0 + 64, kc * TC_BK, b, &x_full[kc]);
        }
        int sl = 0;
        uint32_t ph = 1;                                               // parity of the slot's previous phase: passes on a fresh barrier
        for (int j = 0; j < nh; ++j)
          for (int kc = 0; kc < nk1; ++kc) {
            mbar_wait(&bar_free[sl], ph);
            mbar_expect_tx(&bar_full[sl], (uint32_t)TQ_W_BYTES);
            tma_load_2d(ring + sl * TQ_W_BYTES, &tmapW1, kc * TC_BK, j * TQ_MT, &bar_full[sl]);
            if (++sl == SA) { sl = 0; ph ^= 1u; }
          }
      } else {
        int sl = 0;
        uint32_t ph = 1;
        for (int j = 0; j < nh; ++j)
          for (int q = 0; q < 2 * MT2; ++q) {
            mbar_wait(&bar_free[SA + sl], ph);
            mbar_expect_tx(&bar_full[SA + sl], (uint32_t)TQ_W_BYTES);
            tma_load_2d(ring + (SA + sl) * TQ_W_BYTES, &tmapW2, j * TQ_MT + (q & 1) * TC_BK, (q >> 1) * TQ_MT, &bar_full[SA + sl]);
            if (++sl == SB) { sl = 0; ph ^= 1u; }
          }
      }
    }
    __syncwarp();
  } else if (!SINGLE && (warp == 9 || warp == 11)) {
    // ---- C > 128: MMA issuers (warp 9: hidden = W1 . X per chunk; warp 11: out += W2 . H per chunk) --------------------------------
    const int SA = ST / 2, SB = ST - SA;
    if (lane == 0) {
      const uint32_t idesc = make_idesc_cm(TQ_NP);
      const uint32_t acc1 = tmem_base, acc2 = tmem_base + TQ_NP;
      int sl = 0;
      uint32_t ph = 0;
#if VRCOC_MF_TRACE
      long long tw_ = 0, ta_ = 0, th_ = 0, t_; const long long tl0_ = clock64();
      unsigned long long* q_ = g_tc_trace ? g_tc_trace + 4096 * 8 + (blockIdx.z * (size_t)gridDim.x + blockIdx.x) * 8 : nullptr;
#endif
      if (warp == 9) {
        for (int j = 0; j < nh; ++j) {
          if (j > 0) {
            MF_T0();
            mbar_wait(acc1_empty, (uint32_t)(j - 1) & 1);              // the epilogue warps have read hidden chunk j-1 out of TMEM
            MF_T1(ta_);
            tc_fence_after();
          }
          MF_STAMP(j, 0);
          for (int kc = 0; kc < nk1; ++kc) {
            if (j == 0) mbar_wait(&x_ready[kc], 0);
            MF_T0();
            mbar_wait(&bar_full[sl], ph);
            MF_T1(tw_);
            tc_fence_after();
            tc_issue_slab<32 / 16, 2048 / 16>(acc1, tc_desc_lo(smem_u32(ring + sl * TQ_W_BYTES), 16),
                                              tc_desc_lo(smem_u32(sX + kc * TQ_X_BYTES), TC_A_LBO), idesc, kc == 0 ? 0u : 1u, 4);
            tc_commit(&bar_free[sl]);
            if (++sl == SA) { sl = 0; ph ^= 1u; }
          }
          tc_commit(acc1_full);
          MF_STAMP(j, 1);
        }
#if VRCOC_MF_TRACE
        if (q_) { q_[0] = tw_; q_[1] = ta_; q_[3] = clock64() - tl0_; }
#endif
      } else {
        for (int j = 0; j < nh; ++j) {
          const int hb = j % HB;
          MF_T0();
          mbar_wait(&h_full[hb], (uint32_t)(j / HB) & 1);              // hidden chunk j is in shared memory (bf16 operand layout)
          MF_T1(th_);
          MF_STAMP(j, 2);
          tc_fence_after();
          const uint32_t hbuf = smem_u32(sH + hb * 2 * TQ_X_BYTES);
          for (int q = 0; q < 2 * MT2; ++q) {
            MF_T0();
            mbar_wait(&bar_full[SA + sl], ph);
            MF_T1(tw_);
            tc_fence_after();
            tc_issue_slab<32 / 16, 2048 / 16>(acc2 + (uint32_t)((q >> 1) * TQ_NP), tc_desc_lo(smem_u32(ring + (SA + sl) * TQ_W_BYTES), 16),
                                              tc_desc_lo(hbuf + (uint32_t)((q & 1) * TQ_X_BYTES), TC_A_LBO), idesc,
                                              (j == 0 && (q & 1) == 0) ? 0u : 1u, 4);
            tc_commit(&bar_free[SA + sl]);
            if (++sl == SB) { sl = 0; ph ^= 1u; }
          }
          tc_commit(&h_empty[hb]);
          MF_STAMP(j, 3);
          if (j == nh - 1) tc_commit(acc2_full);
        }
#if VRCOC_MF_TRACE
        if (q_) { q_[2] = th_; q_[7] = tw_; }
#endif
      }
    }
    __syncwarp();
  } else if (warp == 8) {
    // ---- C <= 128: TMA producer: X slabs, then the weight slabs in MMA order --------------------------------------------------------
    if (lane == 0) {
      for (int kc = 0; kc < nk1; ++kc) {
        mbar_expect_tx(&x_full[kc], (uint32_t)TQ_X_BYTES);
        tma_load_3d(sX + kc * TQ_X_BYTES, &tmapX, p0, kc * TC_BK, b, &x_full[kc]);
        tma_load_3d(sX + kc * TQ_X_BYTES + TC_A_LBO, &tmapX, p0 + 64, kc * TC_BK, b, &x_full[kc]);
      }
      const int bs = nk1 + 2 * MT2;                                    // steps of one hidden chunk: W1 of the NEXT chunk, then its own W2
      const int total = nk1 + (nh - 1) * bs + 2 * MT2;
      for (int it = 0; it < total; ++it) {
        const CUtensorMap* map;
        int x, y;
        if (it < nk1) {
          map = &tmapW1; x = it * TC_BK; y = 0;
        } else {
          const int u = it - nk1, j = u / bs, r = u - j * bs;
          if (j < nh - 1 && r < nk1) {
            map = &tmapW1; x = r * TC_BK; y = (j + 1) * TQ_MT;
          } else {
            const int q = j < nh - 1 ? r - nk1 : r;
            map = &tmapW2; x = j * TQ_MT + (q & 1) * TC_BK; y = (q >> 1) * TQ_MT;
          }
        }
        const int s = it % ST;
        if (it >= ST) mbar_wait(&bar_free[s], (uint32_t)((it / ST) - 1) & 1);
        mbar_expect_tx(&bar_full[s], (uint32_t)TQ_W_BYTES);
        tma_load_2d(ring + s * TQ_W_BYTES, map, x, y, &bar_full[s]);
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    // ---- MMA issuer -----------------------------------------------------------------------------------------------------
    if (lane == 0) {
      const uint32_t idesc = make_idesc_cm(TQ_NP);
      const uint32_t acc1 = tmem_base, acc2 = tmem_base + TQ_NP;
      int s = 0;                                                       // ring slot and phase, advanced without divisions
      uint32_t ph = 0;
#if VRCOC_MF_TRACE
      long long tw_ = 0, ta_ = 0, th_ = 0, t_; const long long tl0_ = clock64();
#endif
      auto slab = [&](uint32_t tacc, uint32_t x_addr, bool first) {
        MF_T0();
        mbar_wait(&bar_full[s], ph);
        MF_T1(tw_);
        tc_fence_after();
        const uint32_t w_addr = smem_u32(ring + s * TQ_W_BYTES);
        tc_issue_slab<32 / 16, 2048 / 16>(tacc, tc_desc_lo(w_addr, 16), tc_desc_lo(x_addr, TC_A_LBO), idesc, first ? 0u : 1u, 4);
        tc_commit(&bar_free[s]);
        if (++s == ST) { s = 0; ph ^= 1u; }
      };
      auto gemm1 = [&](int j) {
        for (int kc = 0; kc < nk1; ++kc) {
          if (j == 0) mbar_wait(&x_ready[kc], 0);
          slab(acc1, smem_u32(sX + kc * TQ_X_BYTES), kc == 0);
        }
        tc_commit(acc1_full);
      };
      gemm1(0);
      for (int j = 0; j < nh; ++j) {
        if (j + 1 < nh) {
          MF_T0();
          mbar_wait(acc1_empty, (uint32_t)j & 1);                      // the epilogue warps have read hidden chunk j out of TMEM
          MF_T1(ta_);
          tc_fence_after();
          gemm1(j + 1);
        }
        const int hb = j % HB;
        MF_T0();
        mbar_wait(&h_full[hb], (uint32_t)(j / HB) & 1);                // hidden chunk j is in shared memory (bf16 operand layout)
        MF_T1(th_);
        tc_fence_after();
        unsigned char* hbuf = sH + hb * 2 * TQ_X_BYTES;
        for (int m = 0; m < MT2; ++m) {
          slab(acc2 + (uint32_t)(m * TQ_NP), smem_u32(hbuf), j == 0);
          slab(acc2 + (uint32_t)(m * TQ_NP), smem_u32(hbuf + TQ_X_BYTES), false);
        }
        tc_commit(&h_empty[hb]);
        if (j == nh - 1) tc_commit(acc2_full);
      }
#if VRCOC_MF_TRACE
      if (g_tc_trace) {
        unsigned long long* q = g_tc_trace + 4096 * 8 + (blockIdx.z * (size_t)gridDim.x + blockIdx.x) * 8;
        q[0] = tw_; q[1] = ta_; q[2] = th_; q[3] = clock64() - tl0_;
      }
#endif
    }
    __syncwarp();
  } else {
    // ---- GroupNorm in place on the TMA-landed X slabs ---------------------------------------------------------------------------
    if (tid == 0) trace(1);
    for (int kc = 0; kc < nk1; ++kc) {
      mbar_wait(&x_full[kc], 0);
      const uint32_t sl = smem_u32(sX + kc * TQ_X_BYTES);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int u = tid + 256 * i;
        const int chn = kc * TC_BK + ((u >> 3) & 63);
        if (chn < a1.Cin) {
          const float4 t = tab[chn];
          float f[8];
          unpack8_bf16(lds128(sl + (uint32_t)u * 16u), f);
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = fmaf(f[e], t.x, t.y);
          sts128(sl + (uint32_t)u * 16u, pack8_bf16(f));
        }
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&x_ready[kc])) : "memory");
    }
    if (tid == 0) trace(2);

    // ---- hidden layer: TMEM -> + b1 -> GELU -> bf16 -> operand layout of the second GEMM ------------------------------------------
    const int lq = warp & 3, ch = warp >> 2;
    const int hrow = 32 * (lq & 1) + lane;                             // k-row inside the 64-row slab
    const uint32_t hbase0 = smem_u32(sH) + (uint32_t)((lq >> 1) * TQ_X_BYTES + ch * TC_A_LBO + hrow * 128);
    const int sw = lane & 7;
    unsigned char* out_region = sX + (lq >> 1) * TQ_X_BYTES + ch * TC_A_LBO + (lq & 1) * 4096;   // [32 channels][64 points] of X
    const uint32_t tbase = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(ch * (TQ_NP / 2));
#if VRCOC_MF_TRACE
    long long ea_ = 0, eh_ = 0, et_; const long long el0_ = clock64();
#endif
    for (int j = 0; j < nh; ++j) {
      const float bias = __ldg(b1 + j * TQ_MT + lq * 32 + lane);
      const int hb = j % HB;
      const uint32_t hbase = hbase0 + (uint32_t)(hb * 2 * TQ_X_BYTES);
      const uint64_t one2 = pk2(1.f, 1.f), bias2 = pk2(bias, bias);
#if VRCOC_MF_TRACE
      et_ = clock64();
#endif
      mbar_wait(acc1_full, (uint32_t)j & 1);
#if VRCOC_MF_TRACE
      ea_ += clock64() - et_;
      if (tid == 0 && !SINGLE) MF_STAMP(j, 4);
#endif
      tc_fence_after();
      if (tid == 0 && j == 0) trace(3);
      if (MT2 == 1 && j == nh - 1 && lq * 32 < a2.O && lane == 0) {
        // the last first-GEMM has read X: its slab becomes the staging region of the output epilogue and the residual (the raw
        // x again) is fetched into it now, under the last GELU pass
        mbar_expect_tx(&res_bar[warp], 4096u);
        tma_load_3d(out_region, &tmapR, p0 + ch * (TQ_NP / 2), lq * 32, b, &res_bar[warp]);
      }
      // 32 columns (two 16-point operand rows of this lane's hidden channel) at a time: bias, GELU, bf16, operand layout.  sH is first
      // written after the first 16 values are computed: by then the previous chunk's second GEMM (which reads it) has normally completed.
      auto half = [&](const uint32_t (&r)[32], int hh) {
#pragma unroll
        for (int sub = 0; sub < 2; ++sub) {
          const int c = 2 * hh + sub;
          float y[16];
#pragma unroll
          for (int i = 0; i < 8; ++i)
            upk2(gelu2(fma2(pk2(__uint_as_float(r[16 * sub + 2 * i]), __uint_as_float(r[16 * sub + 2 * i + 1])), one2, bias2)), y[2 * i],
                 y[2 * i + 1]);
          float lo[8], hi[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) { lo[i] = y[i]; hi[i] = y[8 + i]; }
#if VRCOC_MF_TRACE
          et_ = clock64();
#endif
          if (c == 0 && j >= HB) mbar_wait(&h_empty[hb], (uint32_t)(j / HB - 1) & 1);   // the second GEMM of chunk j-HB has read it
#if VRCOC_MF_TRACE
          eh_ += clock64() - et_;
#endif
          sts128(hbase + (uint32_t)(((2 * c) ^ sw) << 4), pack8_bf16(lo));
          sts128(hbase + (uint32_t)(((2 * c + 1) ^ sw) << 4), pack8_bf16(hi));
        }
      };
      auto release_acc1 = [&]() {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(acc1_empty)) : "memory");
      };
      if (SINGLE) {
        // two 32-column reads; after the second one acc1 is free, so the next chunk's first GEMM runs under the second half of the GELU
        // work (two CTAs per SM: 32 more registers per thread are not available)
#pragma unroll 1
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t r[32];
          tmem_ld32(tbase + (uint32_t)(32 * hh), r);
          if (hh == 1) release_acc1();
          half(r, hh);
        }
      } else {
        // one CTA per SM: both reads first, so acc1 is released BEFORE any GELU work - the per-chunk time stamps (tools/trace_mlpf.py)
        // showed the first-GEMM issuer starting 1.0 us after the accumulator was seen (half of the GELU pass) and its MMAs then
        // sharing the pipe with the second GEMM of the same chunk
        uint32_t ra[32], rb[32];
        tmem_ld32(tbase, ra);
        tmem_ld32(tbase + 32u, rb);
        release_acc1();
        half(ra, 0);
        half(rb, 1);
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&h_full[hb])) : "memory");
#if VRCOC_MF_TRACE
      if (tid == 0 && !SINGLE) MF_STAMP(j, 5);
#endif
    }
#if VRCOC_MF_TRACE
    if (g_tc_trace && tid == 0) {
      unsigned long long* q = g_tc_trace + 4096 * 8 + (blockIdx.z * (size_t)gridDim.x + blockIdx.x) * 8;
      q[4] = ea_; q[5] = eh_; q[6] = clock64() - el0_;
    }
#endif
    // ---- output: + b2, layer scale, residual, statistics, TMA store (staged in the warp's part of the X slab) ----------------------
    if (MT2 == 1)
      cm_epilogue<VRCOC_ACT_NONE, false, 2>(a2, tmem_base + TQ_NP, acc2_full, acc2_empty, &res_bar[warp], out_region, 0, &tmapO, &tmapO,
                                            &tmapR, b, p0, 0, 1);
    else   // output tile m is staged in X slabs 2m, 2m+1 (its own 128 channels); all residual boxes requested at once
      cm_epilogue<VRCOC_ACT_NONE, false, 3, true>(a2, tmem_base + TQ_NP, acc2_full, acc2_empty, &res_bar[warp], out_region, 2 * TQ_X_BYTES,
                                                  &tmapO, &tmapO, &tmapR, b, p0, 0, MT2);
    if (tid == 0) trace(4);
  }
  tc_fence_before();
  __syncthreads();
  if (tid == 0) trace(5);
  if (warp == 9) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)L.tmem_cols));
}

}  // namespace vrcoc
