// Token-mixer half of a ClusterBlock in ONE persistent kernel (inference, bf16, stages 1 and 2 of the backbone):
//
//     out = x + ls1 * ( W2 . cluster_core( W1 . GN(x) + b1 ,  Wv . GN(x) + bv ) + b2 )        (+ GroupNorm statistics of `out`)
//
// reference backbone/fusion/vr_coc.py:155-192 (Cluster.forward) inside :264-267 (ClusterBlock.forward, first half).
//
// As three launches ([GN + fc1|fc_v] -> [cluster core] -> [fc2 + ls + residual]) the fp32 `feat`, the bf16 `value` and the core
// output round-trip HBM: 319 MB moved per stage-1 block at batch 8 for 33.5 MB of algorithmic traffic (x in, out out).  Here one
// CTA owns one 16x16 REGION of one sample (all heads) at a time and nothing but x and out touches HBM:
//
//   TMA    x tile  [C][16 rows][16 cols] bf16  -> smem in the order [row][channel][16 cols] (32-byte rows, SWIZZLE_32B): this IS
//          the MN-major B operand of the first GEMM (N = the region's 256 points, K = channels), later the residual, later the
//          staging tile of the output (rewritten in place, stored by TMA) — the tile is fetched once and never copied.
//   GEMM 1 tcgen05, GroupNorm FOLDED into the weights (include/vrcoc.h, gn_fold_k1): the tensor core contracts the RAW
//          activations with W.diag(gamma) split into two bf16 terms (hi | lo) for `feat` (exact to 2^-17: it decides the hard
//          assignments) and with the hi term for `value`; the per-sample statistics enter when the accumulators are read.
//          TMEM: feat = columns 0-255 (lane = channel e*32+d), value = columns 256-511.
//   core   16 warps; warp (e, j) = head e (its TMEM lane quarter), points 64j..64j+63 (region rows 4j..4j+3):
//            pass 1  quadrant sums of feat straight from TMEM registers (lane = channel) -> centres, normalised per warp;
//            pass 2  32-point chunks of feat transposed through a 4 KB swizzled scratch: a lane owns 2 points x 16 channels,
//                    FFMA2 streams for |f|^2 and the 4 centre dot products, one shuffle step, arg-max, sigmoid gate ->
//                    one-hot weights w[n][m] in smem, member counts;
//            pass 3  value from TMEM (lane = channel): A[m] += w[n][m] * v[n] with broadcast weight reads, quadrant sums;
//                    partials of the four warps of a head combined through smem;
//            pass 4  o[n] = sum_m w[n][m] * a[m], rounded to bf16 and written as the MN-major SW128 B operand of GEMM 2
//                    (the 4 KB chunk a warp writes is the one its scratch lived in).
//   GEMM 2 tcgen05  D2[C][256] = W2[C][128] . o   into the dead feat columns.
//   epilogue        + b2, * ls1, + x (from the tile), statistics for the next GroupNorm, bf16, in place; one TMA store.
//
// One extra warp runs TMA and issues every MMA.  W1 (hi|lo, 48-96 KB) is re-fetched from L2 per region into the operand /
// scratch region while the previous epilogue runs; x tiles are double-buffered at C = 64.
#include <stdlib.h>

#include "conv_common.cuh"
#include "tma.cuh"

namespace vrcoc {
namespace {

constexpr int TM_E = 4, TM_D = 32, TM_ED = 128, TM_RS = 16, TM_N = 256;
constexpr int TM_CWARPS = 16;
constexpr int TM_THREADS = (TM_CWARPS + 1) * 32;
constexpr float TM_EPS = 1e-12f;

// ---- tcgen05 / packed-math helpers (same encodings as conv_tc.cu) ---------------------------------------------------------------
__device__ __forceinline__ void tm_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tm_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tm_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// shared-memory matrix descriptor = {lo: start >> 4 | (LBO >> 4) << 16, hi: SBO >> 4 | version 1 << 14 | layout << 29}
constexpr uint32_t TM_HI_SW128 = (1024u >> 4) | (1u << 14) | (2u << 29);      // SBO 1024 B, SWIZZLE_128B
constexpr uint32_t TM_HI_SW32 = (256u >> 4) | (1u << 14) | (6u << 29);        // SBO 256 B (8 k-rows x 32 B), SWIZZLE_32B
__device__ __forceinline__ uint32_t tm_desc_lo(uint32_t saddr, uint32_t lbo_bytes) {
  return ((saddr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
// D[128 x 256] (+)= A[128 x 16] (K-major, SW128) . B[16 x 256] (MN-major): kind::f16, bf16 operands, fp32 accumulate
constexpr uint32_t TM_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | (0u << 15) | (1u << 16) | ((uint32_t)(TM_N >> 3) << 17) | ((128u >> 4) << 24);
__device__ __forceinline__ void tm_mma(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
      "}" ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(TM_IDESC), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tm_ld32(uint32_t taddr, uint32_t (&r)[32]) {
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
__device__ __forceinline__ void tm_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ uint64_t pk(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
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
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ float hsum(uint64_t v) { float a, b; upk(v, a, b); return a + b; }
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.b32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint32_t bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float bf16_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
__device__ __forceinline__ void named_bar(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct TmArgs {
  const double* gn_sums; float gn_eps;          // per-sample slot sums of x (GroupNorm 1)
  const float* k0; const float* k1;             // [2*ED] folded GroupNorm constants (ops.fold_gn_weights)
  const float* alpha; const float* beta;        // Cluster.sim_alpha / sim_beta
  const float* b2; const float* ls;             // fc2 bias, layer_scale_1 (nullable = 1)
  double* out_sums;                             // nullable: slot sums of out
  uint8_t* idx_out; float* smax_out;            // nullable auxiliary outputs [B][E][H][W]
  int B, H, W, F1, F2, regions;
};

template <int C> struct TmSmem {
  static constexpr int KC = C / 64;                                    // k-slabs of x
  static constexpr int NXB = C == 64 ? 2 : 1;                          // x tile buffers
  static constexpr int XB = C * TM_N * 2;                              // bytes of one x tile
  static constexpr int off_x = 0;
  static constexpr int off_o = off_x + NXB * XB;                       // 64 KB: W1 (feat hi|lo [, value hi]) / scratch / o operand
  static constexpr int off_d = off_o + 65536;                          // 32 KB: W2 (C = 128: value hi of W1 before it)
  static constexpr int off_wq = off_d + 32768;                         // [E][128 point pairs][4 centres][2] fp32
  static constexpr int off_cd = off_wq + 16384;                        // [16 warps][32 d][4 centres][2] fp32 normalised centres (dup)
  static constexpr int off_p1 = off_cd + 16384;                        // [E][4 j][2][32] feat quadrant partial sums
  static constexpr int off_p3 = off_p1 + 4096;                         // [E][4 j][4 m][32] aggregation partials
  static constexpr int off_pq = off_p3 + 8192;                         // [E][4 j][2][32] value quadrant partial sums
  static constexpr int off_cnt = off_pq + 4096;                        // [E][4 j][4] member counts
  static constexpr int off_bar = off_cnt + 256;
  static constexpr int total = off_bar + 256 + 1024;                   // + alignment slack
  static constexpr bool W2_RESIDENT = C == 64;                         // C = 128: region d holds value-W1 until GEMM 1 is done
  static constexpr int W1_FEAT_BYTES = 2 * KC * 16384;
  static constexpr int W1_VAL_BYTES = KC * 16384;
};

template <int C>
__global__ void __launch_bounds__(TM_THREADS, 1)
token_mixer_fused_kernel(TmArgs A, const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmOut,
                         const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2) {
  using S = TmSmem<C>;
  constexpr int KC = S::KC, NXB = S::NXB;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  unsigned char* smem = smem_raw + pad;
  unsigned char* sX = smem + S::off_x;
  unsigned char* sO = smem + S::off_o;
  unsigned char* sD = smem + S::off_d;
  float* wq = reinterpret_cast<float*>(smem + S::off_wq);
  float* part1 = reinterpret_cast<float*>(smem + S::off_p1);
  float* part3 = reinterpret_cast<float*>(smem + S::off_p3);
  float* partq = reinterpret_cast<float*>(smem + S::off_pq);
  int* cntp = reinterpret_cast<int*>(smem + S::off_cnt);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::off_bar);
  uint64_t* x_full = bars;              // [2]
  uint64_t* w1_full = bars + 2;
  uint64_t* w2_full = bars + 3;
  uint64_t* acc1_full = bars + 4;
  uint64_t* o_ready = bars + 5;
  uint64_t* acc2_full = bars + 6;
  uint64_t* epi_done = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int regions_per_sample = A.F1 * A.F2;
  const int n_mine = ((int)blockIdx.x < A.regions) ? (A.regions - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp == TM_CWARPS) {
    if (lane == 0) {
      mbar_init(&x_full[0], 1); mbar_init(&x_full[1], 1);
      mbar_init(w1_full, 1); mbar_init(w2_full, 1);
      mbar_init(acc1_full, 1); mbar_init(acc2_full, 1);
      mbar_init(o_ready, TM_CWARPS); mbar_init(epi_done, TM_CWARPS);
      mbar_fence_init();
      tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmOut); tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmW2);
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tm_fence_before();
  __syncthreads();
  tm_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t T_FEAT = tmem_base, T_VAL = tmem_base + 256u;

  auto region_coords = [&](int rg, int& b, int& row0, int& col0) {
    b = rg / regions_per_sample;
    const int q = rg - b * regions_per_sample;
    row0 = (q / A.F2) * TM_RS;
    col0 = (q % A.F2) * TM_RS;
  };

  if (warp == TM_CWARPS) {
    // ================================ producer: TMA + MMA issue (one thread) ==================================================
    if (lane == 0 && n_mine > 0) {
      auto load_x = [&](int it) {
        int b, row0, col0;
        region_coords((int)blockIdx.x + it * (int)gridDim.x, b, row0, col0);
        const int buf = it % NXB;
        mbar_expect_tx(&x_full[buf], (uint32_t)S::XB);
        tma_load_4d(sX + buf * S::XB, &tmX, col0, 0, row0, b, &x_full[buf]);
      };
      auto load_w1 = [&]() {                 // feat: hi | lo slabs of rows 0-127; value: hi slabs of rows 128-255
        mbar_expect_tx(w1_full, (uint32_t)(S::W1_FEAT_BYTES + S::W1_VAL_BYTES));
        for (int s = 0; s < 2 * KC; ++s) tma_load_2d(sO + s * 16384, &tmW1, s * 64, 0, w1_full);
        unsigned char* vdst = S::W2_RESIDENT ? sO + S::W1_FEAT_BYTES : sD;
        for (int s = 0; s < KC; ++s) tma_load_2d(vdst + s * 16384, &tmW1, s * 64, TM_ED, w1_full);
      };
      auto load_w2 = [&]() {                 // [128 rows][64 k] SW128 per slab; C = 64: the 64 rows twice (both lane halves of TMEM)
        mbar_expect_tx(w2_full, 32768u);
        for (int s = 0; s < 2; ++s) {
          if (C == 64) {
            tma_load_2d(sD + s * 16384, &tmW2, s * 64, 0, w2_full);
            tma_load_2d(sD + s * 16384 + 8192, &tmW2, s * 64, 0, w2_full);
          } else {
            tma_load_2d(sD + s * 16384, &tmW2, s * 64, 0, w2_full);
          }
        }
      };
      load_x(0);
      load_w1();
      if (S::W2_RESIDENT) load_w2();
      if (NXB == 2 && n_mine > 1) load_x(1);
      uint32_t w2_phase = 0;
      for (int it = 0; it < n_mine; ++it) {
        const int buf = it % NXB;
        const uint32_t ph = (uint32_t)it & 1u, xph = (uint32_t)(it / NXB) & 1u;
        unsigned char* xt = sX + buf * S::XB;
        mbar_wait(&x_full[buf], xph);
        mbar_wait(w1_full, ph);
        tm_fence_after();
        // ---- GEMM 1: feat (hi then lo against the same x slabs) -> columns 0-255, value (hi) -> columns 256-511 ----------------
        {
          const uint32_t x_addr = smem_u32(xt);
          for (int s = 0; s < 2 * KC; ++s) {
            const uint32_t w_addr = smem_u32(sO + s * 16384);
            const int kx = s % KC;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              tm_mma(T_FEAT, tm_desc_lo(w_addr + ks * 32, 16), TM_HI_SW128, tm_desc_lo(x_addr + (kx * 64 + ks * 16) * 32, C * 32), TM_HI_SW32,
                     (s > 0 || ks > 0) ? 1u : 0u);
          }
          unsigned char* vsrc = S::W2_RESIDENT ? sO + S::W1_FEAT_BYTES : sD;
          for (int s = 0; s < KC; ++s) {
            const uint32_t w_addr = smem_u32(vsrc + s * 16384);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              tm_mma(T_VAL, tm_desc_lo(w_addr + ks * 32, 16), TM_HI_SW128, tm_desc_lo(x_addr + (s * 64 + ks * 16) * 32, C * 32), TM_HI_SW32,
                     (s > 0 || ks > 0) ? 1u : 0u);
          }
          tm_commit(acc1_full);
        }
        mbar_wait(acc1_full, ph);                                          // W1 consumed: operand region free for the scratch / o
        if (!S::W2_RESIDENT) load_w2();
        // ---- GEMM 2: D2[C x 256] = W2 . o   into the feat columns ----------------------------------------------------------------
        mbar_wait(o_ready, ph);
        if (!S::W2_RESIDENT || it == 0) { mbar_wait(w2_full, w2_phase); w2_phase ^= 1u; }
        tm_fence_after();
        {
          const uint32_t o_addr = smem_u32(sO);
          for (int s = 0; s < 2; ++s) {
            const uint32_t w_addr = smem_u32(sD + s * 16384);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              tm_mma(T_FEAT, tm_desc_lo(w_addr + ks * 32, 16), TM_HI_SW128, tm_desc_lo(o_addr + (s * 4 + ks) * 2048, 16384), TM_HI_SW128,
                     (s > 0 || ks > 0) ? 1u : 0u);
          }
          tm_commit(acc2_full);
        }
        mbar_wait(acc2_full, ph);                                          // o and W2 consumed
        if (it + 1 < n_mine) load_w1();
        // ---- the epilogue has rewritten the x tile in place: store it, then reuse the buffer ------------------------------------
        mbar_wait(epi_done, ph);
        {
          int b, row0, col0;
          region_coords((int)blockIdx.x + it * (int)gridDim.x, b, row0, col0);
          tma_store_4d(&tmOut, xt, col0, 0, row0, b);
          tma_store_commit();
        }
        if (it + NXB < n_mine) {
          tma_store_wait_read();                                            // (NXB = 2: this also covers the other buffer's older store)
          load_x(it + NXB);
        }
      }
      tma_store_wait_all();
    }
    __syncwarp();
  } else {
    // ================================ compute warps ===============================================================================
    const int e = warp & 3, j = warp >> 2;                               // head (TMEM lane quarter), 64-point block
    const int d = lane;
    const uint32_t lane_off = (uint32_t)(32 * e) << 16;
    const float alpha = __ldg(A.alpha), beta = __ldg(A.beta);
    const float k0f = __ldg(A.k0 + 32 * e + d), k1f = __ldg(A.k1 + 32 * e + d);
    const float k0v = __ldg(A.k0 + TM_ED + 32 * e + d), k1v = __ldg(A.k1 + TM_ED + 32 * e + d);
    // epilogue role: C = 128: channel 32e + lane, points 64j..+63;  C = 64: channel 32(e&1) + lane, points 64j + 32(e>>1)..+31
    const int oc = C == 128 ? 32 * e + lane : 32 * (e & 1) + lane;
    const float b2 = __ldg(A.b2 + oc), ls = A.ls ? __ldg(A.ls + oc) : 1.f;
    const uint32_t scratch = smem_u32(sO) + (uint32_t)(j * 16384 + e * 4096);        // == this warp's chunk of the o operand
    const uint32_t cdw = smem_u32(smem + S::off_cd) + (uint32_t)warp * 1024u;
    const uint32_t wq_e = smem_u32(wq) + (uint32_t)e * 4096u;
    float ssum = 0.f, ssq = 0.f;
    int cur_b = -1;
    float mu = 0.f, rstd = 1.f;
    constexpr float inv_q = 1.0f / 64.0f;

    for (int it = 0; it < n_mine; ++it) {
      const int rg = (int)blockIdx.x + it * (int)gridDim.x;
      int b, row0, col0;
      region_coords(rg, b, row0, col0);
      const uint32_t ph = (uint32_t)it & 1u;
      if (b != cur_b) {
        if (A.out_sums && cur_b >= 0) {                                    // flush the statistics of the previous sample
          ssum = warp_sum(ssum); ssq = warp_sum(ssq);
          if (lane == 0) {
            double* dst = A.out_sums + ((int64_t)cur_b * VRCOC_STAT_SLOTS + ((blockIdx.x * TM_CWARPS + warp) & (VRCOC_STAT_SLOTS - 1))) * 2;
            atomicAdd(dst, (double)ssum); atomicAdd(dst + 1, (double)ssq);
          }
          ssum = 0.f; ssq = 0.f;
        }
        gn_mean_rstd(A.gn_sums, b, (double)C * (double)A.H * (double)A.W, A.gn_eps, mu, rstd);
        cur_b = b;
      }
      const float ehf = fmaf(-rstd * mu, k1f, k0f), ehv = fmaf(-rstd * mu, k1v, k0v);   // feat = rstd*acc + ehf, value = rstd*acc + ehv

      mbar_wait(acc1_full, ph);
      tm_fence_after();

      // ---- pass 1: quadrant sums of feat over this warp's 4 region rows (16 columns = one row; cols 0-7 left, 8-15 right) -------
      {
        float sl = 0.f, sr = 0.f;
#pragma unroll 1
        for (int h2 = 0; h2 < 2; ++h2) {
          uint32_t r[32];
          tm_ld32(T_FEAT + lane_off + (uint32_t)(64 * j + 32 * h2), r);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if ((i & 15) < 8) sl += __uint_as_float(r[i]); else sr += __uint_as_float(r[i]);
          }
        }
        // sum of (rstd*acc + ehf) over 32 points per side
        part1[((e * 4 + j) * 2 + 0) * 32 + d] = fmaf(rstd, sl, 32.f * ehf);
        part1[((e * 4 + j) * 2 + 1) * 32 + d] = fmaf(rstd, sr, 32.f * ehf);
      }
      named_bar(1 + e, 128);
      {
        // centres m = 2*(bottom half) + (right half); rows 0-7 = warps j 0,1
        float c[4];
        const float* p = part1 + e * 4 * 2 * 32 + d;
        c[0] = (p[0 * 64] + p[1 * 64]) * inv_q;
        c[1] = (p[0 * 64 + 32] + p[1 * 64 + 32]) * inv_q;
        c[2] = (p[2 * 64] + p[3 * 64]) * inv_q;
        c[3] = (p[2 * 64 + 32] + p[3 * 64 + 32]) * inv_q;
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const float ss = warp_sum(c[m] * c[m]);
          c[m] *= 1.0f / fmaxf(sqrtf(ss), TM_EPS);
        }
        sts128(cdw + (uint32_t)d * 32u, __float_as_uint(c[0]), __float_as_uint(c[0]), __float_as_uint(c[1]), __float_as_uint(c[1]));
        sts128(cdw + (uint32_t)d * 32u + 16u, __float_as_uint(c[2]), __float_as_uint(c[2]), __float_as_uint(c[3]), __float_as_uint(c[3]));
      }
      __syncwarp();

      // ---- pass 2: similarity / arg-max / gate, 32 points per chunk ----------------------------------------------------------------
      int cnt0 = 0, cnt1 = 0, cnt2 = 0, cnt3 = 0;
      {
        const int p = lane & 15, h = lane >> 4;
#pragma unroll 1
        for (int q = 0; q < 2; ++q) {
          {
            uint32_t r[32];
            tm_ld32(T_FEAT + lane_off + (uint32_t)(64 * j + 32 * q), r);
            // row d of the scratch: 32 points fp32 = 128 B, 16-byte chunks XOR-swizzled by (d & 7)
#pragma unroll
            for (int c16 = 0; c16 < 8; ++c16)
              sts128(scratch + (uint32_t)d * 128u + (uint32_t)(((c16 ^ (d & 7))) << 4),
                     __float_as_uint(fmaf(rstd, __uint_as_float(r[4 * c16]), ehf)), __float_as_uint(fmaf(rstd, __uint_as_float(r[4 * c16 + 1]), ehf)),
                     __float_as_uint(fmaf(rstd, __uint_as_float(r[4 * c16 + 2]), ehf)), __float_as_uint(fmaf(rstd, __uint_as_float(r[4 * c16 + 3]), ehf)));
          }
          __syncwarp();
          uint64_t ss = 0ull, d0 = 0ull, d1 = 0ull, d2 = 0ull, d3 = 0ull;
#pragma unroll 8
          for (int i = 0; i < 16; ++i) {
            const int dd = 16 * h + i;
            const uint2 xr = lds64(scratch + (uint32_t)dd * 128u + (uint32_t)((((p >> 1) ^ (dd & 7)) << 4) + (p & 1) * 8));
            const uint64_t x = ((uint64_t)xr.y << 32) | xr.x;
            const uint4 c01 = lds128(cdw + (uint32_t)dd * 32u), c23 = lds128(cdw + (uint32_t)dd * 32u + 16u);
            ss = fma2(x, x, ss);
            d0 = fma2(((uint64_t)c01.y << 32) | c01.x, x, d0);
            d1 = fma2(((uint64_t)c01.w << 32) | c01.z, x, d1);
            d2 = fma2(((uint64_t)c23.y << 32) | c23.x, x, d2);
            d3 = fma2(((uint64_t)c23.w << 32) | c23.z, x, d3);
          }
          float ssv[2], dv[4][2];
          upk(ss, ssv[0], ssv[1]);
          upk(d0, dv[0][0], dv[0][1]); upk(d1, dv[1][0], dv[1][1]); upk(d2, dv[2][0], dv[2][1]); upk(d3, dv[3][0], dv[3][1]);
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            ssv[t] += __shfl_xor_sync(0xffffffffu, ssv[t], 16);
#pragma unroll
            for (int m = 0; m < 4; ++m) dv[m][t] += __shfl_xor_sync(0xffffffffu, dv[m][t], 16);
          }
          int kb[2];
          float gv[2];
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            const float inv = fminf(rsqrtf(ssv[t]), 1.0f / TM_EPS);
            float tb = alpha * dv[0][t], db = dv[0][t];
            int k = 0;
            if (alpha * dv[1][t] > tb) { tb = alpha * dv[1][t]; db = dv[1][t]; k = 1; }
            if (alpha * dv[2][t] > tb) { tb = alpha * dv[2][t]; db = dv[2][t]; k = 2; }
            if (alpha * dv[3][t] > tb) { tb = alpha * dv[3][t]; db = dv[3][t]; k = 3; }
            kb[t] = k;
            gv[t] = __fdividef(1.0f, 1.0f + __expf(-fmaf(alpha, db * inv, beta)));
          }
          // one-hot weights of the point pair: [pair][m][2]; half h stores centres 2h, 2h+1
          const int pi = (64 * j + 32 * q) / 2 + p;
          const int m0 = 2 * h;
          sts128(wq_e + (uint32_t)pi * 32u + (uint32_t)h * 16u, __float_as_uint(kb[0] == m0 ? gv[0] : 0.f), __float_as_uint(kb[1] == m0 ? gv[1] : 0.f),
                 __float_as_uint(kb[0] == m0 + 1 ? gv[0] : 0.f), __float_as_uint(kb[1] == m0 + 1 ? gv[1] : 0.f));
          if (h == 0) {
            if (A.idx_out || A.smax_out) {
              const int n = 64 * j + 32 * q + 2 * p;
              const int64_t io = ((int64_t)(b * TM_E + e) * A.H + row0 + (n >> 4)) * A.W + col0 + (n & 15);
              if (A.idx_out) *reinterpret_cast<uint16_t*>(A.idx_out + io) = (uint16_t)(kb[0] | (kb[1] << 8));
              if (A.smax_out) *reinterpret_cast<float2*>(A.smax_out + io) = make_float2(gv[0], gv[1]);
            }
          }
          const unsigned half = 0x0000ffffu;
          cnt0 += __popc(__ballot_sync(0xffffffffu, kb[0] == 0) & half) + __popc(__ballot_sync(0xffffffffu, kb[1] == 0) & half);
          cnt1 += __popc(__ballot_sync(0xffffffffu, kb[0] == 1) & half) + __popc(__ballot_sync(0xffffffffu, kb[1] == 1) & half);
          cnt2 += __popc(__ballot_sync(0xffffffffu, kb[0] == 2) & half) + __popc(__ballot_sync(0xffffffffu, kb[1] == 2) & half);
          cnt3 += __popc(__ballot_sync(0xffffffffu, kb[0] == 3) & half) + __popc(__ballot_sync(0xffffffffu, kb[1] == 3) & half);
          __syncwarp();                                                    // the scratch is rewritten by the next chunk
        }
        if (lane == 0) *reinterpret_cast<int4*>(cntp + (e * 4 + j) * 4) = make_int4(cnt0, cnt1, cnt2, cnt3);
      }
      named_bar(1 + e, 128);                                               // weights + counts of the head complete

      // ---- pass 3: aggregate value (TMEM, lane = channel) to the centres --------------------------------------------------------------
      {
        uint64_t A2[4] = {0ull, 0ull, 0ull, 0ull};
        uint64_t ql = 0ull, qr = 0ull;
#pragma unroll 1
        for (int h2 = 0; h2 < 2; ++h2) {
          uint32_t r[32];
          tm_ld32(T_VAL + lane_off + (uint32_t)(64 * j + 32 * h2), r);
          const uint32_t wbase = wq_e + (uint32_t)((64 * j + 32 * h2) / 2) * 32u;
#pragma unroll
          for (int pp = 0; pp < 16; ++pp) {
            const uint4 w01 = lds128(wbase + (uint32_t)pp * 32u), w23 = lds128(wbase + (uint32_t)pp * 32u + 16u);
            // value is a bf16 tensor in the reference pipeline (stored, then read by the core): same rounding here
            const uint64_t v2 = pk(bf16_round(fmaf(rstd, __uint_as_float(r[2 * pp]), ehv)), bf16_round(fmaf(rstd, __uint_as_float(r[2 * pp + 1]), ehv)));
            A2[0] = fma2(((uint64_t)w01.y << 32) | w01.x, v2, A2[0]);
            A2[1] = fma2(((uint64_t)w01.w << 32) | w01.z, v2, A2[1]);
            A2[2] = fma2(((uint64_t)w23.y << 32) | w23.x, v2, A2[2]);
            A2[3] = fma2(((uint64_t)w23.w << 32) | w23.z, v2, A2[3]);
            if ((pp & 7) < 4) ql = add2(ql, v2); else qr = add2(qr, v2);
          }
        }
        float* p3 = part3 + (e * 4 + j) * 4 * 32 + d;
        p3[0] = hsum(A2[0]); p3[32] = hsum(A2[1]); p3[64] = hsum(A2[2]); p3[96] = hsum(A2[3]);
        float* pq = partq + (e * 4 + j) * 2 * 32 + d;
        pq[0] = hsum(ql); pq[32] = hsum(qr);
      }
      named_bar(1 + e, 128);

      // ---- pass 4: centre aggregates, dispatch, bf16, MN-major SW128 operand of GEMM 2 --------------------------------------------------
      {
        float a[4];
        {
          int cn[4] = {0, 0, 0, 0};
          float As[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const int4 c4 = *reinterpret_cast<const int4*>(cntp + (e * 4 + jj) * 4);
            cn[0] += c4.x; cn[1] += c4.y; cn[2] += c4.z; cn[3] += c4.w;
            const float* p3 = part3 + (e * 4 + jj) * 4 * 32 + d;
            As[0] += p3[0]; As[1] += p3[32]; As[2] += p3[64]; As[3] += p3[96];
          }
          const float* pq = partq + e * 4 * 2 * 32 + d;
          const float Q0 = pq[0 * 64] + pq[1 * 64], Q1 = pq[0 * 64 + 32] + pq[1 * 64 + 32];
          const float Q2 = pq[2 * 64] + pq[3 * 64], Q3 = pq[2 * 64 + 32] + pq[3 * 64 + 32];
          a[0] = fmaf(Q0, inv_q, As[0]) * __fdividef(1.0f, (float)(cn[0] + 1));
          a[1] = fmaf(Q1, inv_q, As[1]) * __fdividef(1.0f, (float)(cn[1] + 1));
          a[2] = fmaf(Q2, inv_q, As[2]) * __fdividef(1.0f, (float)(cn[2] + 1));
          a[3] = fmaf(Q3, inv_q, As[3]) * __fdividef(1.0f, (float)(cn[3] + 1));
        }
        const uint64_t a0 = pk(a[0], a[0]), a1 = pk(a[1], a[1]), a2 = pk(a[2], a[2]), a3 = pk(a[3], a[3]);
        // this lane's k-row of the o operand: n-block j, row 32e + d -> (4e + d/8) * 1024 + (d % 8) * 128, chunks XOR (d & 7)
        const uint32_t orow = smem_u32(sO) + (uint32_t)(j * 16384 + (4 * e + (d >> 3)) * 1024 + (d & 7) * 128);
#pragma unroll 1
        for (int h2 = 0; h2 < 2; ++h2) {
          const uint32_t wbase = wq_e + (uint32_t)((64 * j + 32 * h2) / 2) * 32u;
          uint32_t ow[16];
#pragma unroll
          for (int pp = 0; pp < 16; ++pp) {
            const uint4 w01 = lds128(wbase + (uint32_t)pp * 32u), w23 = lds128(wbase + (uint32_t)pp * 32u + 16u);
            const uint64_t o2 = fma2(((uint64_t)w01.y << 32) | w01.x, a0,
                                     fma2(((uint64_t)w01.w << 32) | w01.z, a1, fma2(((uint64_t)w23.y << 32) | w23.x, a2, mul2(((uint64_t)w23.w << 32) | w23.z, a3))));
            float lo, hi;
            upk(o2, lo, hi);
            ow[pp] = bf16x2(lo, hi);
          }
#pragma unroll
          for (int t = 0; t < 4; ++t)
            sts128(orow + (uint32_t)((((4 * h2 + t) ^ (d & 7))) << 4), ow[4 * t], ow[4 * t + 1], ow[4 * t + 2], ow[4 * t + 3]);
        }
      }
      fence_async_smem();
      tm_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_ready);

      // ---- epilogue: + b2, * ls, + x (the tile itself), statistics, bf16 in place ------------------------------------------------------------
      mbar_wait(acc2_full, ph);
      tm_fence_after();
      {
        const uint32_t xt = smem_u32(sX + (it % NXB) * S::XB);
        constexpr int ROWS = C == 128 ? 4 : 2;                            // region rows (16 points each) this warp finishes
        const int n_first = C == 128 ? 64 * j : 64 * j + 32 * (e >> 1);
        const uint32_t sw = (uint32_t)((oc >> 2) & 1);
#pragma unroll 1
        for (int rr = 0; rr < ROWS; ++rr) {
          const int n0 = n_first + 16 * rr;
          uint32_t r[16];
          tm_ld16(T_FEAT + lane_off + (uint32_t)n0, r);
          const uint32_t rowaddr = xt + (uint32_t)((n0 >> 4) * C * 32 + oc * 32);
          const uint4 x0 = lds128(rowaddr + ((0u ^ sw) << 4)), x1 = lds128(rowaddr + ((1u ^ sw) << 4));
          const uint32_t xw[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
          uint32_t yw[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float r0 = __uint_as_float(xw[i] << 16), r1 = __uint_as_float(xw[i] & 0xffff0000u);
            const float y0 = fmaf(__uint_as_float(r[2 * i]) + b2, ls, r0), y1 = fmaf(__uint_as_float(r[2 * i + 1]) + b2, ls, r1);
            ssum += y0 + y1;
            ssq = fmaf(y0, y0, fmaf(y1, y1, ssq));
            yw[i] = bf16x2(y0, y1);
          }
          sts128(rowaddr + ((0u ^ sw) << 4), yw[0], yw[1], yw[2], yw[3]);
          sts128(rowaddr + ((1u ^ sw) << 4), yw[4], yw[5], yw[6], yw[7]);
        }
      }
      fence_async_smem();
      tm_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(epi_done);
    }
    if (A.out_sums && cur_b >= 0) {
      ssum = warp_sum(ssum); ssq = warp_sum(ssq);
      if (lane == 0) {
        double* dst = A.out_sums + ((int64_t)cur_b * VRCOC_STAT_SLOTS + ((blockIdx.x * TM_CWARPS + warp) & (VRCOC_STAT_SLOTS - 1))) * 2;
        atomicAdd(dst, (double)ssum); atomicAdd(dst + 1, (double)ssq);
      }
    }
  }
  tm_fence_before();
  __syncthreads();
  if (warp == TM_CWARPS) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
}

template <int C>
int launch_tm(const TmArgs& A, const CUtensorMap& tx, const CUtensorMap& to, const CUtensorMap& tw1, const CUtensorMap& tw2, cudaStream_t st) {
  using S = TmSmem<C>;
  auto kern = token_mixer_fused_kernel<C>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::total);
  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  int grid = sm_count();
  if (grid > A.regions) grid = A.regions;
  kern<<<grid, TM_THREADS, S::total, st>>>(A, tx, to, tw1, tw2);
  return check_launch("token_mixer_fused");
}

}  // namespace
}  // namespace vrcoc

using namespace vrcoc;

extern "C" int vrcoc_token_mixer_supported(int dtype, int C, int H, int W, int heads, int head_dim, int fold_w, int fold_h,
                                           int proposal_w, int proposal_h) {
  const char* knob = getenv("VRCOC_TM_FUSED");                          // "0": A/B switch (tools/microbench.py, parity tests)
  if (knob && knob[0] == '0') return 0;
  const bool folded = fold_w > 1 && fold_h > 1;
  const int f1 = folded ? fold_w : 1, f2 = folded ? fold_h : 1;
  return dtype == VRCOC_BF16 && (C == 64 || C == 128) && heads == TM_E && head_dim == TM_D && proposal_w == 2 && proposal_h == 2 &&
         H % f1 == 0 && W % f2 == 0 && H / f1 == TM_RS && W / f2 == TM_RS && tma_encode_fn() != nullptr;
}

extern "C" int vrcoc_token_mixer_fwd(const void* x, const double* gn_sums, float gn_eps, const void* w_fold, const float* k0, const float* k1,
                                     const float* alpha, const float* beta, const void* w2, const float* b2, const float* layer_scale,
                                     void* out, double* out_sample_sums, uint8_t* idx, float* sim_max, int B, int C, int H, int W,
                                     int heads, int head_dim, int fold_w, int fold_h, void* stream) {
  VRCOC_REQUIRE(x && gn_sums && w_fold && k0 && k1 && alpha && beta && w2 && b2 && out && B > 0, "token_mixer: null pointer / empty batch");
  VRCOC_REQUIRE(vrcoc_token_mixer_supported(VRCOC_BF16, C, H, W, heads, head_dim, fold_w, fold_h, 2, 2),
                "token_mixer: unsupported geometry C=%d %dx%d heads=%d head_dim=%d fold=%dx%d (needs C in {64,128}, 4 heads x 32, 16x16 regions)",
                C, H, W, heads, head_dim, fold_w, fold_h);
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  VRCOC_REQUIRE(al(x) && al(out) && al(w_fold) && al(w2), "token_mixer: tensors must be 16-byte aligned");
  VRCOC_REQUIRE(!idx || (reinterpret_cast<uintptr_t>(idx) & 1) == 0, "token_mixer: idx must be 2-byte aligned");
  VRCOC_REQUIRE(!sim_max || (reinterpret_cast<uintptr_t>(sim_max) & 7) == 0, "token_mixer: sim_max must be 8-byte aligned");
  TmArgs A;
  A.gn_sums = gn_sums; A.gn_eps = gn_eps; A.k0 = k0; A.k1 = k1; A.alpha = alpha; A.beta = beta; A.b2 = b2; A.ls = layer_scale;
  A.out_sums = out_sample_sums; A.idx_out = idx; A.smax_out = sim_max;
  A.B = B; A.H = H; A.W = W; A.F1 = H / TM_RS; A.F2 = W / TM_RS; A.regions = B * A.F1 * A.F2;
  CUtensorMap tx, to, tw1, tw2;
  int rc;
  {
    // x / out [B][C][H][W] viewed as (col, channel, row, sample): a box of 16 cols x C channels x 16 rows lands as [row][channel][32 B]
    cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)C, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)H * W * 2, (cuuint64_t)W * 2, (cuuint64_t)C * H * W * 2};
    cuuint32_t box[4] = {(cuuint32_t)TM_RS, (cuuint32_t)C, (cuuint32_t)TM_RS, 1};
    if ((rc = tma_encode_sw(&tx, VRCOC_BF16, x, 4, dims, strides, box, 32))) return rc;
    if ((rc = tma_encode_sw(&to, VRCOC_BF16, out, 4, dims, strides, box, 32))) return rc;
  }
  {
    cuuint64_t d1[2] = {(cuuint64_t)(2 * C), (cuuint64_t)(2 * TM_ED)}, s1[1] = {(cuuint64_t)(2 * C) * 2};
    cuuint32_t b1[2] = {64, 128};
    if ((rc = tma_encode_sw(&tw1, VRCOC_BF16, w_fold, 2, d1, s1, b1, 128))) return rc;
    cuuint64_t d2[2] = {(cuuint64_t)TM_ED, (cuuint64_t)C}, s2[1] = {(cuuint64_t)TM_ED * 2};
    cuuint32_t bx2[2] = {64, (cuuint32_t)(C == 64 ? 64 : 128)};
    if ((rc = tma_encode_sw(&tw2, VRCOC_BF16, w2, 2, d2, s2, bx2, 128))) return rc;
  }
  if (C == 64) return launch_tm<64>(A, tx, to, tw1, tw2, (cudaStream_t)stream);
  return launch_tm<128>(A, tx, to, tw1, tw2, (cudaStream_t)stream);
}
