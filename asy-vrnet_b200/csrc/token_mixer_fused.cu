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
//            pass 2  32-point chunks of feat transposed through a 4 KB swizzled scratch: a lane then owns 4 channels x 8 points with
//                    its centres in registers (no broadcast loads), FFMA2 streams for |f|^2 and the 4 centre dot products, a
//                    transpose-reduce over 8 lanes leaves lane L with point L: arg-max, sigmoid gate -> (gate | centre) word;
//            pass 3  value from TMEM (lane = channel): A[k_n] += g_n * v[n] with one 16-byte broadcast read per 4 points;
//                    partials of the four warps of a head combined through smem;
//            pass 4  lane = point: o[n][:] = g_n * a[k_n][:], rounded to bf16 and written as the K-major SW128 B operand of GEMM 2
//                    (row = point; the first version kept lane = channel everywhere and was bound by the shared-memory pipe:
//                    broadcast LDS.128 of one-hot weights, 4 cycles per point and warp, 12.5 us per region).
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
constexpr int TM_THREADS = (TM_CWARPS + 4) * 32;                     // + one warpgroup: warp 16 = TMA / MMA issue, warps 17-19 idle
constexpr int TM_REGS_COMPUTE = 104, TM_REGS_AUX = 56;               // setmaxnreg re-divides the CTA's OWN launch pool (640 x 96 = 61440 registers):
                                                                     // 16*32*104 + 4*32*56 = 60416 fits (120 would dead-lock the last warpgroup; at 24 the MMA-issuing thread spills and crawls)
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
constexpr uint32_t TM_IDESC_KK = TM_IDESC & ~(1u << 16);             // the same with a K-major B operand (GEMM 2: rows = points)
// aggregation GEMM (pass 3): D[128 x 48] = V[128 x 256 points] (bf16 in TMEM) . WB[256 points x 48] (K-major smem)
constexpr int TM_NW = 48;                                            // 16 (gate hi) + 16 (gate lo) + 4 quadrant means + 12 zero rows
constexpr uint32_t TM_IDESC_AGG = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TM_NW >> 3) << 17) | ((128u >> 4) << 24);
constexpr int TM_WB_SLAB = TM_NW * 128;                              // bytes of one 64-point k-slab of WB
template <uint32_t IDESC = TM_IDESC>
__device__ __forceinline__ void tm_mma(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
      "}" ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(IDESC), "r"(accumulate)
      : "memory");
}
// D[128 x N] (+)= A[128 x 16] (bf16 pairs in TMEM, lane = row, 8 columns) . B[16 x N] (smem, K-major SW128): pass 3 on the tensor core
__device__ __forceinline__ void tm_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 db;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t"
      "}" ::"r"(tmem_d), "r"(tmem_a), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tm_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,"
      "%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
      "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]),
      "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tm_ld4_issue(uint32_t taddr, uint32_t (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
}
// the destination registers of the split loads are only defined after the wait: they are in/out operands of it so that the
// compiler orders every use after it
__device__ __forceinline__ void tm_ld_wait12(uint32_t (&a)[4], uint32_t (&b)[4], uint32_t (&c)[4]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(b[0]), "+r"(b[1]), "+r"(b[2]), "+r"(b[3]), "+r"(c[0]), "+r"(c[1]),
                 "+r"(c[2]), "+r"(c[3])
               :
               : "memory");
}
__device__ __forceinline__ void sts16(uint32_t addr, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((unsigned short)v) : "memory"); }
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

// debug trace (vrcoc_debug_set_tm_trace): per CTA and region iteration 16 u64 %globaltimer stamps
//   producer: 0 x+W1 landed, 1 GEMM 1 issued, 2 GEMM 1 done, 3 o ready seen, 4 GEMM 2 done, 5 epilogue done seen, 6 store issued
//   warp 0:   8 GEMM 1 seen, 9 pass 1, 10 pass 2, 11 pass 3, 12 pass 4, 13 GEMM 2 seen, 14 epilogue written
__device__ unsigned long long* g_tm_trace = nullptr;
constexpr int TM_TRACE_ITERS = 4;
__device__ __forceinline__ unsigned long long tm_time() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void tm_trace(unsigned long long* tr, int it, int slot) {
  if (tr && it < TM_TRACE_ITERS) tr[((size_t)blockIdx.x * TM_TRACE_ITERS + it) * 16 + slot] = tm_time();
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

// Everything between the projection GEMM and the dispatch, shared by the two kernels below: pass 1 (centres), pass 2 (similarity,
// arg-max, gate, one-hot rows of the aggregation GEMM), value -> bf16 TMEM operand, wait for the aggregation GEMM, centre aggregates
// into the warp's private table.  Leaves gk[q] = (gate bits & ~3) | centre of this lane's point in chunk q.
struct TmCore {
  uint32_t T_FEAT, T_VAL, lane_off, scratch, wb_base, ph;
  int e, j, lane, it, Wimg;
  float* part1; int* cntp; float* apriv;
  uint64_t* va_ready; uint64_t* accA_full;
  float rstd, ehf, ehv, alpha, beta;
  uint8_t* idx_out; float* smax_out; int64_t io_base;
  unsigned long long* tr;
};

__device__ __forceinline__ void tm_core_passes(const TmCore& c, uint32_t (&gk)[2]) {
  const uint32_t T_FEAT = c.T_FEAT, T_VAL = c.T_VAL, lane_off = c.lane_off, scratch = c.scratch, wb_base = c.wb_base;
  const int e = c.e, j = c.j, lane = c.lane, d = c.lane;
  const int dq = lane & 3, pg = lane >> 2;                             // pass-2 read mapping: channels 8dq..8dq+7, points 4pg..4pg+3
  float* part1 = c.part1; int* cntp = c.cntp; float* apriv = c.apriv;
  const float rstd = c.rstd, ehf = c.ehf, ehv = c.ehv, alpha = c.alpha, beta = c.beta;
  constexpr float inv_q = 1.0f / 64.0f;
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
      // normalised centres of THIS LANE's eight channels 8*dq .. 8*dq+7 (pass-2 mapping), kept in registers:
      // cc[t][m], m = 2*(bottom half) + (right half); rows 0-7 = warps j 0,1
      float cc[8][4];
      {
        const uint32_t p1 = smem_u32(part1) + (uint32_t)((e * 4 * 2 * 32 + 8 * dq) * 4);
        float ssm[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {                                   // channels 8dq + 4hf .. + 3
          float q[4][2][4];                                                // [j][side][channel]
#pragma unroll
          for (int jj = 0; jj < 4; ++jj)
#pragma unroll
            for (int sd = 0; sd < 2; ++sd) {
              const uint4 v = lds128(p1 + (uint32_t)(((jj * 2 + sd) * 32 + 4 * hf) * 4));
              q[jj][sd][0] = __uint_as_float(v.x); q[jj][sd][1] = __uint_as_float(v.y);
              q[jj][sd][2] = __uint_as_float(v.z); q[jj][sd][3] = __uint_as_float(v.w);
            }
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            cc[4 * hf + t][0] = (q[0][0][t] + q[1][0][t]) * inv_q; cc[4 * hf + t][1] = (q[0][1][t] + q[1][1][t]) * inv_q;
            cc[4 * hf + t][2] = (q[2][0][t] + q[3][0][t]) * inv_q; cc[4 * hf + t][3] = (q[2][1][t] + q[3][1][t]) * inv_q;
#pragma unroll
            for (int m = 0; m < 4; ++m) ssm[m] = fmaf(cc[4 * hf + t][m], cc[4 * hf + t][m], ssm[m]);
          }
        }
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          ssm[m] += __shfl_xor_sync(0xffffffffu, ssm[m], 1);
          ssm[m] += __shfl_xor_sync(0xffffffffu, ssm[m], 2);
          const float inv = fminf(rsqrtf(ssm[m]), 1.0f / TM_EPS);         // 1 / max(|c|, eps) to 2 ulp
#pragma unroll
          for (int t = 0; t < 8; ++t) cc[t][m] *= inv;
        }
      }
      // zero this warp's part of the aggregation GEMM's B operand (rows of head e, its 64 points): pass 2 then only writes the
      // two non-zero entries of each point
      {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int u = lane + 32 * i;                                     // 8 rows (4 hi + 4 lo) x 8 chunks of 16 B
          const int r = (u >> 3) < 4 ? 4 * e + (u >> 3) : 16 + 4 * e + (u >> 3) - 4;
          sts128(wb_base + (uint32_t)(j * TM_WB_SLAB + (r >> 3) * 1024 + (r & 7) * 128 + (u & 7) * 16), 0u, 0u, 0u, 0u);
        }
      }
      __syncwarp();
      tm_trace(c.tr, c.it, 9);

      // ---- pass 2: similarity / arg-max / gate, 32 points per chunk ----------------------------------------------------------------
      // The chunk is transposed through the scratch: written lane = channel (TMEM order), read lane = (dq, pg): 4 channels x 8 points,
      // centres from registers -> no broadcast loads; 40 partial sums are then transpose-reduced over the 8 dq lanes so that lane L
      // ends up with the complete |f|^2 and 4 dot products of point L of the chunk.
      int cnt0 = 0, cnt1 = 0, cnt2 = 0, cnt3 = 0;
      {
        const uint32_t swzw = (uint32_t)((((d >> 3) << 1) ^ (d & 7)) & 7);   // conflict-free for both access patterns
#pragma unroll 1
        for (int q = 0; q < 2; ++q) {
          {
            uint32_t r[32];
            tm_ld32(T_FEAT + lane_off + (uint32_t)(64 * j + 32 * q), r);
#pragma unroll
            for (int c16 = 0; c16 < 8; ++c16)
              sts128(scratch + (uint32_t)d * 128u + (((uint32_t)c16 ^ swzw) << 4),
                     __float_as_uint(fmaf(rstd, __uint_as_float(r[4 * c16]), ehf)), __float_as_uint(fmaf(rstd, __uint_as_float(r[4 * c16 + 1]), ehf)),
                     __float_as_uint(fmaf(rstd, __uint_as_float(r[4 * c16 + 2]), ehf)), __float_as_uint(fmaf(rstd, __uint_as_float(r[4 * c16 + 3]), ehf)));
          }
          __syncwarp();
          uint64_t acc2[2][5];
#pragma unroll
          for (int pr = 0; pr < 2; ++pr)
#pragma unroll
            for (int v = 0; v < 5; ++v) acc2[pr][v] = 0ull;
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            const int dd = 8 * dq + t;
            const uint32_t swzr = (uint32_t)(((2 * dq) ^ t) & 7);
            const uint4 f = lds128(scratch + (uint32_t)dd * 128u + ((((uint32_t)pg) ^ swzr) << 4));     // points 4pg .. 4pg+3 of channel dd
            const uint64_t x2[2] = {((uint64_t)f.y << 32) | f.x, ((uint64_t)f.w << 32) | f.z};
            const uint64_t c0 = pk(cc[t][0], cc[t][0]), c1 = pk(cc[t][1], cc[t][1]), c2 = pk(cc[t][2], cc[t][2]), c3 = pk(cc[t][3], cc[t][3]);
#pragma unroll
            for (int pr = 0; pr < 2; ++pr) {
              acc2[pr][0] = fma2(x2[pr], x2[pr], acc2[pr][0]);
              acc2[pr][1] = fma2(c0, x2[pr], acc2[pr][1]);
              acc2[pr][2] = fma2(c1, x2[pr], acc2[pr][2]);
              acc2[pr][3] = fma2(c2, x2[pr], acc2[pr][3]);
              acc2[pr][4] = fma2(c3, x2[pr], acc2[pr][4]);
            }
          }
          // transpose-reduce over the 4 dq lanes: bit 1 halves the 4 points to 2, bit 0 to 1 -> lane L holds point L of the chunk
          float a4[4][5];
#pragma unroll
          for (int pr = 0; pr < 2; ++pr)
#pragma unroll
            for (int v = 0; v < 5; ++v) upk(acc2[pr][v], a4[2 * pr][v], a4[2 * pr + 1][v]);
          float a2[2][5], a1[5];
          const bool b1_ = (dq & 2) != 0, b0_ = (dq & 1) != 0;
#pragma unroll
          for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int v = 0; v < 5; ++v) {
              const float keep = b1_ ? a4[i + 2][v] : a4[i][v], send = b1_ ? a4[i][v] : a4[i + 2][v];
              a2[i][v] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
            }
#pragma unroll
          for (int v = 0; v < 5; ++v) {
            const float keep = b0_ ? a2[1][v] : a2[0][v], send = b0_ ? a2[0][v] : a2[1][v];
            a1[v] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
          }
          // this lane's point: n = 64j + 32q + lane
          const float inv = fminf(rsqrtf(a1[0]), 1.0f / TM_EPS);
          float tb = alpha * a1[1], db = a1[1];
          int k = 0;
          if (alpha * a1[2] > tb) { tb = alpha * a1[2]; db = a1[2]; k = 1; }
          if (alpha * a1[3] > tb) { tb = alpha * a1[3]; db = a1[3]; k = 2; }
          if (alpha * a1[4] > tb) { tb = alpha * a1[4]; db = a1[4]; k = 3; }
          const float gv = __fdividef(1.0f, 1.0f + __expf(-fmaf(alpha, db * inv, beta)));
          // the centre index rides in the two lowest mantissa bits of the gate (2^-22 relative); passes 3 and 4 both use the masked gate
          gk[q] = (__float_as_uint(gv) & ~3u) | (uint32_t)k;
          const int n = 64 * j + 32 * q + lane;
          {
            // B operand of the aggregation GEMM, K-major SW128 [k-slab = n / 64][row][128 B]: rows 4e+m = bf16 hi part of the gate
            // where m is this point's centre (0 elsewhere), rows 16+4e+m = the lo part (hi + lo carries the gate to 2^-17)
            const float gm = __uint_as_float(gk[q] & ~3u);
            const __nv_bfloat16 hb = __float2bfloat16_rn(gm);
            const uint32_t hi16 = (uint32_t)__bfloat16_as_ushort(hb);
            const uint32_t lo16 = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(gm - __bfloat162float(hb)));
            const uint32_t colb = wb_base + (uint32_t)((n >> 6) * TM_WB_SLAB + (n & 7) * 2);
            const uint32_t ch = (uint32_t)((n & 63) >> 3);
            const int r1 = 4 * e + k, r2 = 16 + 4 * e + k;                 // the other rows of this point were zeroed above
            sts16(colb + (uint32_t)((r1 >> 3) * 1024 + (r1 & 7) * 128) + ((ch ^ (uint32_t)(r1 & 7)) << 4), hi16);
            sts16(colb + (uint32_t)((r2 >> 3) * 1024 + (r2 & 7) * 128) + ((ch ^ (uint32_t)(r2 & 7)) << 4), lo16);
          }
          if (c.idx_out || c.smax_out) {
            const int64_t io = c.io_base + (int64_t)(n >> 4) * c.Wimg + (n & 15);
            if (c.idx_out) c.idx_out[io] = (uint8_t)k;
            if (c.smax_out) c.smax_out[io] = __uint_as_float(gk[q] & ~3u);
          }
          cnt0 += __popc(__ballot_sync(0xffffffffu, k == 0));
          cnt1 += __popc(__ballot_sync(0xffffffffu, k == 1));
          cnt2 += __popc(__ballot_sync(0xffffffffu, k == 2));
          cnt3 += __popc(__ballot_sync(0xffffffffu, k == 3));
          __syncwarp();                                                    // the scratch is rewritten by the next chunk
        }
        if (lane == 0) *reinterpret_cast<int4*>(cntp + (e * 4 + j) * 4) = make_int4(cnt0, cnt1, cnt2, cnt3);
      }
      // the scratch of the two heads of a k-slab lives in that slab of the o operand: both heads must have left pass 2 before
      // either writes o (pass 4)
      named_bar(5 + (e >> 1), 256);
      tm_trace(c.tr, c.it, 10);

      // ---- pass 3 on the tensor core: value -> bf16 pairs in the (dead) feat columns of TMEM = the A operand of the aggregation GEMM;
      //      the one-hot gate rows written in pass 2 are its B operand.  (The CUDA-core forms of this pass were the hot spot of the
      //      first two versions: 4 shared-memory cycles per point and warp for broadcast one-hot weights, or 19 instructions per point
      //      for predicated accumulation from packed (gate | centre) words.) -----------------------------------------------------------
      {
        uint32_t pw[32];
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          uint32_t r[32];
          tm_ld32(T_VAL + lane_off + (uint32_t)(64 * j + 32 * h2), r);
#pragma unroll
          for (int i = 0; i < 16; ++i)
            pw[16 * h2 + i] = bf16x2(fmaf(rstd, __uint_as_float(r[2 * i]), ehv), fmaf(rstd, __uint_as_float(r[2 * i + 1]), ehv));
        }
        tm_st32(T_FEAT + lane_off + (uint32_t)(32 * j), pw);               // points 64j + 2c, 64j + 2c + 1 -> column 32j + c
      }
      fence_async_smem();                                                  // the gate rows (generic-proxy stores) before the MMA reads them
      tm_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(c.va_ready);
      tm_trace(c.tr, c.it, 11);
      // ---- centre aggregates (lane = channel) from the aggregation GEMM -> this warp's private table a[m][d] -----------------------------
      {
        mbar_wait(c.accA_full, c.ph);
        tm_fence_after();
        tm_trace(c.tr, c.it, 7);
        {
          // D_A columns (from T_FEAT + 128): [4e, 4e+4) = sum of hi-gated value, [16 + 4e, ..) = lo part, [32, 36) = quadrant means
          uint32_t dh[4], dl[4], dqm[4];
          tm_ld4_issue(T_FEAT + lane_off + (uint32_t)(128 + 4 * e), dh);
          tm_ld4_issue(T_FEAT + lane_off + (uint32_t)(128 + 16 + 4 * e), dl);
          tm_ld4_issue(T_FEAT + lane_off + (uint32_t)(128 + 32), dqm);
          tm_ld_wait12(dh, dl, dqm);
          int cn[4] = {0, 0, 0, 0};
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const int4 c4 = *reinterpret_cast<const int4*>(cntp + (e * 4 + jj) * 4);
            cn[0] += c4.x; cn[1] += c4.y; cn[2] += c4.z; cn[3] += c4.w;
          }
          // rows of 36 floats: lanes that read different centres hit different banks
#pragma unroll
          for (int m = 0; m < 4; ++m)
            apriv[m * 36 + d] = (__uint_as_float(dh[m]) + __uint_as_float(dl[m]) + __uint_as_float(dqm[m])) * __fdividef(1.0f, (float)(cn[m] + 1));
        }
      }
}

template <int C> struct TmSmem {
  static constexpr int KC = C / 64;                                    // k-slabs of x
  static constexpr int NXB = C == 64 ? 2 : 1;                          // x tile buffers
  static constexpr int XB = C * TM_N * 2;                              // bytes of one x tile
  static constexpr int off_x = 0;
  static constexpr int off_o = off_x + NXB * XB;                       // 64 KB: W1 (feat hi|lo [, value hi]) / scratch / o operand
  static constexpr int off_d = off_o + 65536;                          // 32 KB: W2 (C = 128: value hi of W1 before it)
  static constexpr int off_wq = off_d + 32768;                         // [E][256 points] u32: (gate bits & ~3) | centre index
  static constexpr int off_cd = off_wq + 4096;                         // [16 warps][4 centres][36] fp32 centre aggregates a[m][d]
  static constexpr int off_p1 = off_cd + 16 * 4 * 36 * 4;              // [E][4 j][2][32] feat quadrant partial sums
  static constexpr int off_wb = off_p1 + 4096;                         // [4 k-slabs][48 rows][128 B] bf16: B operand of the aggregation GEMM
  static constexpr int off_stat = off_wb + 4 * TM_WB_SLAB;             // [256 samples] float2 (mean, rstd) of GroupNorm 1
  static constexpr int off_cnt = off_stat + 2048;                      // [E][4 j][4] member counts
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
  float* part1 = reinterpret_cast<float*>(smem + S::off_p1);
  float2* stat = reinterpret_cast<float2*>(smem + S::off_stat);
  int* cntp = reinterpret_cast<int*>(smem + S::off_cnt);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::off_bar);
  uint64_t* x_full = bars;              // [2]
  uint64_t* w1_full = bars + 2;
  uint64_t* w2_full = bars + 3;
  uint64_t* acc1_full = bars + 4;
  uint64_t* o_ready = bars + 5;
  uint64_t* acc2_full = bars + 6;
  uint64_t* epi_done = bars + 7;
  uint64_t* va_ready = bars + 8;
  uint64_t* accA_full = bars + 9;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  unsigned long long* const tr = (tid == 0 || tid == TM_CWARPS * 32) ? g_tm_trace : nullptr;
  if (tid == 0) tm_trace(tr, 0, 15);                                   // kernel entry
  const int regions_per_sample = A.F1 * A.F2;
  // a CTA owns a contiguous range of regions (mostly one sample: statistics look-ups and flushes stay rare, neighbours share L2 lines)
  const int rg_begin = (int)(((long long)blockIdx.x * A.regions) / gridDim.x);
  const int n_mine = (int)(((long long)(blockIdx.x + 1) * A.regions) / gridDim.x) - rg_begin;

  if (warp == TM_CWARPS) {
    if (lane == 0) {
      mbar_init(&x_full[0], 1); mbar_init(&x_full[1], 1);
      mbar_init(w1_full, 1); mbar_init(w2_full, 1);
      mbar_init(acc1_full, 1); mbar_init(acc2_full, 1);
      mbar_init(o_ready, TM_CWARPS); mbar_init(epi_done, TM_CWARPS);
      mbar_init(va_ready, TM_CWARPS); mbar_init(accA_full, 1);
      mbar_fence_init();
      tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmOut); tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmW2);
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (warp < TM_CWARPS) {
    // per-sample GroupNorm statistics, once (fp64 reduction of the slot sums: 16 warps, sample b = warp, warp + 16, ...)
    for (int b = warp; b < A.B; b += TM_CWARPS) {
      float mu, rstd;
      gn_mean_rstd(A.gn_sums, b, (double)C * (double)A.H * (double)A.W, A.gn_eps, mu, rstd);
      if (lane == 0) stat[b] = make_float2(mu, rstd);
    }
    // constant rows of WB: 32 + m = quadrant-mean weights (1/64 on the 64 points of quadrant m), 36..47 = 0.  One 16-byte chunk
    // (8 points of one row of one slab) per thread: 4 slabs x 16 rows x 8 chunks = 512.
    {
      const int slab = tid >> 7, r = 32 + ((tid >> 3) & 15), c = tid & 7;
      uint32_t w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint32_t v = 0;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int n = slab * 64 + c * 8 + 2 * i + hf;
          const int qd = 2 * ((n >> 4) >= 8) + ((n & 15) >= 8);
          if (r < 36 && qd == r - 32) v |= 0x3C80u << (16 * hf);            // bf16(1/64)
        }
        w[i] = v;
      }
      sts128(smem_u32(smem + S::off_wb) + (uint32_t)(slab * TM_WB_SLAB + (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)), w[0], w[1], w[2], w[3]);
    }
    fence_async_smem();
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

  if (warp >= TM_CWARPS) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(TM_REGS_AUX));
    // ================================ producer: TMA + MMA issue (one thread) ==================================================
    if (warp == TM_CWARPS && lane == 0 && n_mine > 0) {
      auto load_x = [&](int it) {
        int b, row0, col0;
        region_coords(rg_begin + it, b, row0, col0);
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
        tm_trace(tr, it, 0);
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
        tm_trace(tr, it, 1);
        mbar_wait(acc1_full, ph);                                          // W1 consumed: operand region free for the scratch / o
        tm_trace(tr, it, 2);
        if (!S::W2_RESIDENT) load_w2();
        // ---- aggregation GEMM (pass 3): [gate-weighted sums | quadrant means] of value, A operand = bf16 value in TMEM -------------
        mbar_wait(va_ready, ph);
        tm_fence_after();
        {
          const uint32_t wb_addr = smem_u32(smem + S::off_wb);
#pragma unroll 4
          for (int ks = 0; ks < 16; ++ks)
            tm_mma_ts(T_FEAT + 128u, T_FEAT + (uint32_t)(8 * ks), tm_desc_lo(wb_addr + (ks >> 2) * TM_WB_SLAB + (ks & 3) * 32, 16), TM_HI_SW128,
                      TM_IDESC_AGG, ks > 0 ? 1u : 0u);
          tm_commit(accA_full);
        }
        // ---- GEMM 2: D2[C x 256] = W2 . o   into the feat columns ----------------------------------------------------------------
        mbar_wait(o_ready, ph);
        if (!S::W2_RESIDENT || it == 0) { mbar_wait(w2_full, w2_phase); w2_phase ^= 1u; }
        tm_fence_after();
        tm_trace(tr, it, 3);
        {
          const uint32_t o_addr = smem_u32(sO);
          for (int s = 0; s < 2; ++s) {
            const uint32_t w_addr = smem_u32(sD + s * 16384);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              tm_mma<TM_IDESC_KK>(T_FEAT, tm_desc_lo(w_addr + ks * 32, 16), TM_HI_SW128, tm_desc_lo(o_addr + s * 32768 + ks * 32, 16), TM_HI_SW128,
                                  (s > 0 || ks > 0) ? 1u : 0u);
          }
          tm_commit(acc2_full);
        }
        mbar_wait(acc2_full, ph);                                          // o and W2 consumed
        tm_trace(tr, it, 4);
        if (it + 1 < n_mine) load_w1();
        // ---- the epilogue has rewritten the x tile in place: store it, then reuse the buffer ------------------------------------
        mbar_wait(epi_done, ph);
        tm_trace(tr, it, 5);
        {
          int b, row0, col0;
          region_coords(rg_begin + it, b, row0, col0);
          tma_store_4d(&tmOut, xt, col0, 0, row0, b);
          tma_store_commit();
        }
        tm_trace(tr, it, 6);
        if (it + NXB < n_mine) {
          tma_store_wait_read();                                            // (NXB = 2: this also covers the other buffer's older store)
          load_x(it + NXB);
        }
      }
      tma_store_wait_all();
    }
    __syncwarp();
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(TM_REGS_COMPUTE));
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
    const int dq = lane & 3, pg = lane >> 2;                             // pass-2 read mapping: channels 8dq..8dq+7, points 4pg..4pg+3
    // 4 KB transposition scratch, inside the k-slab of the o operand that this head pair writes later (pass 4)
    const uint32_t scratch = smem_u32(sO) + (uint32_t)((e >> 1) * 32768 + ((e & 1) * 4 + j) * 4096);
    const uint32_t wb_base = smem_u32(smem + S::off_wb);
    float* apriv = reinterpret_cast<float*>(smem + S::off_cd) + warp * (4 * 36);    // this warp's copy of a[m][d]
    float ssum = 0.f, ssq = 0.f;
    int cur_b = -1;
    float mu = 0.f, rstd = 1.f;
    constexpr float inv_q = 1.0f / 64.0f;

    for (int it = 0; it < n_mine; ++it) {
      const int rg = rg_begin + it;
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
        const float2 st = stat[b];
        mu = st.x; rstd = st.y;
        cur_b = b;
      }
      const float ehf = fmaf(-rstd * mu, k1f, k0f), ehv = fmaf(-rstd * mu, k1v, k0v);   // feat = rstd*acc + ehf, value = rstd*acc + ehv

      mbar_wait(acc1_full, ph);
      tm_fence_after();
      tm_trace(tr, it, 8);

      // ---- passes 1-3 + centre aggregates (tm_core_passes) ----------------------------------------------------------------------------
      uint32_t gk[2];
      {
        TmCore cc_;
        cc_.T_FEAT = T_FEAT; cc_.T_VAL = T_VAL; cc_.lane_off = lane_off; cc_.scratch = scratch; cc_.wb_base = wb_base; cc_.ph = ph;
        cc_.e = e; cc_.j = j; cc_.lane = lane; cc_.it = it; cc_.Wimg = A.W;
        cc_.part1 = part1; cc_.cntp = cntp; cc_.apriv = apriv; cc_.va_ready = va_ready; cc_.accA_full = accA_full;
        cc_.rstd = rstd; cc_.ehf = ehf; cc_.ehv = ehv; cc_.alpha = alpha; cc_.beta = beta;
        cc_.idx_out = A.idx_out; cc_.smax_out = A.smax_out;
        cc_.io_base = ((int64_t)(b * TM_E + e) * A.H + row0) * A.W + col0;
        cc_.tr = tr;
        tm_core_passes(cc_, gk);
      }
      // ---- pass 4: dispatch with lane = POINT: o[n][:] = g_n * a[k_n][:], bf16, written as the K-major SW128 B operand of GEMM 2
      //      (row = point, 64 bytes = this head's 32 channels) ------------------------------------------------------------------------------
      {
        __syncwarp();
#pragma unroll 1
        for (int q = 0; q < 2; ++q) {
          const float g = __uint_as_float(gk[q] & ~3u);
          const uint32_t arow = smem_u32(apriv) + (gk[q] & 3u) * 144u;
          const int n = 64 * j + 32 * q + lane;
          uint32_t ow[16];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint4 a4 = lds128(arow + (uint32_t)c * 16u);
            ow[2 * c] = bf16x2(g * __uint_as_float(a4.x), g * __uint_as_float(a4.y));
            ow[2 * c + 1] = bf16x2(g * __uint_as_float(a4.z), g * __uint_as_float(a4.w));
          }
          const uint32_t orow = smem_u32(sO) + (uint32_t)((e >> 1) * 32768 + n * 128);
#pragma unroll
          for (int t = 0; t < 4; ++t)
            sts128(orow + ((((uint32_t)(4 * (e & 1) + t)) ^ (uint32_t)(n & 7)) << 4), ow[4 * t], ow[4 * t + 1], ow[4 * t + 2], ow[4 * t + 3]);
        }
      }
      fence_async_smem();
      tm_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_ready);
      tm_trace(tr, it, 12);

      // ---- epilogue: + b2, * ls, + x (the tile itself), statistics, bf16 in place ------------------------------------------------------------
      mbar_wait(acc2_full, ph);
      tm_fence_after();
      tm_trace(tr, it, 13);
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
      tm_trace(tr, it, 14);
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

// ================================================================================================================================
// Projection + cluster core of the token mixer for the WIDE stages (stage 3: C = 320, 8 heads): one CTA per (16x16 region, group of
// four heads).  fc2 contracts over all heads, so it stays a separate launch; what is fused is GN-folded fc1|fc_v -> core, i.e. the
// fp32 `feat` / bf16 `value` round trip and one launch per block disappear, and a launch needs only regions x head-groups CTAs
// (64 at batch 8) — the image and the radar stacks run side by side on the 148 SMs.
//   x tile (C x 256 points, 160 KB) by TMA in two channel blocks (box dimensions are limited to 256); the fc1|fc_v weight slabs
//   (hi | lo for feat, hi for value) stream through a 3-deep TMA ring; after the projection the x tile is dead and its shared
//   memory holds the transposition scratch, the aggregation operand, the aggregate tables and the output tile, which leaves by
//   one TMA store in NCHW order.
template <int C> struct TkSmem {
  static_assert(C == 320, "projection + core kernel: the channel blocks below are laid out for C = 320");
  static constexpr int KC = C / 64;
  static constexpr int CA = 128, CB = C - 128;                         // channel blocks of the x tile (whole 64-channel slabs each)
  static constexpr int XB = C * TM_N * 2;
  static constexpr int off_x = 0;                                      // block A [16 rows][128][32 B], block B [16][192][32 B]
  static constexpr int off_xb = CA * TM_N * 2;
  // after GEMM 1 (x dead):
  static constexpr int off_scratch = 0;                                // 16 warps x 4 KB
  static constexpr int off_wb = 65536;                                 // aggregation operand, 4 x 6 KB
  static constexpr int off_ot = off_wb + 4 * TM_WB_SLAB;               // output tile [16 rows][128 channels][32 B]
  static_assert(off_ot % 1024 == 0 && off_ot + 65536 <= XB, "reuse of the x tile does not fit");
  static constexpr int off_ring = XB;                                  // 3 x 16 KB weight slabs
  static constexpr int off_p1 = off_ring + 3 * 16384;
  static constexpr int off_ap = off_p1 + 4096;                         // 16 warps x 4 x 36 floats: centre aggregates a[m][d]
  static constexpr int off_stat = off_ap + 16 * 4 * 36 * 4;
  static constexpr int off_cnt = off_stat + 2048;
  static constexpr int off_bar = off_cnt + 256;
  static constexpr int total = off_bar + 256 + 1024;
};

template <int C>
__global__ void __launch_bounds__(TM_THREADS, 1)
token_mixer_core_kernel(TmArgs A, int heads, const __grid_constant__ CUtensorMap tmXa, const __grid_constant__ CUtensorMap tmXb,
                        const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmW1) {
  using S = TkSmem<C>;
  constexpr int KC = S::KC;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  unsigned char* smem = smem_raw + pad;
  unsigned char* sX = smem + S::off_x;
  unsigned char* ring = smem + S::off_ring;
  float* part1 = reinterpret_cast<float*>(smem + S::off_p1);
  float2* stat = reinterpret_cast<float2*>(smem + S::off_stat);
  int* cntp = reinterpret_cast<int*>(smem + S::off_cnt);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::off_bar);
  uint64_t* x_full = bars;
  uint64_t* w_full = bars + 1;          // [3]
  uint64_t* w_free = bars + 4;          // [3]
  uint64_t* acc1_full = bars + 7;
  uint64_t* va_ready = bars + 8;
  uint64_t* accA_full = bars + 9;
  uint64_t* epi_done = bars + 10;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  unsigned long long* const tr = (tid == 0 || tid == TM_CWARPS * 32) ? g_tm_trace : nullptr;
  if (tid == 0) tm_trace(tr, 0, 15);
  const int HG = heads / 4, ED = heads * TM_D;
  const int regions_per_sample = A.F1 * A.F2;
  const int units = A.regions * HG;
  const int u_begin = (int)(((long long)blockIdx.x * units) / gridDim.x);
  const int n_mine = (int)(((long long)(blockIdx.x + 1) * units) / gridDim.x) - u_begin;

  if (warp == TM_CWARPS) {
    if (lane == 0) {
      mbar_init(x_full, 1);
      for (int i = 0; i < 3; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_free[i], 1); }
      mbar_init(acc1_full, 1); mbar_init(accA_full, 1);
      mbar_init(va_ready, TM_CWARPS); mbar_init(epi_done, TM_CWARPS);
      mbar_fence_init();
      tma_prefetch_desc(&tmXa); tma_prefetch_desc(&tmXb); tma_prefetch_desc(&tmO); tma_prefetch_desc(&tmW1);
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (warp < TM_CWARPS) {
    for (int b = warp; b < A.B; b += TM_CWARPS) {
      float mu, rstd;
      gn_mean_rstd(A.gn_sums, b, (double)C * (double)A.H * (double)A.W, A.gn_eps, mu, rstd);
      if (lane == 0) stat[b] = make_float2(mu, rstd);
    }
  }
  tm_fence_before();
  __syncthreads();
  tm_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t T_FEAT = tmem_base, T_VAL = tmem_base + 256u;

  auto unit_coords = [&](int u, int& b, int& g, int& row0, int& col0) {
    const int rg = u / HG;
    g = u - rg * HG;
    b = rg / regions_per_sample;
    const int q = rg - b * regions_per_sample;
    row0 = (q / A.F2) * TM_RS;
    col0 = (q % A.F2) * TM_RS;
  };

  if (warp >= TM_CWARPS) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(TM_REGS_AUX));
    if (warp > TM_CWARPS && lane == 0) {
      // ---- weight ring producers: per unit 2*KC feat slabs (hi | lo) then KC value slabs.  One issuing thread sustains one TMA box
      //      per ~0.33 us whatever its size (tools/tma_bw_probe.cu), so the three spare warps each own one slot of the 3-deep ring ----
      const int slot = warp - TM_CWARPS - 1;
      int cnt = 0;
      for (int it = 0; it < n_mine; ++it) {
        int b, g, row0, col0;
        unit_coords(u_begin + it, b, g, row0, col0);
        for (int s = 0; s < 3 * KC; ++s, ++cnt) {
          if (cnt % 3 != slot) continue;
          if (cnt >= 3) mbar_wait(&w_free[slot], (uint32_t)((cnt / 3) - 1) & 1u);
          mbar_expect_tx(&w_full[slot], 16384u);
          if (s < 2 * KC) tma_load_2d(ring + slot * 16384, &tmW1, s * 64, 128 * g, &w_full[slot]);
          else tma_load_2d(ring + slot * 16384, &tmW1, (s - 2 * KC) * 64, ED + 128 * g, &w_full[slot]);
        }
      }
    }
    if (warp == TM_CWARPS && lane == 0 && n_mine > 0) {
      // ---- x loads, MMA issue, output stores -------------------------------------------------------------------------------------
      auto load_x = [&](int it) {
        int b, g, row0, col0;
        unit_coords(u_begin + it, b, g, row0, col0);
        mbar_expect_tx(x_full, (uint32_t)S::XB);
        tma_load_4d(sX, &tmXa, col0, 0, row0, b, x_full);
        tma_load_4d(sX + S::off_xb, &tmXb, col0, S::CA, row0, b, x_full);
      };
      load_x(0);
      int cnt = 0;
      for (int it = 0; it < n_mine; ++it) {
        const uint32_t ph = (uint32_t)it & 1u;
        mbar_wait(x_full, ph);
        tm_fence_after();
        tm_trace(tr, it, 0);
        for (int s = 0; s < 3 * KC; ++s, ++cnt) {
          const int slot = cnt % 3;
          mbar_wait(&w_full[slot], (uint32_t)(cnt / 3) & 1u);
          tm_fence_after();
          const bool is_feat = s < 2 * KC;
          const int kx = is_feat ? s % KC : s - 2 * KC;                   // x slab: 0,1 in block A, 2.. in block B
          const uint32_t x_addr = kx < 2 ? smem_u32(sX) + (uint32_t)(kx * 64 * 32) : smem_u32(sX + S::off_xb) + (uint32_t)((kx - 2) * 64 * 32);
          const uint32_t x_lbo = kx < 2 ? (uint32_t)(S::CA * 32) : (uint32_t)(S::CB * 32);
          const uint32_t w_addr = smem_u32(ring + slot * 16384);
          const bool first = is_feat ? s == 0 : s == 2 * KC;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            tm_mma(is_feat ? T_FEAT : T_VAL, tm_desc_lo(w_addr + ks * 32, 16), TM_HI_SW128, tm_desc_lo(x_addr + ks * 16 * 32, x_lbo), TM_HI_SW32,
                   (first && ks == 0) ? 0u : 1u);
          tm_commit(&w_free[slot]);
        }
        tm_commit(acc1_full);
        tm_trace(tr, it, 1);
        // ---- aggregation GEMM (pass 3) ---------------------------------------------------------------------------------------------
        mbar_wait(va_ready, ph);
        tm_fence_after();
        tm_trace(tr, it, 3);
        {
          const uint32_t wb_addr = smem_u32(sX + S::off_wb);
#pragma unroll 4
          for (int ks = 0; ks < 16; ++ks)
            tm_mma_ts(T_FEAT + 128u, T_FEAT + (uint32_t)(8 * ks), tm_desc_lo(wb_addr + (ks >> 2) * TM_WB_SLAB + (ks & 3) * 32, 16), TM_HI_SW128,
                      TM_IDESC_AGG, ks > 0 ? 1u : 0u);
          tm_commit(accA_full);
        }
        // ---- the output tile is complete: store it, then the x region may be refilled ----------------------------------------------------
        mbar_wait(epi_done, ph);
        tm_trace(tr, it, 5);
        {
          int b, g, row0, col0;
          unit_coords(u_begin + it, b, g, row0, col0);
          tma_store_4d(&tmO, sX + S::off_ot, col0, 128 * g, row0, b);
          tma_store_commit();
        }
        tm_trace(tr, it, 6);
        if (it + 1 < n_mine) {
          tma_store_wait_read();
          load_x(it + 1);
        }
      }
      tma_store_wait_all();
    }
    __syncwarp();
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(TM_REGS_COMPUTE));
    const int e = warp & 3, j = warp >> 2;
    const int d = lane;
    const uint32_t lane_off = (uint32_t)(32 * e) << 16;
    const float alpha = __ldg(A.alpha), beta = __ldg(A.beta);
    float* apriv = reinterpret_cast<float*>(smem + S::off_ap) + warp * (4 * 36);
    const uint32_t ot = smem_u32(sX + S::off_ot);

    for (int it = 0; it < n_mine; ++it) {
      int b, g, row0, col0;
      unit_coords(u_begin + it, b, g, row0, col0);
      const uint32_t ph = (uint32_t)it & 1u;
      const int ch = 128 * g + 32 * e + d;                               // this lane's feat / value channel
      const float2 st = stat[b];
      const float ehf = fmaf(-st.y * st.x, __ldg(A.k1 + ch), __ldg(A.k0 + ch));
      const float ehv = fmaf(-st.y * st.x, __ldg(A.k1 + ED + ch), __ldg(A.k0 + ED + ch));
      mbar_wait(acc1_full, ph);
      tm_fence_after();
      tm_trace(tr, it, 8);
      // the x tile is dead: constant rows of the aggregation operand (quadrant means, zero padding), one 16-byte chunk per thread
      {
        const int slab = tid >> 7, r = 32 + ((tid >> 3) & 15), c = tid & 7;
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint32_t v = 0;
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            const int n = slab * 64 + c * 8 + 2 * i + hf;
            const int qd = 2 * ((n >> 4) >= 8) + ((n & 15) >= 8);
            if (r < 36 && qd == r - 32) v |= 0x3C80u << (16 * hf);
          }
          w[i] = v;
        }
        sts128(smem_u32(sX + S::off_wb) + (uint32_t)(slab * TM_WB_SLAB + (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)), w[0], w[1], w[2], w[3]);
      }
      uint32_t gk[2];
      {
        TmCore cc_;
        cc_.T_FEAT = T_FEAT; cc_.T_VAL = T_VAL; cc_.lane_off = lane_off;
        cc_.scratch = smem_u32(sX + S::off_scratch) + (uint32_t)((e * 4 + j) * 4096);
        cc_.wb_base = smem_u32(sX + S::off_wb); cc_.ph = ph;
        cc_.e = e; cc_.j = j; cc_.lane = lane; cc_.it = it; cc_.Wimg = A.W;
        cc_.part1 = part1; cc_.cntp = cntp; cc_.apriv = apriv; cc_.va_ready = va_ready; cc_.accA_full = accA_full;
        cc_.rstd = st.y; cc_.ehf = ehf; cc_.ehv = ehv; cc_.alpha = alpha; cc_.beta = beta;
        cc_.idx_out = A.idx_out; cc_.smax_out = A.smax_out;
        cc_.io_base = ((int64_t)(b * heads + 4 * g + e) * A.H + row0) * A.W + col0;
        cc_.tr = tr;
        tm_core_passes(cc_, gk);
      }
      // ---- pass 4: lane = point, o[n][:] = g_n * a[k_n][:] into the NCHW-ordered output tile [row][channel][16 cols] (SWIZZLE_32B) --
      __syncwarp();
#pragma unroll 1
      for (int q = 0; q < 2; ++q) {
        const float gq = __uint_as_float(gk[q] & ~3u);
        const uint32_t arow = smem_u32(apriv) + (gk[q] & 3u) * 144u;
        const int n = 64 * j + 32 * q + lane;
        const uint32_t pbase = ot + (uint32_t)((n >> 4) * 4096 + (n & 7) * 2);
        const uint32_t hi8 = (uint32_t)((n >> 3) & 1);
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
          const uint4 a4 = lds128(arow + (uint32_t)c4 * 16u);
          const float v[4] = {gq * __uint_as_float(a4.x), gq * __uint_as_float(a4.y), gq * __uint_as_float(a4.z), gq * __uint_as_float(a4.w)};
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const int c = 32 * e + 4 * c4 + t;                             // channel inside the tile
            sts16(pbase + (uint32_t)(c * 32) + ((hi8 ^ (uint32_t)((c >> 2) & 1)) << 4), (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v[t])));
          }
        }
      }
      fence_async_smem();
      tm_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(epi_done);
      tm_trace(tr, it, 12);
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

extern "C" int vrcoc_debug_set_tm_trace(unsigned long long* buf) {
  return cudaMemcpyToSymbol(g_tm_trace, &buf, sizeof(buf)) == cudaSuccess ? 0 : -2;
}

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

extern "C" int vrcoc_token_mixer_core_supported(int dtype, int C, int H, int W, int heads, int head_dim, int fold_w, int fold_h,
                                                int proposal_w, int proposal_h) {
  const char* knob = getenv("VRCOC_TM_FUSED");
  if (knob && knob[0] == '0') return 0;
  const bool folded = fold_w > 1 && fold_h > 1;
  const int f1 = folded ? fold_w : 1, f2 = folded ? fold_h : 1;
  return dtype == VRCOC_BF16 && C == 320 && heads > 0 && heads % TM_E == 0 && head_dim == TM_D && proposal_w == 2 && proposal_h == 2 &&
         H % f1 == 0 && W % f2 == 0 && H / f1 == TM_RS && W / f2 == TM_RS && tma_encode_fn() != nullptr;
}

extern "C" int vrcoc_token_mixer_core_fwd(const void* x, const double* gn_sums, float gn_eps, const void* w_fold, const float* k0,
                                          const float* k1, const float* alpha, const float* beta, void* o, uint8_t* idx, float* sim_max,
                                          int B, int C, int H, int W, int heads, int head_dim, int fold_w, int fold_h, void* stream) {
  VRCOC_REQUIRE(x && gn_sums && w_fold && k0 && k1 && alpha && beta && o && B > 0 && B <= 256, "token_mixer_core: null pointer / bad batch");
  VRCOC_REQUIRE(vrcoc_token_mixer_core_supported(VRCOC_BF16, C, H, W, heads, head_dim, fold_w, fold_h, 2, 2),
                "token_mixer_core: unsupported geometry C=%d %dx%d heads=%d head_dim=%d fold=%dx%d", C, H, W, heads, head_dim, fold_w, fold_h);
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  VRCOC_REQUIRE(al(x) && al(o) && al(w_fold), "token_mixer_core: tensors must be 16-byte aligned");
  VRCOC_REQUIRE(!idx || (reinterpret_cast<uintptr_t>(idx) & 1) == 0, "token_mixer_core: idx must be 2-byte aligned");
  const int ED = heads * head_dim;
  TmArgs A;
  A.gn_sums = gn_sums; A.gn_eps = gn_eps; A.k0 = k0; A.k1 = k1; A.alpha = alpha; A.beta = beta; A.b2 = nullptr; A.ls = nullptr;
  A.out_sums = nullptr; A.idx_out = idx; A.smax_out = sim_max;
  A.B = B; A.H = H; A.W = W; A.F1 = H / TM_RS; A.F2 = W / TM_RS; A.regions = B * A.F1 * A.F2;
  CUtensorMap txa, txb, to, tw1;
  int rc;
  {
    cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)C, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)H * W * 2, (cuuint64_t)W * 2, (cuuint64_t)C * H * W * 2};
    cuuint32_t boxa[4] = {(cuuint32_t)TM_RS, 128, (cuuint32_t)TM_RS, 1}, boxb[4] = {(cuuint32_t)TM_RS, (cuuint32_t)(C - 128), (cuuint32_t)TM_RS, 1};
    if ((rc = tma_encode_sw(&txa, VRCOC_BF16, x, 4, dims, strides, boxa, 32))) return rc;
    if ((rc = tma_encode_sw(&txb, VRCOC_BF16, x, 4, dims, strides, boxb, 32))) return rc;
    cuuint64_t odims[4] = {(cuuint64_t)W, (cuuint64_t)ED, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t ostr[3] = {(cuuint64_t)H * W * 2, (cuuint64_t)W * 2, (cuuint64_t)ED * H * W * 2};
    cuuint32_t obox[4] = {(cuuint32_t)TM_RS, 128, (cuuint32_t)TM_RS, 1};
    if ((rc = tma_encode_sw(&to, VRCOC_BF16, o, 4, odims, ostr, obox, 32))) return rc;
    cuuint64_t d1[2] = {(cuuint64_t)(2 * C), (cuuint64_t)(2 * ED)}, s1[1] = {(cuuint64_t)(2 * C) * 2};
    cuuint32_t b1[2] = {64, 128};
    if ((rc = tma_encode_sw(&tw1, VRCOC_BF16, w_fold, 2, d1, s1, b1, 128))) return rc;
  }
  using S = TkSmem<320>;
  auto kern = token_mixer_core_kernel<320>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::total);
  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  const int units = A.regions * (heads / TM_E);
  int grid = sm_count();
  if (grid > units) grid = units;
  kern<<<grid, TM_THREADS, S::total, (cudaStream_t)stream>>>(A, heads, txa, txb, to, tw1);
  return check_launch("token_mixer_core");
}

