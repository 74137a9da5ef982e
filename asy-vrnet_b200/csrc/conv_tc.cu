// Convolution-as-GEMM engine, tcgen05 path (5th-gen tensor cores, bf16 operands, fp32 accumulation in TMEM).
// Same operator contract as the CUDA-core path (include/vrcoc.h, conv_common.cuh); used whenever the weights are bf16.
//
//   D[128 points x N outs] (TMEM, fp32)  +=  A[128 points x 64 k] (smem, bf16)  *  B[N outs x 64 k]^T (smem, bf16)
//
// * M = points.  In NCHW the points of one channel are contiguous, so A is an MN-major operand: each K-row of the
//   smem tile holds 64 consecutive points (128 B) of one logical input channel, two 64-point blocks per tile,
//   128-byte swizzled — the canonical UMMA layout  Sw<3,4,3> o ((8,2),(8,k)) : ((1,LBO),(8,SBO))  in 16-byte units
//   with LBO = 8192 B (next 64-point block) and SBO = 1024 B (next group of 8 k-rows).
// * B = weights [O][K] row-major = K-major operand, fetched by TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B box of
//   64 k x N rows, out-of-range rows/columns zero-filled by the hardware) straight into the UMMA layout.
// * Two kernels share the epilogue:
//   conv_tc_tma_kernel   — prologue-free 1x1 projections (Cluster.fc2, Mlp.fc2, BaseConv 1x1, dgrad): A also arrives
//                          by TMA (3-D map [B][C][P], two 64-point x 64-channel boxes per slab land directly in the
//                          MN-major layout).  Warp-specialised: one producer thread runs the TMA ring ahead, one thread
//                          issues tcgen05.mma; no SM instruction touches the operands.
//   conv_tc_xform_kernel — everything with a prologue or a gather (GroupNorm->fc1|fc_v, GroupNorm->mlp.fc1, the
//                          attention/ECA/shuffle prologue of RadarEnhanceByImage, k x k convs): A is produced by the
//                          CTA's threads ("transform on load"): 128-bit coalesced loads, prologue in fp32 registers, one
//                          rounding to bf16, one swizzled 16-byte st.shared; the raw loads of slab k+1 are in flight while
//                          the tensor core works on slab k.  The normalised/gated/gathered activation never exists in HBM.
// * tcgen05.commit on per-stage mbarriers frees smem stages; a last commit publishes the accumulator.
// * Epilogue: 8 warps read the accumulator with tcgen05.ld (32 lanes x 16 columns per instruction; warp w owns TMEM lanes
//   32*(w%4).. and column half w/4), apply bias/BN/activation/layer-scale/residual/BN (coefficients staged in smem) and the
//   side statistics in registers and store NCHW directly: lane = point, so every output channel is one coalesced
//   64/128-byte store per warp.  ncu (profiles/) showed the first version of this epilogue was issue-bound at ~160 SASS
//   instructions per output column; pointers are therefore hoisted, coefficients read with ld.shared, the activation is
//   a template parameter and GELU uses a branch-free erfc (|err| < 2e-7, bf16 outputs).
#include <stdio.h>
#include <stdlib.h>

#include "conv_common.cuh"
#include "tma.cuh"

namespace vrcoc {

constexpr int TC_THREADS = 256;
constexpr int TC_BM = 128;            // points per CTA
constexpr int TC_BK = 64;             // k per smem slab (128 B of bf16 per B row)
constexpr int TC_A_BYTES = TC_BM * TC_BK * 2;       // 16 KB
constexpr int TC_A_LBO = 64 * TC_BK * 2;            // 8192 B: second 64-point block
constexpr int TC_MAX_STAGES = 4;

__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) |
// version=1 [46,48) | layout SWIZZLE_128B=2 [61,64)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | (2ull << 61);
}

// instruction descriptor (cute::UMMA::InstrDescriptor) for kind::f16: D=f32, A=B=bf16, A MN-major, B K-major, M=128
__host__ __device__ constexpr uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (0u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
}

__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// ---- lean MMA issue ---------------------------------------------------------------------------------------------------
// ONE thread issues every tcgen05.mma of a CTA, and its instruction stream is on the critical path: the per-tile timestamp trace
// of the channel-major kernel showed the issue loop taking 2.8 us for the 20 MMAs of a 128x128x320 tile (245 clk per MMA, the
// tensor pipe needs 64) - rebuilding two 64-bit descriptors with shifts / masks, run-time accumulate flags and loop control per
// MMA, in a single dependent chain that also competes for issue slots with two epilogue warps.  Here a descriptor is a
// {lo, hi} register pair: hi (SBO, version, swizzle) is a constant, lo (address and LBO in 16-byte units) advances by an
// immediate per k-step; the four k-steps of a 64-deep slab are straight-line code.
constexpr uint32_t TC_DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);          // SBO = 1024 B, version 1, SWIZZLE_128B
__device__ __forceinline__ uint32_t tc_desc_lo(uint32_t saddr, uint32_t lbo_bytes) {
  return ((saddr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
template <bool ACC>
__device__ __forceinline__ void tc_mma_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate = 1u) {
  if (ACC) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b64 da, db;\n\t"
        "setp.eq.u32 p, 0, 0;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t"
        "}" ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(TC_DESC_HI), "r"(idesc)
        : "memory");
  } else {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t"
        "}" ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(TC_DESC_HI), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// the k-steps of one slab: operand A advances by a_step, B by b_step (16-byte units) per 16 k; `acc0` = accumulate flag of the
// first MMA (0 only for the first slab of a tile)
template <int A_STEP, int B_STEP>
__device__ __forceinline__ void tc_issue_slab(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t acc0, int ksteps) {
  if (ksteps == 4) {
    tc_mma_lo<false>(tmem_d, a_lo, b_lo, idesc, acc0);
    tc_mma_lo<true>(tmem_d, a_lo + A_STEP, b_lo + B_STEP, idesc);
    tc_mma_lo<true>(tmem_d, a_lo + 2 * A_STEP, b_lo + 2 * B_STEP, idesc);
    tc_mma_lo<true>(tmem_d, a_lo + 3 * A_STEP, b_lo + 3 * B_STEP, idesc);
  } else {
    for (int j = 0; j < ksteps; ++j) tc_mma_lo<false>(tmem_d, a_lo + j * A_STEP, b_lo + j * B_STEP, idesc, (acc0 || j > 0) ? 1u : 0u);
  }
}

// issue the K=16 MMAs of one slab: A MN-major (16 k-rows = two 8-row groups = 2048 B), B K-major (16 k = 32 B inside the row)
__device__ __forceinline__ void issue_slab_mmas(uint32_t tmem_base, uint32_t a_addr, uint32_t b_addr, uint32_t idesc, int ksteps,
                                                bool first_slab) {
  tc_issue_slab<2048 / 16, 32 / 16>(tmem_base, tc_desc_lo(a_addr, TC_A_LBO), tc_desc_lo(b_addr, 16), idesc, first_slab ? 0u : 1u, ksteps);
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ uint4 pack8_bf16(const float (&v)[8]) {
  uint4 r;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  return r;
}

// branch-free erf GELU with ONE special-function op:
//   gelu(x) = max(x, 0) - a * 2^P(a),   a = min(|x|, AMAX),   2^P(a) ~ 0.5 erfc(a / sqrt 2)
// P is a fit of -(log2(e) * -ln erfc(a/sqrt2)) - 1 on [0, AMAX], weighted for the error of a * 0.5 erfc (tools/fit_gelu.py).
// These kernels contract bf16 operands and (the hidden layer of the channel MLP) store bf16: P has DEGREE 3, |gelu error|
// < 5.5e-5 over the whole line — two orders below the bf16 rounding of the operands and of the result — and three FMAs per
// value cheaper than the degree-6 fit (3e-7) used before: ncu showed the hidden-layer epilogues bound by instruction issue
// (profiles/r01_ncu_final_cm_mlpf_core.csv).  The MUFU pipe (16 lanes / clk / SM) is the other bound of a GELU epilogue on
// sm_100a: the earlier A&S 7.1.26 form needed a reciprocal as well as the exponential and ran at half the rate.
// The fp32 engine (conv_simt.cu) keeps erff.
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
#define VRCOC_GELU_AMAX 5.091169f
#define VRCOC_GELU_C0 -1.003531933e+00f
#define VRCOC_GELU_C1 -1.129245043e+00f
#define VRCOC_GELU_C2 -4.988209009e-01f
#define VRCOC_GELU_C3 -2.488539740e-02f
__device__ __forceinline__ float gelu_fast(float x) {
  const float a = fminf(fabsf(x), VRCOC_GELU_AMAX);
  float p = fmaf(VRCOC_GELU_C3, a, VRCOC_GELU_C2);
  p = fmaf(p, a, VRCOC_GELU_C1);
  p = fmaf(p, a, VRCOC_GELU_C0);
  return fmaf(-a, ex2_approx(p), fmaxf(x, 0.f));
}
__device__ __forceinline__ float sigmoid_fast(float x) { return rcp_approx(1.0f + ex2_approx(x * -1.4426950408889634f)); }

template <int ACT>
__device__ __forceinline__ float act_tc(float y) {
  if (ACT == VRCOC_ACT_RELU) return fmaxf(y, 0.f);
  if (ACT == VRCOC_ACT_GELU) return gelu_fast(y);
  if (ACT == VRCOC_ACT_SILU) return y * sigmoid_fast(y);
  if (ACT == VRCOC_ACT_LRELU) return y > 0.f ? y : 0.1f * y;
  return y;
}

struct TcLayout {
  int n_tile;       // output channels per CTA (multiple of 32, <= 256)
  int tmem_cols;    // power of two >= 32
  int b_bytes;      // n_tile * 128
  int stages;
  int use_tma_b;    // weights via TMA
  int plain_epi;    // epilogue is y = act(acc*es + eh) into `out` only (no residual / post / final affine / stats / split)
  int res_tma;      // residual epilogue through a TMA-staged [n_tile][128] bf16 tile (fetched at kernel start, rewritten in
                    // place, stored by TMA): no per-thread global access, full memory-level parallelism
  int off_b, off_tab, off_epi, off_res, off_bar, total;
};

struct TcSmem {
  unsigned char* base;
  unsigned char* sA; unsigned char* sB; float4* tab; float* epi;
  uint64_t* bar_free; uint64_t* bar_full; uint64_t* bar_acc; uint64_t* bar_res; uint32_t* tmem_slot;
  __nv_bfloat16* res_tile;
};

__device__ __forceinline__ TcSmem carve(unsigned char* smem_raw, const TcLayout& L) {
  // 1024-byte alignment is required by the 128-byte swizzle atoms; computed on the shared-space address so that the
  // compiler keeps emitting ld/st.shared for everything derived from it
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  TcSmem s;
  s.base = smem_raw + pad;
  s.sA = s.base;
  s.sB = s.base + L.off_b;
  s.tab = reinterpret_cast<float4*>(s.base + L.off_tab);
  s.epi = reinterpret_cast<float*>(s.base + L.off_epi);
  s.bar_free = reinterpret_cast<uint64_t*>(s.base + L.off_bar);
  s.bar_full = s.bar_free + TC_MAX_STAGES;
  s.bar_acc = s.bar_full + TC_MAX_STAGES;
  s.bar_res = s.bar_acc + 1;
  s.tmem_slot = reinterpret_cast<uint32_t*>(s.bar_res + 1);
  s.res_tile = reinterpret_cast<__nv_bfloat16*>(s.base + L.off_res);
  return s;
}

// common prologue: TMEM allocation, barrier init, epilogue coefficient staging
__device__ __forceinline__ uint32_t tc_setup(const ConvArgs& a, const TcLayout& L, const TcSmem& S, int n0) {
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(S.tmem_slot)), "r"((uint32_t)L.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 32) {
    for (int i = 0; i < L.stages; ++i) { mbar_init(&S.bar_free[i], 1); mbar_init(&S.bar_full[i], 1); }
    mbar_init(S.bar_acc, 1);
    mbar_init(S.bar_res, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int n = tid; n < L.n_tile; n += TC_THREADS) {
    const int o = n0 + n;
    const bool in = o < a.O;
    S.epi[n] = (in && a.e_scale) ? a.e_scale[o] : 1.f;
    S.epi[L.n_tile + n] = (in && a.e_shift) ? a.e_shift[o] : 0.f;
    S.epi[2 * L.n_tile + n] = (in && a.post_scale) ? a.post_scale[o] : 1.f;
    S.epi[3 * L.n_tile + n] = (in && a.f_scale) ? a.f_scale[o] : 1.f;
    S.epi[4 * L.n_tile + n] = (in && a.f_shift) ? a.f_shift[o] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  return *S.tmem_slot;
}

// ---- epilogue -------------------------------------------------------------------------------------------------------------
// PLAIN: out[o] = act(acc*es + eh), single output tensor of type TO.
template <int ACT, typename TO>
__device__ __forceinline__ void epi_plain_group(const uint32_t (&r)[16], const float* es, const float* eh, TO* optr, int64_t P, int lim,
                                                bool valid, bool fast) {
  if (fast) {                                // warp-uniform fast path: no per-column predicates
#pragma unroll
    for (int j = 0; j < 16; ++j) stf<TO>(optr + j * P, act_tc<ACT>(fmaf(__uint_as_float(r[j]), es[j], eh[j])));
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (j < lim && valid) stf<TO>(optr + j * P, act_tc<ACT>(fmaf(__uint_as_float(r[j]), es[j], eh[j])));
  }
}

// PLAIN: out[o] = act(acc*es + eh); channels [0,O_split) go to `out`, the rest to `out2` (a 16-column group never straddles
// the split: O_split % 16 == 0, host-checked), each with its own storage type.
template <int ACT>
__device__ __forceinline__ void epi_plain(const ConvArgs& a, const TcLayout& L, const float* epi, uint32_t tbase, int c_begin,
                                          int c_end, int n0, int b, int q, bool valid) {
  const int64_t P = a.P_out;
  const bool all_valid = __all_sync(0xffffffffu, valid);
  for (int c0 = c_begin; c0 < c_end; c0 += 16) {
    uint32_t r[16];
    tmem_ld16(tbase + (uint32_t)c0, r);
    const int o0 = n0 + c0;
    const int lim = min(16, a.O - o0);
    const bool fast = lim == 16 && all_valid;
    const bool second = o0 >= a.O_split;
    const int odt = second ? a.out2_dtype : a.out_dtype;
    void* obase = second ? a.out2 : a.out;
    const int64_t oidx = (second ? ((int64_t)b * (a.O - a.O_split) + (o0 - a.O_split)) : ((int64_t)b * a.O_split + o0)) * P + q;
    const float* es = epi + c0;
    const float* eh = epi + L.n_tile + c0;
    if (odt == VRCOC_BF16) epi_plain_group<ACT, __nv_bfloat16>(r, es, eh, reinterpret_cast<__nv_bfloat16*>(obase) + oidx, P, lim, valid, fast);
    else epi_plain_group<ACT, float>(r, es, eh, reinterpret_cast<float*>(obase) + oidx, P, lim, valid, fast);
  }
}

// FULL: y = act(acc*es + eh)*ps + res; y = y*fs + fh; split outputs; side statistics.
// The residual values of a 16-column group are requested one group ahead (the first group before the accumulator is
// waited for), so their DRAM latency overlaps the main loop / the previous group instead of stalling every column.
__device__ __forceinline__ void load_res16(const ConvArgs& a, int64_t ridx, int64_t P, int lim, bool valid, float (&rv)[16]) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    rv[j] = 0.f;
    if (j < lim && valid)
      rv[j] = (a.res_dtype == VRCOC_BF16) ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(a.res)[ridx + j * P])
                                           : reinterpret_cast<const float*>(a.res)[ridx + j * P];
  }
}

template <int ACT>
__device__ __forceinline__ void epi_full(const ConvArgs& a, const TcLayout& L, const float* epi, uint32_t tbase, int c_begin,
                                         int c_end, int n0, int b, int q, bool valid, uint64_t* bar_acc) {
  const int64_t P = a.P_out;
  const int nt = L.n_tile;
  float ssum = 0.f, ssq = 0.f, vmax = 0.f, vmin = __int_as_float(0x7f800000);
  float rnext[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) rnext[j] = 0.f;
  const bool has_res = a.res != nullptr;
  if (has_res && c_begin < c_end)
    load_res16(a, ((int64_t)b * a.O + n0 + c_begin) * P + q, P, min(16, a.O - n0 - c_begin), valid, rnext);
  mbar_wait(bar_acc, 0);
  tc_fence_after();
  for (int c0 = c_begin; c0 < c_end; c0 += 16) {
    float rcur[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) rcur[j] = rnext[j];
    if (has_res && c0 + 16 < c_end)
      load_res16(a, ((int64_t)b * a.O + n0 + c0 + 16) * P + q, P, min(16, a.O - n0 - c0 - 16), valid, rnext);
    uint32_t r[16];
    tmem_ld16(tbase + (uint32_t)c0, r);
    const int o0 = n0 + c0;
    const int lim = min(16, a.O - o0);
    // a 16-column group never straddles O_split when O_split % 16 == 0 (checked on the host)
    const bool second = o0 >= a.O_split;
    const int odt = second ? a.out2_dtype : a.out_dtype;
    unsigned char* obase = reinterpret_cast<unsigned char*>(second ? a.out2 : a.out);
    const int64_t ochan = second ? ((int64_t)b * (a.O - a.O_split) + (o0 - a.O_split)) : ((int64_t)b * a.O_split + o0);
    const int64_t oidx = ochan * P + q;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (j < lim && valid) {
        const int n = c0 + j;
        float y = act_tc<ACT>(fmaf(__uint_as_float(r[j]), epi[n], epi[nt + n]));
        y = fmaf(y, epi[2 * nt + n], rcur[j]);
        y = fmaf(y, epi[3 * nt + n], epi[4 * nt + n]);
        ssum += y; ssq = fmaf(y, y, ssq);
        vmax = fmaxf(vmax, y); vmin = fminf(vmin, y);
        if (odt == VRCOC_BF16) reinterpret_cast<__nv_bfloat16*>(obase)[oidx + j * P] = __float2bfloat16_rn(y);
        else reinterpret_cast<float*>(obase)[oidx + j * P] = y;
      }
    }
  }
  emit_side_stats(a, b, ssum, ssq, vmax, vmin);
}

// FULL epilogue on the TMA-staged tile: res_tile[n][m] (bf16, n = output channel in the tile, m = point in the tile) holds
// the residual on entry and the result on exit.
template <int ACT>
__device__ __forceinline__ void epi_full_tile(const ConvArgs& a, const TcLayout& L, const TcSmem& S, uint32_t tbase, int c_begin,
                                              int c_end, int n0, int b, int m, bool valid) {
  const int nt = L.n_tile;
  const float* epi = S.epi;
  float ssum = 0.f, ssq = 0.f, vmax = 0.f, vmin = __int_as_float(0x7f800000);
  mbar_wait(S.bar_acc, 0);
  tc_fence_after();
  mbar_wait(S.bar_res, 0);
  for (int c0 = c_begin; c0 < c_end; c0 += 16) {
    uint32_t r[16];
    tmem_ld16(tbase + (uint32_t)c0, r);
    const int lim = min(16, a.O - n0 - c0);
    __nv_bfloat16* rt = S.res_tile + c0 * TC_BM + m;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (j < lim) {
        const int n = c0 + j;
        float y = act_tc<ACT>(fmaf(__uint_as_float(r[j]), epi[n], epi[nt + n]));
        y = fmaf(y, epi[2 * nt + n], __bfloat162float(rt[j * TC_BM]));
        y = fmaf(y, epi[3 * nt + n], epi[4 * nt + n]);
        rt[j * TC_BM] = __float2bfloat16_rn(y);
        if (valid) {
          ssum += y; ssq = fmaf(y, y, ssq);
          vmax = fmaxf(vmax, y); vmin = fminf(vmin, y);
        }
      }
    }
  }
  emit_side_stats(a, b, ssum, ssq, vmax, vmin);
}

// waits for the accumulator, then drains it
__device__ __forceinline__ void tc_epilogue(const ConvArgs& a, const TcLayout& L, const TcSmem& S, uint32_t tmem_base, int p0,
                                            int n0, int b, const CUtensorMap* tmapO) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int lq = warp & 3, chalf = warp >> 2;
  const int q = p0 + lq * 32 + lane;
  const bool valid = q < a.P_out;
  const int ncols = L.n_tile >> 1;
  const uint32_t tbase = tmem_base + ((uint32_t)(lq * 32) << 16);
  const int c_begin = chalf * ncols;
  int c_end = (chalf + 1) * ncols;
  if (c_end > a.O - n0) c_end = a.O - n0;        // warp-uniform: skip all-padding column groups
  if (L.res_tma) {
    const int m = lq * 32 + lane;
    switch (a.act) {
      case VRCOC_ACT_NONE: epi_full_tile<VRCOC_ACT_NONE>(a, L, S, tbase, c_begin, c_end, n0, b, m, valid); break;
      case VRCOC_ACT_RELU: epi_full_tile<VRCOC_ACT_RELU>(a, L, S, tbase, c_begin, c_end, n0, b, m, valid); break;
      default: epi_full_tile<VRCOC_ACT_GELU>(a, L, S, tbase, c_begin, c_end, n0, b, m, valid); break;
    }
    fence_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
      tma_store_3d(tmapO, S.res_tile, p0, n0, b);     // clipped to the tensor: padding channels / points are not written
      tma_store_commit();
      tma_store_wait_read();
    }
  } else if (L.plain_epi) {
    mbar_wait(S.bar_acc, 0);
    tc_fence_after();
#define PLAIN(ACTV) epi_plain<ACTV>(a, L, S.epi, tbase, c_begin, c_end, n0, b, q, valid)
    switch (a.act) {
      case VRCOC_ACT_NONE: PLAIN(VRCOC_ACT_NONE); break;
      case VRCOC_ACT_RELU: PLAIN(VRCOC_ACT_RELU); break;
      case VRCOC_ACT_GELU: PLAIN(VRCOC_ACT_GELU); break;
      case VRCOC_ACT_SILU: PLAIN(VRCOC_ACT_SILU); break;
      default: PLAIN(VRCOC_ACT_LRELU); break;
    }
#undef PLAIN
  } else {
    switch (a.act) {
      case VRCOC_ACT_NONE: epi_full<VRCOC_ACT_NONE>(a, L, S.epi, tbase, c_begin, c_end, n0, b, q, valid, S.bar_acc); break;
      case VRCOC_ACT_RELU: epi_full<VRCOC_ACT_RELU>(a, L, S.epi, tbase, c_begin, c_end, n0, b, q, valid, S.bar_acc); break;
      case VRCOC_ACT_GELU: epi_full<VRCOC_ACT_GELU>(a, L, S.epi, tbase, c_begin, c_end, n0, b, q, valid, S.bar_acc); break;
      case VRCOC_ACT_SILU: epi_full<VRCOC_ACT_SILU>(a, L, S.epi, tbase, c_begin, c_end, n0, b, q, valid, S.bar_acc); break;
      default: epi_full<VRCOC_ACT_LRELU>(a, L, S.epi, tbase, c_begin, c_end, n0, b, q, valid, S.bar_acc); break;
    }
  }
}

__device__ __forceinline__ void tc_teardown(const TcLayout& L, uint32_t tmem_base) {
  tc_fence_before();
  __syncthreads();
  if ((threadIdx.x >> 5) == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)L.tmem_cols));
  }
}

// optional per-CTA phase trace (debug aid for tools/trace_cta.py): 8 x u64 nanosecond timestamps per CTA
__device__ unsigned long long* g_tc_trace = nullptr;
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void trace(int slot) {
  unsigned long long* tr = g_tc_trace;
  if (tr) tr[((blockIdx.z * gridDim.y + blockIdx.y) * (size_t)gridDim.x + blockIdx.x) * 8 + slot] = gtime();
}

// One 64-point x 64-channel activation box (MN-major SW128 block).  Row-tap mode (ConvArgs::rt_taps): virtual channel k0 = (tap,
// c0); the box is the one of the horizontal-tap expansion [P | rt_C | B] moved by whole rows in the flattened point index - a
// multiple of 8 elements, so the 16-byte rule on the innermost box start holds - and the map's own bounds [0, P) zero-fill the
// rows above / below the image.  (A 4-D [W | H | C | B] box does not work for W < 64: under SWIZZLE_128B every inner row of the
// box takes a full 128-byte row of shared memory, tools/tma_box_probe.cu.)
__device__ __forceinline__ void tma_load_xbox(const ConvArgs& a, void* dst, const CUtensorMap* map, int p, int k0, int b, uint64_t* bar) {
  if (a.rt_taps) {
    const int tap = k0 / a.rt_C, c0 = k0 - tap * a.rt_C;
    tma_load_3d(dst, map, p + (tap - (a.rt_taps >> 1)) * a.rt_dil * a.rt_W, c0, b, bar);
  } else {
    tma_load_3d(dst, map, p, k0, b, bar);
  }
}
// host side: the activation map of the TMA kernels, [P | C | B] (row-tap mode: C = the channels of the expansion, not the virtual K)
static int encode_xmap(CUtensorMap* tm, const ConvArgs& a) {
  cuuint64_t dims[3] = {(cuuint64_t)a.P_in, (cuuint64_t)(a.rt_taps ? a.rt_C : a.C0), (cuuint64_t)a.B};
  cuuint64_t strides[2] = {(cuuint64_t)a.P_in * 2, (cuuint64_t)a.src0_bstride * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)TC_BK, 1};
  return tma_encode(tm, VRCOC_BF16, a.src0, 3, dims, strides, box, true);
}

// ---- kernel 1: both operands by TMA, warp-specialised -----------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS, 3) conv_tc_tma_kernel(ConvArgs a, TcLayout L, const __grid_constant__ CUtensorMap tmapA,
                                                                 const __grid_constant__ CUtensorMap tmapB,
                                                                 const __grid_constant__ CUtensorMap tmapR,
                                                                 const __grid_constant__ CUtensorMap tmapO) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const TcSmem S = carve(smem_raw, L);
  const int tid = threadIdx.x;
  const int b = blockIdx.z, p0 = blockIdx.x * TC_BM, n0 = blockIdx.y * L.n_tile;
  if (tid == 64) {
    trace(0);
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmapA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmapB)) : "memory");
  }
  const uint32_t tmem_base = tc_setup(a, L, S, n0);
  if (tid == 64) trace(1);
  if (tid == 96 && L.res_tma) {                       // residual tile: in flight during the whole main loop
    mbar_expect_tx(S.bar_res, (uint32_t)(L.n_tile * TC_BM * 2));
    tma_load_3d(S.res_tile, &tmapR, p0, n0, b, S.bar_res);
  }
  const int nk = (a.K + TC_BK - 1) / TC_BK;
  const int NS = L.stages;
  if (tid == 0) {
    // producer: keeps the TMA ring full
    for (int kc = 0; kc < nk; ++kc) {
      const int s = kc % NS;
      if (kc >= NS) mbar_wait(&S.bar_free[s], (uint32_t)((kc / NS) - 1) & 1);
      mbar_expect_tx(&S.bar_full[s], (uint32_t)(TC_A_BYTES + L.b_bytes));
      unsigned char* As = S.sA + s * TC_A_BYTES;
      tma_load_xbox(a, As, &tmapA, p0, kc * TC_BK, b, &S.bar_full[s]);
      tma_load_xbox(a, As + TC_A_LBO, &tmapA, p0 + 64, kc * TC_BK, b, &S.bar_full[s]);
      tma_load_2d(S.sB + s * L.b_bytes, &tmapB, kc * TC_BK, n0, &S.bar_full[s]);
    }
  } else if (tid == 32) {
    // MMA issuer
    const uint32_t idesc = make_idesc(L.n_tile);
    int s = 0;
    uint32_t ph = 0;
    for (int kc = 0; kc < nk; ++kc) {
      mbar_wait(&S.bar_full[s], ph);
      if (kc == 0) trace(2);
      tc_fence_after();
      const int ksteps = (min(TC_BK, a.K - kc * TC_BK) + 15) >> 4;
      issue_slab_mmas(tmem_base, smem_u32(S.sA + s * TC_A_BYTES), smem_u32(S.sB + s * L.b_bytes), idesc, ksteps, kc == 0);
      tc_commit(&S.bar_free[s]);
      if (kc == nk - 1) { tc_commit(S.bar_acc); trace(3); }
      if (++s == NS) { s = 0; ph ^= 1u; }
    }
  }
  __syncwarp();
  tc_epilogue(a, L, S, tmem_base, p0, n0, b, &tmapO);
  if (tid == 64) trace(4);
  tc_teardown(L, tmem_base);
  if (tid == 64) trace(5);
}

// ---- kernel 2: transform-on-load A ------------------------------------------------------------------------------------------
// raw (untransformed) slab data of one thread: 4 k-rows x 8 points, kept in the source dtype so that no instruction
// depends on the loads until the next iteration's transform (the loads stay in flight across the MMA issue)
template <typename TS>
struct __align__(16) RawSlab {
  TS v[32];
  uint32_t mask;   // bit (i*8+j) = element (row i, point j) is a real in-range input
};

template <typename TS, bool FAST>
__device__ __forceinline__ void slab_gload(const ConvArgs& a, int b, int kc, int a_krow0, int q0, int P, int taps, RawSlab<TS>& raw) {
  raw.mask = 0;
  // q0 is a multiple of 8: when W_out % 8 == 0 (every live map) the thread's 8 points share one output row
  const int oy0 = FAST ? 0 : q0 / a.W_out, ox0 = FAST ? 0 : q0 - oy0 * a.W_out;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int kk = kc * TC_BK + a_krow0 + 16 * i;
    if (kk >= a.K || q0 >= P) continue;
    int c = kk, tap = 0;
    if (!FAST) {
      if (a.k_order) { tap = kk / a.Cin; c = kk - tap * a.Cin; }    // tap-major: the slab's rows are consecutive channels
      else { c = kk / taps; tap = kk - c * taps; }
    }
    const int sc = a.chan_src ? __ldg(a.chan_src + c) : c;
    const TS* src;
    if (sc < a.C0) src = reinterpret_cast<const TS*>(a.src0) + (int64_t)b * a.src0_bstride + (int64_t)sc * a.P_in;
    else src = reinterpret_cast<const TS*>(a.src1) + (int64_t)b * a.src1_bstride + (int64_t)(sc - a.C0) * a.P_in;
    if (FAST) {
      raw.mask |= 0xFFu << (8 * i);
      if (sizeof(TS) == 2) {
        reinterpret_cast<uint4*>(raw.v)[i] = __ldg(reinterpret_cast<const uint4*>(src + q0));
      } else {
        reinterpret_cast<float4*>(raw.v)[2 * i] = __ldg(reinterpret_cast<const float4*>(src + q0));
        reinterpret_cast<float4*>(raw.v)[2 * i + 1] = __ldg(reinterpret_cast<const float4*>(src + q0) + 1);
      }
    } else {
      const int ky = tap / a.kw, kx = tap - ky * a.kw;
      int oy = oy0, ox = ox0;
      int iy = oy * a.stride - a.pad + ky * a.dil;
      const TS* row = src + (int64_t)iy * a.W_in;
      int ix = ox * a.stride - a.pad + kx * a.dil;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (q0 + j < P && iy >= 0 && iy < a.H_in && ix >= 0 && ix < a.W_in) {
          raw.v[8 * i + j] = row[ix];
          raw.mask |= 1u << (8 * i + j);
        }
        ix += a.stride;
        if (++ox == a.W_out) {                       // only when W_out % 8 != 0
          ox = 0; ++oy;
          iy = oy * a.stride - a.pad + ky * a.dil;
          row = src + (int64_t)iy * a.W_in;
          ix = -a.pad + kx * a.dil;
        }
      }
    }
  }
}

template <typename TS, bool FAST, bool GATE>
__device__ __forceinline__ void slab_sstore(int kc, int ksteps, int a_krow0, int a_blk, int a_c, int taps, int k_order, int cin,
                                            const float4* tab, const RawSlab<TS>& raw, unsigned char* As) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = a_krow0 + 16 * i;
    if (k >= ksteps * 16) continue;
    const int kk = kc * TC_BK + k;
    float v[8];
    const uint32_t mk = (raw.mask >> (8 * i)) & 0xFFu;
    if (mk == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = 0.f;
    } else {
      if (FAST && sizeof(TS) == 2) {
        const uint4 r = reinterpret_cast<const uint4*>(raw.v)[i];
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
        for (int j = 0; j < 4; ++j) { float2 f = __bfloat1622float2(h[j]); v[2 * j] = f.x; v[2 * j + 1] = f.y; }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = ((mk >> j) & 1u) ? (float)raw.v[8 * i + j] : 0.f;
      }
      const float4 t = tab[FAST ? kk : (k_order ? kk % cin : kk / taps)];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float x = v[j];
        float y = fmaf(x, t.x, t.y);
        if (GATE) y *= sigmoid_fast(fmaf(t.z, x, t.w));
        v[j] = (FAST || ((mk >> j) & 1u)) ? y : 0.f;
      }
    }
    *reinterpret_cast<uint4*>(As + a_blk * TC_A_LBO + k * 128 + ((a_c ^ (k & 7)) << 4)) = pack8_bf16(v);
  }
}

template <typename TS, bool FAST>
__global__ void __launch_bounds__(TC_THREADS, 3) conv_tc_xform_kernel(ConvArgs a, TcLayout L, const __grid_constant__ CUtensorMap tmapB,
                                                                   const __grid_constant__ CUtensorMap tmapR,
                                                                   const __grid_constant__ CUtensorMap tmapO) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const TcSmem S = carve(smem_raw, L);
  const int tid = threadIdx.x;
  const int b = blockIdx.z, p0 = blockIdx.x * TC_BM, n0 = blockIdx.y * L.n_tile;
  const int P = a.P_out;
  const int taps = a.kh * a.kw;
  const int NS = L.stages;
  if (tid == 64 && L.use_tma_b) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmapB)) : "memory");
  build_prologue_table(a, b, S.tab);
  const uint32_t tmem_base = tc_setup(a, L, S, n0);
  if (tid == 96 && L.res_tma) {
    mbar_expect_tx(S.bar_res, (uint32_t)(L.n_tile * TC_BM * 2));
    tma_load_3d(S.res_tile, &tmapR, p0, n0, b, S.bar_res);
  }

  const int nk = (a.K + TC_BK - 1) / TC_BK;
  const uint32_t idesc = make_idesc(L.n_tile);
  // A loader geometry: thread -> (8-point chunk, k-rows krow0 + 16*i)
  const int a_chunk = tid & 15;            // points a_chunk*8 .. +8 of the tile
  const int a_krow0 = tid >> 4;            // 0..15
  const int a_blk = a_chunk >> 3;          // 64-point block
  const int a_c = a_chunk & 7;             // 16-byte chunk inside the 128-byte row
  const int q0 = p0 + a_chunk * 8;

  RawSlab<TS> raw;
  slab_gload<TS, FAST>(a, b, 0, a_krow0, q0, P, taps, raw);

  for (int kc = 0; kc < nk; ++kc) {
    const int s = kc % NS;
    const uint32_t use = (uint32_t)(kc / NS);
    if (kc >= NS) mbar_wait(&S.bar_free[s], (use - 1) & 1);     // MMAs that read this stage have completed
    const int ksteps = (min(TC_BK, a.K - kc * TC_BK) + 15) >> 4;
    unsigned char* As = S.sA + s * TC_A_BYTES;
    unsigned char* Bs = S.sB + s * L.b_bytes;

    if (L.use_tma_b) {
      if (tid == 0) {
        mbar_expect_tx(&S.bar_full[s], (uint32_t)L.b_bytes);
        tma_load_2d(Bs, &tmapB, kc * TC_BK, n0, &S.bar_full[s]);
      }
    } else {
      // tiny / unaligned K: the threads stage B themselves
      const int cpr = ksteps * 2;                            // 16-byte chunks per row
      const int units = L.n_tile * cpr;
      const __nv_bfloat16* W = reinterpret_cast<const __nv_bfloat16*>(a.weight);
      for (int u = tid; u < units; u += TC_THREADS) {
        const int n = u / cpr, c = u - n * cpr;
        const int o = n0 + n;
        const int kk = kc * TC_BK + c * 8;
        __align__(16) __nv_bfloat16 tmp[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) tmp[j] = (o < a.O && kk + j < a.K) ? W[(int64_t)o * a.K + kk + j] : __float2bfloat16_rn(0.f);
        *reinterpret_cast<uint4*>(Bs + (n >> 3) * 1024 + (n & 7) * 128 + ((c ^ (n & 7)) << 4)) = *reinterpret_cast<uint4*>(tmp);
      }
    }

    // A slab: transform the prefetched raw data and store it in the UMMA layout
    if (a.has_gate) slab_sstore<TS, FAST, true>(kc, ksteps, a_krow0, a_blk, a_c, taps, a.k_order, a.Cin, S.tab, raw, As);
    else slab_sstore<TS, FAST, false>(kc, ksteps, a_krow0, a_blk, a_c, taps, a.k_order, a.Cin, S.tab, raw, As);
    fence_async_smem();          // generic-proxy smem writes -> visible to the tensor core (async proxy)
    __syncthreads();

    if (tid == 0) {
      if (L.use_tma_b) mbar_wait(&S.bar_full[s], use & 1);
      tc_fence_after();
      issue_slab_mmas(tmem_base, smem_u32(As), smem_u32(Bs), idesc, ksteps, kc == 0);
      tc_commit(&S.bar_free[s]);                     // stage s reusable once these MMAs have read it
      if (kc == nk - 1) tc_commit(S.bar_acc);
    }
    // prefetch the next slab's raw data: in flight while the tensor core works on this slab
    if (kc + 1 < nk) slab_gload<TS, FAST>(a, b, kc + 1, a_krow0, q0, P, taps, raw);
  }

  tc_epilogue(a, L, S, tmem_base, p0, n0, b, &tmapO);
  tc_teardown(L, tmem_base);
}

// ---- kernel 3: persistent GroupNorm/table-prologue projection ------------------------------------------------------------------
// For 1x1 projections with a prologue and K <= 384 (GroupNorm -> fc1|fc_v, GroupNorm -> mlp.fc1 + GELU): one CTA owns one
// 128-point tile and a range of output channels.
//   * threads 0..255 build the normalised bf16 A operand ONCE (all K slabs resident in shared memory), then become the 8
//     epilogue warps;
//   * warp 8 lane 0 streams weight slabs by TMA through a 4-deep ring (starts before A is ready);
//   * warp 9 lane 0 issues tcgen05.mma for N tile j into TMEM buffer j&1 while the epilogue warps drain buffer (j-1)&1:
//     the activation epilogue (the issue-bound part, profiles/) overlaps the tensor work and the weight traffic, and the A
//     transform is not repeated per N tile as in conv_tc_xform_kernel.
constexpr int TP_THREADS = 320;
constexpr int TP_STAGES = 4;

struct TpLayout {
  int nt;            // N tile width (TMEM buffer = nt columns, two buffers)
  int tiles;         // N tiles per CTA
  int n_range;       // output channels per CTA = tiles * nt (last CTA / tile clipped by O)
  int nslabs;        // K slabs (resident A)
  int tmem_cols;
  int b_bytes;       // nt * 128
  int off_a, off_b, off_tab, off_epi, off_bar, total;
};

template <typename TS>
__global__ void __launch_bounds__(TP_THREADS, 2) conv_tc_persist_kernel(ConvArgs a, TpLayout L, const __grid_constant__ CUtensorMap tmapB) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  unsigned char* smem = smem_raw + pad;
  unsigned char* sA = smem + L.off_a;                                // [nslabs][16 KB]
  unsigned char* sB = smem + L.off_b;                                // [TP_STAGES][b_bytes]
  float4* tab = reinterpret_cast<float4*>(smem + L.off_tab);
  float* epi = reinterpret_cast<float*>(smem + L.off_epi);           // [2][n_range]: scale, shift
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + L.off_bar);
  uint64_t* bar_free = bar_full + TP_STAGES;
  uint64_t* acc_full = bar_free + TP_STAGES;                         // [2]
  uint64_t* acc_empty = acc_full + 2;                                // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.z, p0 = blockIdx.x * TC_BM;
  const int n_begin = blockIdx.y * L.n_range;
  const int P = a.P_out;
  int tiles = L.tiles;
  if (n_begin + tiles * L.nt > a.O) tiles = (a.O - n_begin + L.nt - 1) / L.nt;

  if (warp == 9) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)L.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 256) {
    for (int i = 0; i < TP_STAGES; ++i) { mbar_init(&bar_full[i], 1); mbar_init(&bar_free[i], 1); }
    mbar_init(&acc_full[0], 1); mbar_init(&acc_full[1], 1);
    mbar_init(&acc_empty[0], 8); mbar_init(&acc_empty[1], 8);      // one arrival per epilogue warp
    mbar_fence_init();
    tma_prefetch_desc(&tmapB);
  }
  if (tid == 0) trace(0);
  build_prologue_table(a, b, tab);                                     // strides by blockDim.x: every thread takes part
  if (tid < 256) {
    for (int n = tid; n < L.n_range; n += 256) {
      const int o = n_begin + n;
      const bool in = o < a.O;
      epi[n] = (in && a.e_scale) ? a.e_scale[o] : 1.f;
      epi[L.n_range + n] = (in && a.e_shift) ? a.e_shift[o] : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();                                                     // barriers + TMEM slot + tables visible
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int nk = L.nslabs;

  if (warp == 8) {
    // ---- weight producer --------------------------------------------------------------------------------------------
    // the first TP_STAGES slabs need no free-slot wait and are issued before (A), i.e. while the other warps still build
    // the A operand; everything after that depends on MMA progress and therefore has to come after the barrier
    const int total = tiles * nk;
    auto issue = [&](int it) {
      const int s = it % TP_STAGES;
      const int j = it / nk, kc = it - j * nk;
      if (it >= TP_STAGES) mbar_wait(&bar_free[s], (uint32_t)((it / TP_STAGES) - 1) & 1);
      mbar_expect_tx(&bar_full[s], (uint32_t)L.b_bytes);
      tma_load_2d(sB + s * L.b_bytes, &tmapB, kc * TC_BK, n_begin + j * L.nt, &bar_full[s]);
    };
    if (lane == 0)
      for (int it = 0; it < total && it < TP_STAGES; ++it) issue(it);
    __syncwarp();
    __syncthreads();                                                   // (A) the A operand is complete
    if (lane == 0)
      for (int it = TP_STAGES; it < total; ++it) issue(it);
    __syncwarp();
  } else if (warp == 9) {
    __syncthreads();                                                   // (A)
    // ---- MMA issuer ---------------------------------------------------------------------------------------------------
    if (lane == 0) {
      tc_fence_after();
      const uint32_t idesc = make_idesc(L.nt);
      int it = 0;
      for (int j = 0; j < tiles; ++j) {
        const int buf = j & 1;
        if (j >= 2) { mbar_wait(&acc_empty[buf], (uint32_t)((j >> 1) - 1) & 1); tc_fence_after(); }
        const uint32_t tacc = tmem_base + (uint32_t)(buf * L.nt);
        for (int kc = 0; kc < nk; ++kc, ++it) {
          const int s = it % TP_STAGES;
          mbar_wait(&bar_full[s], (uint32_t)(it / TP_STAGES) & 1);
          tc_fence_after();
          const int ksteps = (min(TC_BK, a.K - kc * TC_BK) + 15) >> 4;
          issue_slab_mmas(tacc, smem_u32(sA + kc * TC_A_BYTES), smem_u32(sB + s * L.b_bytes), idesc, ksteps, kc == 0);
          tc_commit(&bar_free[s]);
        }
        tc_commit(&acc_full[buf]);
      }
    }
    __syncwarp();
  } else {
    // ---- A operand: transform on load, all K slabs resident -----------------------------------------------------------------
    const int a_chunk = tid & 15, a_krow0 = tid >> 4, a_blk = a_chunk >> 3, a_c = a_chunk & 7;
    const int q0 = p0 + a_chunk * 8;
    RawSlab<TS> raw;
    if (tid == 0) trace(1);
    slab_gload<TS, true>(a, b, 0, a_krow0, q0, P, 1, raw);
    for (int kc = 0; kc < nk; ++kc) {
      const int ksteps = (min(TC_BK, a.K - kc * TC_BK) + 15) >> 4;
      RawSlab<TS> cur = raw;
      if (kc + 1 < nk) slab_gload<TS, true>(a, b, kc + 1, a_krow0, q0, P, 1, raw);
      if (a.has_gate) slab_sstore<TS, true, true>(kc, ksteps, a_krow0, a_blk, a_c, 1, 0, a.Cin, tab, cur, sA + kc * TC_A_BYTES);
      else slab_sstore<TS, true, false>(kc, ksteps, a_krow0, a_blk, a_c, 1, 0, a.Cin, tab, cur, sA + kc * TC_A_BYTES);
    }
    fence_async_smem();
    if (tid == 0) trace(2);
    __syncthreads();                                                   // (A)

    // ---- epilogue warps ---------------------------------------------------------------------------------------------------------
    const int lq = warp & 3, chalf = warp >> 2;
    const int q = p0 + lq * 32 + lane;
    const bool valid = q < P;
    const bool all_valid = __all_sync(0xffffffffu, valid);
    const int ncols = L.nt >> 1;
    for (int j = 0; j < tiles; ++j) {
      const int buf = j & 1;
      mbar_wait(&acc_full[buf], (uint32_t)(j >> 1) & 1);
      tc_fence_after();
      if (tid == 0 && j == 0) trace(3);
      const uint32_t tbase = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(buf * L.nt);
      const int nloc = j * L.nt;                                       // tile offset inside the CTA's range
      int c_end = (chalf + 1) * ncols;
      if (c_end > a.O - n_begin - nloc) c_end = a.O - n_begin - nloc;
      for (int c0 = chalf * ncols; c0 < c_end; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(tbase + (uint32_t)c0, r);
        const int o0 = n_begin + nloc + c0;
        const int lim = min(16, a.O - o0);
        const bool fast = lim == 16 && all_valid;
        const bool second = o0 >= a.O_split;
        const int odt = second ? a.out2_dtype : a.out_dtype;
        void* obase = second ? a.out2 : a.out;
        const int64_t oidx = (second ? ((int64_t)b * (a.O - a.O_split) + (o0 - a.O_split)) : ((int64_t)b * a.O_split + o0)) * P + q;
        const float* es = epi + nloc + c0;
        const float* eh = epi + L.n_range + nloc + c0;
#define TP_EPI(ACTV)                                                                                                             \
  if (odt == VRCOC_BF16) epi_plain_group<ACTV, __nv_bfloat16>(r, es, eh, reinterpret_cast<__nv_bfloat16*>(obase) + oidx, P, lim, valid, fast); \
  else epi_plain_group<ACTV, float>(r, es, eh, reinterpret_cast<float*>(obase) + oidx, P, lim, valid, fast)
        switch (a.act) {
          case VRCOC_ACT_NONE: TP_EPI(VRCOC_ACT_NONE); break;
          case VRCOC_ACT_RELU: TP_EPI(VRCOC_ACT_RELU); break;
          case VRCOC_ACT_GELU: TP_EPI(VRCOC_ACT_GELU); break;
          case VRCOC_ACT_SILU: TP_EPI(VRCOC_ACT_SILU); break;
          default: TP_EPI(VRCOC_ACT_LRELU); break;
        }
#undef TP_EPI
      }
      // this warp is done reading TMEM buffer `buf`
      tc_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&acc_empty[buf])) : "memory");
    }
    if (tid == 0) trace(4);
  }
  tc_fence_before();
  __syncthreads();
  if (tid == 0) trace(5);
  if (warp == 9) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)L.tmem_cols));
  }
}

}  // namespace vrcoc
#include "conv_tc_cm.cuh"
#include "mlp_fused.cuh"
namespace vrcoc {

// ---- host side ---------------------------------------------------------------------------------------------------------
static bool tma_a_eligible(const ConvArgs& a) {
  return a.fast1x1 && a.src0_dtype == VRCOC_BF16 && !a.gn_sums && !a.table && !a.chan_src && a.C1 == 0 && (a.K % 8 == 0) &&
         ((reinterpret_cast<uintptr_t>(a.weight) & 15) == 0) && tma_encode_fn() != nullptr;
}

static bool res_tma_eligible(const ConvArgs& a) {
  return a.res && a.res_dtype == VRCOC_BF16 && a.out_dtype == VRCOC_BF16 && a.O_split == a.O && a.P_out % 8 == 0 &&
         ((reinterpret_cast<uintptr_t>(a.res) | reinterpret_cast<uintptr_t>(a.out)) & 15) == 0 && tma_encode_fn() != nullptr &&
         (a.act == VRCOC_ACT_NONE || a.act == VRCOC_ACT_RELU || a.act == VRCOC_ACT_GELU);
}

static int tc_smem_bytes(const ConvArgs& a, int n_tile, int stages) {
  return stages * (TC_A_BYTES + n_tile * 128) + a.Cin * 16 + 5 * n_tile * 4 + (res_tma_eligible(a) ? n_tile * TC_BM * 2 : 0) +
         (2 * TC_MAX_STAGES + 2) * 8 + 64 + 1024 + 128;
}

static int tc_resident(const ConvArgs& a, int n_tile, int stages) {
  int tmem = 32;
  while (tmem < n_tile) tmem *= 2;
  int by_tmem = 512 / tmem, by_smem = (227 * 1024) / tc_smem_bytes(a, n_tile, stages), by_regs = 3;
  int r = by_tmem < by_smem ? by_tmem : by_smem;
  return r < by_regs ? (r < 1 ? 1 : r) : by_regs;
}

static TcLayout tc_layout(const ConvArgs& a) {
  TcLayout L{};
  const int64_t m_tiles = cdiv(a.P_out, TC_BM) * a.B;
  const int nk = (a.K + TC_BK - 1) / TC_BK;
  const int sms = sm_count();
  // N tiling by a small cost model (measured on B200, profiles/): a CTA costs a fixed part + one slab latency per 64 k
  // + an epilogue part per output column; the launch costs that times the number of waves over the resident-CTA slots,
  // which depend on the tile through TMEM columns and shared memory.  Narrow tiles re-do the A transform more often
  // but keep more CTAs resident (their phases overlap) and quantise better; wide tiles win when there are many point tiles.
  const bool heavy_act = a.act == VRCOC_ACT_GELU || a.act == VRCOC_ACT_SILU;
  double best = 1e30;
  int best_tile = 32, best_stages = 1;
  for (int nt = 32; nt <= 256; nt += 32) {
    const int64_t n_tiles = cdiv(a.O, nt);
    if (nt > 32 && n_tiles * (nt - 32) >= a.O) continue;          // a narrower tile covers O with the same count
    int stages = nk < TC_MAX_STAGES ? nk : TC_MAX_STAGES;
    while (stages > 2 && tc_resident(a, nt, stages) < tc_resident(a, nt, 2)) --stages;
    if (tc_smem_bytes(a, nt, stages) > 220 * 1024) continue;
    const int64_t slots = (int64_t)sms * tc_resident(a, nt, stages);
    const int64_t waves = cdiv(m_tiles * n_tiles, slots);
    const double cta = 2.0 + 1.0 * nk + (heavy_act ? 0.040 : 0.022) * nt;
    const double cost = (double)waves * cta;
    if (cost < best - 1e-9) { best = cost; best_tile = nt; best_stages = stages; }
  }
  L.n_tile = best_tile;
  L.stages = best_stages;
  L.tmem_cols = 32;
  while (L.tmem_cols < L.n_tile) L.tmem_cols *= 2;
  L.b_bytes = L.n_tile * 128;
  L.use_tma_b = (a.K % 8 == 0) && ((reinterpret_cast<uintptr_t>(a.weight) & 15) == 0) && tma_encode_fn() != nullptr;
  L.plain_epi = !a.res && !a.post_scale && !a.f_scale && !a.f_shift && !a.out_sample_sums && !a.out_minmax;
  L.off_b = L.stages * TC_A_BYTES;
  L.off_tab = L.off_b + L.stages * L.b_bytes;
  L.off_epi = L.off_tab + a.Cin * 16;
  L.res_tma = res_tma_eligible(a) ? 1 : 0;
  L.off_res = (L.off_epi + 5 * L.n_tile * 4 + 127) & ~127;
  L.off_bar = L.off_res + (L.res_tma ? L.n_tile * TC_BM * 2 : 0);
  L.total = L.off_bar + (2 * TC_MAX_STAGES + 2) * 8 + 16 + 1024;   // + alignment slack
  return L;
}

static TpLayout tp_layout(const ConvArgs& a) {
  TpLayout T{};
  T.nslabs = (a.K + TC_BK - 1) / TC_BK;
  T.nt = a.O >= 128 ? 128 : (int)cdiv(a.O, 32) * 32;
  const int64_t m_tiles = cdiv(a.P_out, TC_BM) * a.B;
  const int n_tiles_total = (int)cdiv(a.O, T.nt);
  // resident CTAs per SM by shared memory / TMEM (2 x nt columns each)
  const int smem1 = T.nslabs * TC_A_BYTES + TP_STAGES * T.nt * 128 + a.Cin * 16 + 2048;
  int res = (220 * 1024) / (smem1 + 2 * n_tiles_total * T.nt * 4);
  if (res > 2) res = 2;
  if (res < 1) res = 1;
  const int64_t slots = (int64_t)sm_count() * res;
  // split the output channels over as many CTAs as it takes to fill the chip once, but never re-build A more than needed
  int splits = (int)cdiv(slots, m_tiles);
  if (splits > n_tiles_total) splits = n_tiles_total;
  if (splits < 1) splits = 1;
  T.tiles = (int)cdiv(n_tiles_total, splits);
  T.n_range = T.tiles * T.nt;
  T.tmem_cols = 32;
  while (T.tmem_cols < 2 * T.nt) T.tmem_cols *= 2;
  T.b_bytes = T.nt * 128;
  T.off_a = 0;
  T.off_b = T.nslabs * TC_A_BYTES;
  T.off_tab = T.off_b + TP_STAGES * T.b_bytes;
  T.off_epi = T.off_tab + a.Cin * 16;
  T.off_bar = (T.off_epi + 2 * T.n_range * 4 + 15) & ~15;
  T.total = T.off_bar + (2 * TP_STAGES + 4) * 8 + 16 + 1024;
  return T;
}

static bool persist_eligible(const ConvArgs& a, const TcLayout& L) {
  if (!a.fast1x1 || !L.plain_epi || !L.use_tma_b || a.chan_src || a.C1 != 0) return false;
  if (!(a.gn_sums || a.table)) return false;             // prologue-free projections take the TMA-only kernel
  if (a.K > 384 || a.O < 32) return false;
  return tp_layout(a).total <= 220 * 1024;
}

// channel-major kernel: operand sourcing, ring depth and output tiles per CTA from a small cost model (time unit ~ us)
struct CmPlan { TqLayout T; int xmode; double cost; };

static bool cm_plan(const ConvArgs& a, CmPlan& best) {
  if (!a.fast1x1 || !a.vec_out || a.chan_src) return false;
  // a second source (virtual concat, RadarEnhanceByImage) only through the TMA + in-place prologue mode: bf16, the boundary
  // on a k-slab, and no second output (its tensor-map slot carries the second source)
  const bool two_src = a.C1 != 0;
  if (two_src && (!(a.table || a.gn_sums || a.has_gate) || a.src1_dtype != VRCOC_BF16 || a.src0_dtype != VRCOC_BF16 || (a.C0 % TC_BK) != 0 ||
                  a.src1_bstride == 0 || (a.src1_bstride % 8) != 0 || (reinterpret_cast<uintptr_t>(a.src1) & 15) != 0 || a.O_split != a.O))
    return false;
  if (a.weight_dtype != VRCOC_BF16 || (a.K % 8) != 0 || (reinterpret_cast<uintptr_t>(a.weight) & 15) != 0 || !tma_encode_fn()) return false;
  // outputs / residual travel as TMA boxes of 32 channels: 16-byte aligned bases, the split on a 32-channel boundary,
  // bf16 residual, and no residual into fp32 outputs (the staging region holds one or the other)
  if (a.O_split != a.O && (a.O_split % 32) != 0) return false;
  if ((reinterpret_cast<uintptr_t>(a.out) & 15) != 0 || (a.O_split != a.O && (reinterpret_cast<uintptr_t>(a.out2) & 15) != 0)) return false;
  if (a.res && (a.res_dtype != VRCOC_BF16 || (reinterpret_cast<uintptr_t>(a.res) & 15) != 0 || a.out_dtype != VRCOC_BF16 ||
                (a.O_split != a.O && a.out2_dtype != VRCOC_BF16)))
    return false;
  const bool fold = a.gn_fold_k1 != nullptr;                // GroupNorm folded into [hi | lo] weights: raw X by TMA, no prologue
  const bool prologue = !fold && (a.gn_sums || a.table || a.has_gate);
  const bool tma_x = !prologue && a.src0_dtype == VRCOC_BF16 && (reinterpret_cast<uintptr_t>(a.src0) & 15) == 0 &&
                     (a.src0_bstride % 8) == 0;
  if (fold && (!tma_x || two_src)) return false;
  const int nx = (a.K + TC_BK - 1) / TC_BK;                 // X slabs
  const int nslabs = fold ? 2 * nx : nx;                    // weight slabs of a (feat) tile
  const int n_tiles = (int)cdiv(a.O, TQ_MT);
  const int64_t m_tiles = cdiv(a.P_out, TQ_NP) * a.B;
  const bool plain = !(a.post_scale || a.res || a.f_scale || a.f_shift || a.out_sample_sums || a.out_minmax);
  // measured (tools/microbench.py, B200): with both operands by TMA the point-major kernel keeps 3 CTAs per SM and wins
  // whenever the last 128-channel tile would be partly empty (O = 64, 320); the channel-major kernel wins on full tiles
  if (tma_x && !fold && (a.O % TQ_MT) != 0) return false;
  const double epi_us = a.act == VRCOC_ACT_GELU || a.act == VRCOC_ACT_SILU ? 1.6 : (plain ? 0.5 : 0.8);
  bool found = false;
  const int modes[10] = {1, 1, 1, 1, 1, 0, 0, 0, 0, 0}, depth[10] = {2, 3, 4, 6, 8, 2, 3, 4, 5, 6};
  // a bf16 source with a prologue can still come in by TMA and be normalised in place (XMODE 3)
  const bool tma_x3 = prologue && a.src0_dtype == VRCOC_BF16 && (reinterpret_cast<uintptr_t>(a.src0) & 15) == 0 &&
                      (a.src0_bstride % 8) == 0 && nslabs <= TQ_MAX_SLABS3;
  if (two_src && !tma_x3) return false;
  for (int cand = 0; cand < (tma_x ? 10 : 5); ++cand) {
    const int bufs = 1;       // a second staging buffer was measured: no gain (the bulk-store drain is not on the critical path)
    const int mode = tma_x ? modes[cand] : (tma_x3 ? 3 : 2), st = depth[cand];
    const int x_bytes = mode == 0 ? st * TQ_X_BYTES : nx * TQ_X_BYTES;
    const int tab_bytes = mode >= 2 ? a.Cin * 16 : 0;
    const int total = x_bytes + st * TQ_W_BYTES + bufs * 8 * 4096 + tab_bytes + 512 + 1024;
    if (total > 220 * 1024) continue;
    const int res = (227 * 1024) / (total + 1024) >= 2 ? 2 : 1;
    const int64_t slots = (int64_t)sm_count() * res;
    const double x_us = mode == 2 ? 1.1 * nslabs : (mode == 3 ? 1.5 + 0.2 * nslabs : 0.6);
    // one k-slab: four MMAs (0.14 us) or, when the ring is shallow, the L2 -> shared latency of a TMA box (~1 us) / depth
    const double mma_bw_us = nslabs * (mode == 0 ? 0.28 : 0.14);    // bandwidth part: shared by co-resident CTAs
    const double mma_lat_us = nslabs * 1.0 / st;                     // latency part: every CTA has its own ring
    for (int t = 1; t <= n_tiles; ++t) {
      const int64_t ctas = m_tiles * cdiv(n_tiles, t);
      const int64_t waves = cdiv(ctas, slots);
      // co-resident CTAs share the SM's issue slots and fill bandwidth, but one's latency-bound phases (set-up, operand
      // build, pipeline fill) hide behind the other's epilogue: 1.6x, not 2x, and only on the steady-state part
      const double share = (res == 2 && ctas > (int64_t)sm_count()) ? 1.6 : 1.0;
      double tile_us = share * (mma_bw_us > epi_us ? mma_bw_us : epi_us);
      if (mma_lat_us > tile_us) tile_us = mma_lat_us;
      const double cost = waves * (2.0 + x_us + t * tile_us + (mma_bw_us < epi_us ? mma_bw_us : epi_us));
      if (!found || cost < best.cost - 1e-9) {
        found = true;
        best.cost = cost;
        best.xmode = mode;
        TqLayout& T = best.T;
        T.stages = st;
        T.nslabs = nslabs;
        T.nx = nx;
        T.tiles = t;
        T.plain = plain ? 1 : 0;
        T.off_x = 0;
        T.off_w = x_bytes;
        T.off_stage = T.off_w + st * TQ_W_BYTES;
        T.stage_bufs = bufs;
        T.off_tab = T.off_stage + bufs * 8 * 4096;
        T.off_bar = T.off_tab + tab_bytes;
        T.total = total;
      }
    }
  }
  return found;
}

bool conv_tc_supported(const ConvArgs& a) {
  if (a.weight_dtype != VRCOC_BF16) return false;
  if (a.C1 > 0 && a.src1_dtype != a.src0_dtype) return false;
  if (a.O_split != a.O && (a.O_split % 16) != 0) return false;
  TcLayout L = tc_layout(a);
  return L.total <= 220 * 1024;
}

static int encode(CUtensorMap* tm, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                  const cuuint32_t* box) {
  return tma_encode(tm, VRCOC_BF16, base, rank, dims, strides_bytes, box, true);
}

template <typename K>
static void set_smem(K kern, int bytes) {
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  // ask for the largest shared-memory carve-out: residency (CTAs/SM) must not depend on the driver's L1/smem heuristic
  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

static int g_tc_use_cm = 1;

}  // namespace vrcoc
// debug / A-B switch: 0 routes 1x1 projections to the point-major kernels (tools/microbench.py --no-cm)
extern "C" int vrcoc_debug_set_cm(int on) { vrcoc::g_tc_use_cm = on; return 0; }
extern "C" int vrcoc_debug_set_trace(unsigned long long* buf) {
  return cudaMemcpyToSymbol(vrcoc::g_tc_trace, &buf, sizeof(buf)) == cudaSuccess ? 0 : -2;
}
namespace vrcoc {

int launch_conv_tc(const ConvArgs& a, cudaStream_t st) {
  TcLayout L = tc_layout(a);
  VRCOC_REQUIRE(L.total <= 220 * 1024, "conv(tcgen05): shared memory budget exceeded (%d bytes)", L.total);
  VRCOC_REQUIRE(a.C1 == 0 || a.src1_dtype == a.src0_dtype, "conv(tcgen05): both sources must share a dtype");
  CUtensorMap tmB, tmA, tmR, tmO;
  memset(&tmB, 0, sizeof(tmB));
  memset(&tmA, 0, sizeof(tmA));
  memset(&tmR, 0, sizeof(tmR));
  memset(&tmO, 0, sizeof(tmO));
  if (L.res_tma) {
    // residual / output [B][O][P] bf16; box = 128 points x n_tile channels, dense
    cuuint64_t dims[3] = {(cuuint64_t)a.P_out, (cuuint64_t)a.O, (cuuint64_t)a.B};
    cuuint64_t strides[2] = {(cuuint64_t)a.P_out * 2, (cuuint64_t)a.O * a.P_out * 2};
    cuuint32_t box[3] = {(cuuint32_t)TC_BM, (cuuint32_t)L.n_tile, 1};
    int rc = tma_encode(&tmR, VRCOC_BF16, a.res, 3, dims, strides, box, false);
    if (rc) return rc;
    rc = tma_encode(&tmO, VRCOC_BF16, a.out, 3, dims, strides, box, false);
    if (rc) return rc;
  }
  if (L.use_tma_b) {
    // weights [O][K] bf16 row-major; box = 64 k (128 B) x n_tile rows, 128-byte swizzle, zero fill outside
    cuuint64_t dims[2] = {(cuuint64_t)a.K, (cuuint64_t)a.O};
    cuuint64_t strides[1] = {(cuuint64_t)a.K * 2};
    cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)L.n_tile};
    int rc = encode(&tmB, a.weight, 2, dims, strides, box);
    if (rc) return rc;
  }
  const bool f32 = a.src0_dtype == VRCOC_F32;
  CmPlan cm;
  if (g_tc_use_cm && cm_plan(a, cm)) {
    const TqLayout& T = cm.T;
    static const bool dbg = getenv("VRCOC_DEBUG_PLAN") != nullptr;
    if (dbg)
      fprintf(stderr, "[vrcoc] cm plan K=%d O=%d P=%d B=%d: xmode=%d stages=%d tiles=%d smem=%d cost=%.1f\n", a.K, a.O, a.P_out, a.B,
              cm.xmode, T.stages, T.tiles, T.total, cm.cost);
    dim3 cgrid((unsigned)cdiv(a.P_out, TQ_NP), (unsigned)cdiv(a.O, T.tiles * TQ_MT), (unsigned)a.B);
    CUtensorMap tmW, tmX, tmO1, tmO2, tmRes;
    memset(&tmX, 0, sizeof(tmX));
    memset(&tmO2, 0, sizeof(tmO2));
    memset(&tmRes, 0, sizeof(tmRes));
    {
      // outputs [B][C][P]: box = one 128-byte row segment (64 bf16 / 32 fp32 points) x 32 channels, 128-byte swizzle
      const int c1 = a.O_split, c2 = a.O - a.O_split;
      const int e1 = a.out_dtype == VRCOC_BF16 ? 2 : 4;
      cuuint64_t dims[3] = {(cuuint64_t)a.P_out, (cuuint64_t)c1, (cuuint64_t)a.B};
      cuuint64_t strides[2] = {(cuuint64_t)a.P_out * e1, (cuuint64_t)c1 * a.P_out * e1};
      cuuint32_t box[3] = {(cuuint32_t)(128 / e1), 32, 1};
      int rc = tma_encode(&tmO1, a.out_dtype, a.out, 3, dims, strides, box, true);
      if (rc) return rc;
      if (c2 > 0) {
        const int e2 = a.out2_dtype == VRCOC_BF16 ? 2 : 4;
        cuuint64_t dims2[3] = {(cuuint64_t)a.P_out, (cuuint64_t)c2, (cuuint64_t)a.B};
        cuuint64_t strides2[2] = {(cuuint64_t)a.P_out * e2, (cuuint64_t)c2 * a.P_out * e2};
        cuuint32_t box2[3] = {(cuuint32_t)(128 / e2), 32, 1};
        rc = tma_encode(&tmO2, a.out2_dtype, a.out2, 3, dims2, strides2, box2, true);
        if (rc) return rc;
      }
      if (a.res) {
        cuuint64_t dimsr[3] = {(cuuint64_t)a.P_out, (cuuint64_t)a.O, (cuuint64_t)a.B};
        cuuint64_t stridesr[2] = {(cuuint64_t)a.P_out * 2, (cuuint64_t)a.O * a.P_out * 2};
        cuuint32_t boxr[3] = {64, 32, 1};
        rc = tma_encode(&tmRes, VRCOC_BF16, a.res, 3, dimsr, stridesr, boxr, true);
        if (rc) return rc;
      }
    }
    {
      const int wk = a.gn_fold_k1 ? 2 * a.K : a.K;                     // folded GroupNorm: rows are [hi | lo]
      cuuint64_t dims[2] = {(cuuint64_t)wk, (cuuint64_t)a.O};
      cuuint64_t strides[1] = {(cuuint64_t)wk * 2};
      cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)TQ_MT};
      int rc = encode(&tmW, a.weight, 2, dims, strides, box);
      if (rc) return rc;
    }
    if (cm.xmode != 2) {
      cuuint32_t box[3] = {64, (cuuint32_t)TC_BK, 1};
      int rc = encode_xmap(&tmX, a);
      if (rc) return rc;
      if (a.C1 > 0) {                                                  // second source, in the second-output slot (cm_plan)
        cuuint64_t dims1[3] = {(cuuint64_t)a.P_in, (cuuint64_t)a.C1, (cuuint64_t)a.B};
        cuuint64_t strides1[2] = {(cuuint64_t)a.P_in * 2, (cuuint64_t)a.src1_bstride * 2};
        rc = encode(&tmO2, a.src1, 3, dims1, strides1, box);
        if (rc) return rc;
      }
    }
#define LAUNCH_CM(TS, MODE)                                                                     \
  do {                                                                                          \
    set_smem(conv_tc_cm_kernel<TS, MODE>, T.total);                                             \
    conv_tc_cm_kernel<TS, MODE><<<cgrid, TQ_THREADS, T.total, st>>>(a, T, tmW, tmX, tmO1, tmO2, tmRes); \
  } while (0)
    if (cm.xmode == 0) LAUNCH_CM(__nv_bfloat16, 0);
    else if (cm.xmode == 1) LAUNCH_CM(__nv_bfloat16, 1);
    else if (cm.xmode == 3) LAUNCH_CM(__nv_bfloat16, 3);
    else if (f32) LAUNCH_CM(float, 2);
    else LAUNCH_CM(__nv_bfloat16, 2);
#undef LAUNCH_CM
    return check_launch("conv_tc_cm");
  }
  VRCOC_REQUIRE(!a.gn_fold_k1, "conv(tcgen05): the folded-GroupNorm projection needs the channel-major kernel (P %% 8 == 0, aligned "
                               "bf16 tensors, O_split %% 128 == 0, C0 %% 64 == 0); this problem is not covered");
  dim3 grid((unsigned)cdiv(a.P_out, TC_BM), (unsigned)cdiv(a.O, L.n_tile), (unsigned)a.B);
  if (tma_a_eligible(a)) {
    // activations [B][C][P] bf16; box = 64 points (128 B) x 64 channels, lands as one MN-major SW128 block
    int rc = encode_xmap(&tmA, a);
    if (rc) return rc;
    set_smem(conv_tc_tma_kernel, L.total);
    conv_tc_tma_kernel<<<grid, TC_THREADS, L.total, st>>>(a, L, tmA, tmB, tmR, tmO);
    return check_launch("conv_tc_tma");
  }
  VRCOC_REQUIRE(!a.rt_taps, "conv(tcgen05): row-tap mode reached a kernel without shifted TMA boxes (K %% 8, aligned bf16 weight needed)");
  if (persist_eligible(a, L)) {
    TpLayout T = tp_layout(a);
    dim3 pgrid((unsigned)cdiv(a.P_out, TC_BM), (unsigned)cdiv(a.O, T.n_range), (unsigned)a.B);
    // the weight box is nt rows here, not L.n_tile
    cuuint64_t dims[2] = {(cuuint64_t)a.K, (cuuint64_t)a.O};
    cuuint64_t strides[1] = {(cuuint64_t)a.K * 2};
    cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)T.nt};
    int rc = encode(&tmB, a.weight, 2, dims, strides, box);
    if (rc) return rc;
    if (f32) {
      set_smem(conv_tc_persist_kernel<float>, T.total);
      conv_tc_persist_kernel<float><<<pgrid, TP_THREADS, T.total, st>>>(a, T, tmB);
    } else {
      set_smem(conv_tc_persist_kernel<__nv_bfloat16>, T.total);
      conv_tc_persist_kernel<__nv_bfloat16><<<pgrid, TP_THREADS, T.total, st>>>(a, T, tmB);
    }
    return check_launch("conv_tc_persist");
  }
#define LAUNCH(TS, FASTV)                                                        \
  do {                                                                           \
    set_smem(conv_tc_xform_kernel<TS, FASTV>, L.total);                          \
    conv_tc_xform_kernel<TS, FASTV><<<grid, TC_THREADS, L.total, st>>>(a, L, tmB, tmR, tmO); \
  } while (0)
  if (a.fast1x1) { if (f32) LAUNCH(float, true); else LAUNCH(__nv_bfloat16, true); }
  else           { if (f32) LAUNCH(float, false); else LAUNCH(__nv_bfloat16, false); }
#undef LAUNCH
  return check_launch("conv_tc_xform");
}

}  // namespace vrcoc

// Is the folded-GroupNorm projection (vrcoc.h, gn_fold_k1) available for this shape?  Runs the channel-major planner on a
// synthetic problem (aligned pointers), so the host side can decide before it builds the [hi | lo] weights.
extern "C" int vrcoc_gn_fold_supported(int B, int C, int O, int O_split, int P) {
  using namespace vrcoc;
  if (B <= 0 || C <= 0 || (C % 64) != 0 || O <= 0 || O_split <= 0 || O_split > O || P <= 0 || (P % 8) != 0 || !tma_encode_fn()) return 0;
  ConvArgs a;
  memset(&a, 0, sizeof(a));
  a.B = B; a.C0 = a.Cin = a.K = C; a.O = O; a.O_split = O_split; a.P_in = a.P_out = P; a.kh = a.kw = a.stride = a.dil = 1;
  a.H_in = a.H_out = 1; a.W_in = a.W_out = P;
  void* fake = reinterpret_cast<void*>(uintptr_t(1) << 20);
  a.src0 = fake; a.src0_dtype = VRCOC_BF16; a.src0_bstride = (int64_t)C * P;
  a.weight = fake; a.weight_dtype = VRCOC_BF16;
  a.out = fake; a.out_dtype = VRCOC_F32; a.out2 = fake; a.out2_dtype = VRCOC_BF16;
  a.gn_sums = reinterpret_cast<const double*>(fake); a.gn_fold_k1 = reinterpret_cast<const float*>(fake);
  a.e_shift = reinterpret_cast<const float*>(fake);
  a.fast1x1 = 1; a.vec_out = 1;
  CmPlan cm;
  return cm_plan(a, cm) ? 1 : 0;
}

// ---- fused channel MLP ------------------------------------------------------------------------------------------------------
extern "C" int vrcoc_mlp_fused_supported(int dtype, int C, int hidden, int P) {
  return dtype == VRCOC_BF16 && C >= 64 && C % 64 == 0 && C <= 384 && hidden > 0 && hidden % vrcoc::TQ_MT == 0 && P > 0 && P % 8 == 0 &&
         vrcoc::tma_encode_fn() != nullptr;
}

extern "C" int vrcoc_mlp_fused_fwd(const void* x, const double* gn_sums, const float* gamma, const float* beta, float eps, const void* w1,
                                   const float* b1, const void* w2, const float* b2, const float* layer_scale, void* out,
                                   double* out_sample_sums, int B, int C, int hidden, int P, void* stream) {
  using namespace vrcoc;
  VRCOC_REQUIRE(x && gn_sums && gamma && beta && w1 && b1 && w2 && out && B > 0, "mlp_fused: bad argument");
  VRCOC_REQUIRE(vrcoc_mlp_fused_supported(VRCOC_BF16, C, hidden, P), "mlp_fused: unsupported shape C=%d hidden=%d P=%d", C, hidden, P);
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  VRCOC_REQUIRE(al(x) && al(w1) && al(w2) && al(out), "mlp_fused: pointers must be 16-byte aligned");
  ConvArgs a1, a2;
  memset(&a1, 0, sizeof(a1));
  memset(&a2, 0, sizeof(a2));
  a1.B = B; a1.C0 = a1.Cin = a1.K = C; a1.O = hidden; a1.P_in = a1.P_out = P;
  a1.src0 = x; a1.src0_dtype = VRCOC_BF16; a1.src0_bstride = (int64_t)C * P;
  a1.gn_sums = gn_sums; a1.gn_gamma = gamma; a1.gn_beta = beta; a1.gn_eps = eps;
  a2.B = B; a2.C0 = a2.Cin = a2.K = hidden; a2.O = a2.O_split = C; a2.P_in = a2.P_out = P;
  a2.e_shift = b2; a2.post_scale = layer_scale; a2.act = VRCOC_ACT_NONE;
  a2.res = x; a2.res_dtype = VRCOC_BF16;
  a2.out = out; a2.out_dtype = a2.out2_dtype = VRCOC_BF16;
  a2.out_sample_sums = out_sample_sums;
  MlpLayout L;
  L.nk1 = C / TC_BK;
  L.nh = hidden / TQ_MT;
  L.mt2 = (C + TQ_MT - 1) / TQ_MT;
  // one hidden buffer everywhere: at C > 128 the second one bought nothing (the second GEMM of chunk j is 0.8 us and overlaps the
  // accumulator read of chunk j+1) while its 32 KB as two more ring stages do — with a 4-deep ring 64 KB of weights were in
  // flight per SM against ~1 us of L2 latency and each chunk needs 176 KB: 3.6 us per chunk (trace), i.e. latency-bound
  L.h_bufs = 1;
  L.tmem_cols = L.mt2 == 1 ? 256 : 512;
  // ring depth: C <= 128: as deep as two resident CTAs per SM allow (227 KB, 1 KB reserved per CTA); else one CTA per SM
  const int budget = L.mt2 == 1 ? (227 * 1024) / 2 - 1024 : 220 * 1024;
  L.stages = 2;
  for (int stg = MF_MAX_STAGES; stg >= 2; --stg) {
    const int tot = L.nk1 * TQ_X_BYTES + stg * TQ_W_BYTES + L.h_bufs * 2 * TQ_X_BYTES + C * 16 + 512 + 1024;
    if (tot <= budget) { L.stages = stg; break; }
  }
  L.off_ring = L.nk1 * TQ_X_BYTES;
  L.off_h = L.off_ring + L.stages * TQ_W_BYTES;
  L.off_tab = L.off_h + L.h_bufs * 2 * TQ_X_BYTES;
  L.off_bar = L.off_tab + C * 16;
  L.total = L.off_bar + 512 + 1024;
  CUtensorMap tmX, tmW1, tmW2, tmO, tmR;
  int rc;
  {
    cuuint64_t dims[3] = {(cuuint64_t)P, (cuuint64_t)C, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)P * 2, (cuuint64_t)C * P * 2};
    cuuint32_t boxx[3] = {64, (cuuint32_t)TC_BK, 1}, boxo[3] = {64, 32, 1};
    if ((rc = tma_encode(&tmX, VRCOC_BF16, x, 3, dims, strides, boxx, true))) return rc;
    if ((rc = tma_encode(&tmR, VRCOC_BF16, x, 3, dims, strides, boxo, true))) return rc;
    if ((rc = tma_encode(&tmO, VRCOC_BF16, out, 3, dims, strides, boxo, true))) return rc;
  }
  {
    cuuint64_t d1[2] = {(cuuint64_t)C, (cuuint64_t)hidden}, s1[1] = {(cuuint64_t)C * 2};
    cuuint64_t d2[2] = {(cuuint64_t)hidden, (cuuint64_t)C}, s2[1] = {(cuuint64_t)hidden * 2};
    cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)TQ_MT};
    if ((rc = tma_encode(&tmW1, VRCOC_BF16, w1, 2, d1, s1, box, true))) return rc;
    if ((rc = tma_encode(&tmW2, VRCOC_BF16, w2, 2, d2, s2, box, true))) return rc;
  }
  dim3 grid((unsigned)cdiv(P, TQ_NP), 1, (unsigned)B);
  if (L.mt2 == 1) {
    set_smem(mlp_fused_kernel<true>, L.total);
    mlp_fused_kernel<true><<<grid, MF_THREADS, L.total, (cudaStream_t)stream>>>(a1, a2, L, b1, tmX, tmW1, tmW2, tmO, tmR);
  } else {
    set_smem(mlp_fused_kernel<false>, L.total);
    mlp_fused_kernel<false><<<grid, MF_THREADS_WIDE, L.total, (cudaStream_t)stream>>>(a1, a2, L, b1, tmX, tmW1, tmW2, tmO, tmR);
  }
  return check_launch("mlp_fused");
}

