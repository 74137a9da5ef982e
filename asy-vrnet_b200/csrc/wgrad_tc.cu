// Weight gradient of a 1x1 projection on the tensor cores (training path; reference: autograd of nn.Conv2d in
// backbone/fusion/vr_coc.py:139-141, :205-207).
//
//   dW[o][c] = sum_b sum_p dY[b][o][p] * xhat[b][c][p],   xhat = prologue(x) = s[b][c] * x + t[b][c]   (GroupNorm(1,C) / table)
//            = sum_b ( s[b][c] * G_b[o][c] + t[b][c] * r_b[o] ),   G_b = dY_b . x_b^T,   r_b[o] = sum_p dY[b][o][p]
//
// so the contraction runs on the RAW bf16 activations: in NCHW the points are contiguous in both dY and x, i.e. both are
// K-major operands that TMA drops straight into the SW128 UMMA layout - no transform, no prologue in the main loop.  r_b comes
// out of the same pass as a second, 16-column MMA against a constant tile of ones.  One CTA = 128 outputs x n_tile channels of
// one (sample, point range); the per-(sample, range) partials go to the caller's workspace and a second kernel applies the
// prologue coefficients and reduces them in a fixed order (deterministic, like the CUDA-core path it replaces - which the
// profile of a ClusterBlock fwd+bwd showed taking 61-75 % of the time).
#include <stdlib.h>

#include "conv_common.cuh"
#include "tma.cuh"

namespace vrcoc {
namespace {

constexpr int WT_THREADS = 192;                 // warp 0: TMA producer (+ TMEM allocation), warp 1: MMA issuer, warps 2-5: epilogue
constexpr int WT_BM = 128;                      // outputs per CTA
constexpr int WT_BK = 64;                       // points per slab (128 B of bf16 per operand row)
constexpr int WT_A_BYTES = WT_BM * WT_BK * 2;   // 16 KB
constexpr int WT_ONES_BYTES = 16 * 128;         // 16 rows of ones
constexpr int WT_MAX_STAGES = 6;

struct WtArgs {
  float* ws_g;                                  // [parts][O][C]
  float* ws_r;                                  // [parts][O]
  int B, O, C, P;
  int splits;                                   // point ranges per sample
  int slabs_per_split, slabs;
  int n_tile, tmem_cols, stages;
};

__device__ __forceinline__ void wt_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void wt_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void wt_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// K-major SW128 operand descriptor: start>>4 | LBO>>4 << 16 | SBO>>4 << 32 | version 1 << 46 | SWIZZLE_128B << 61
__device__ __forceinline__ uint64_t wt_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16: D = f32, A = B = bf16, both K-major, M = 128
__host__ __device__ constexpr uint32_t wt_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(WT_BM >> 4) << 24);
}
__device__ __forceinline__ void wt_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void wt_tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(WT_THREADS) wgrad_tc_kernel(WtArgs a, const __grid_constant__ CUtensorMap tmDY,
                                                           const __grid_constant__ CUtensorMap tmX) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int b_bytes = a.n_tile * 128;
  unsigned char* sA = smem;                                          // [stages][16 KB]   dY tile: 128 outputs x 64 points
  unsigned char* sB = sA + a.stages * WT_A_BYTES;                    // [stages][n_tile x 128 B]   x tile: channels x 64 points
  unsigned char* sOnes = sB + a.stages * b_bytes;                    // 16 x 128 B of bf16 1.0 (any swizzle of ones is ones)
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(sOnes + WT_ONES_BYTES);
  uint64_t* bar_free = bar_full + WT_MAX_STAGES;
  uint64_t* bar_acc = bar_free + WT_MAX_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_acc + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int o0 = blockIdx.x * WT_BM, c0 = blockIdx.y * a.n_tile;
  const int part = blockIdx.z, b = part / a.splits, split = part - b * a.splits;
  const int s_begin = split * a.slabs_per_split;
  int s_end = s_begin + a.slabs_per_split;
  if (s_end > a.slabs) s_end = a.slabs;
  const int nk = s_end - s_begin;                                    // >= 1 by construction of the grid
  const bool with_ones = blockIdx.y == 0;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)a.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 32) {
    for (int i = 0; i < a.stages; ++i) { mbar_init(&bar_full[i], 1); mbar_init(&bar_free[i], 1); }
    mbar_init(bar_acc, 1);
    mbar_fence_init();
    tma_prefetch_desc(&tmDY);
    tma_prefetch_desc(&tmX);
  }
  if (warp >= 2) {
    for (int i = tid - 64; i < WT_ONES_BYTES / 4; i += WT_THREADS - 64) reinterpret_cast<uint32_t*>(sOnes)[i] = 0x3F803F80u;
    fence_async_smem();
  }
  wt_fence_before();
  __syncthreads();
  wt_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < nk; ++it) {
        const int s = it % a.stages;
        if (it >= a.stages) mbar_wait(&bar_free[s], (uint32_t)((it / a.stages) - 1) & 1);
        mbar_expect_tx(&bar_full[s], (uint32_t)(WT_A_BYTES + b_bytes));
        const int p = (s_begin + it) * WT_BK;
        tma_load_3d(sA + s * WT_A_BYTES, &tmDY, p, o0, b, &bar_full[s]);
        tma_load_3d(sB + s * b_bytes, &tmX, p, c0, b, &bar_full[s]);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = wt_idesc(a.n_tile), idesc1 = wt_idesc(16);
      const uint64_t ones = wt_desc(smem_u32(sOnes));
      for (int it = 0; it < nk; ++it) {
        const int s = it % a.stages;
        mbar_wait(&bar_full[s], (uint32_t)(it / a.stages) & 1);
        wt_fence_after();
        const uint32_t a_addr = smem_u32(sA + s * WT_A_BYTES), b_addr = smem_u32(sB + s * b_bytes);
#pragma unroll
        for (int j = 0; j < WT_BK / 16; ++j) {                         // 16 points = 32 B inside the 128-byte rows
          const uint64_t ad = wt_desc(a_addr + j * 32);
          const uint32_t acc = (it > 0 || j > 0) ? 1u : 0u;
          wt_mma(tmem_base, ad, wt_desc(b_addr + j * 32), idesc, acc);
          if (with_ones) wt_mma(tmem_base + (uint32_t)a.n_tile, ad, ones, idesc1, acc);
        }
        wt_commit(&bar_free[s]);
      }
      wt_commit(bar_acc);
    }
    __syncwarp();
  } else {
    // ---- epilogue: TMEM lane = output row; the raw partial goes to the workspace as it is ---------------------------------------
    const int lq = warp & 3;                                         // warps 2,3,4,5 -> lane quarters 2,3,0,1
    const int o = o0 + lq * 32 + lane;
    mbar_wait(bar_acc, 0);
    wt_fence_after();
    const uint32_t tbase = tmem_base + ((uint32_t)(lq * 32) << 16);
    float* grow = a.ws_g + ((int64_t)part * a.O + o) * a.C;
    const bool vec = (a.C & 3) == 0;
    for (int cc = 0; cc < a.n_tile; cc += 16) {
      if (c0 + cc >= a.C) break;                                     // warp-uniform
      uint32_t r[16];
      wt_tmem_ld16(tbase + (uint32_t)cc, r);
      if (o < a.O) {
        if (vec && c0 + cc + 16 <= a.C) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            *reinterpret_cast<uint4*>(grow + c0 + cc + 4 * i) = make_uint4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (c0 + cc + i < a.C) grow[c0 + cc + i] = __uint_as_float(r[i]);
        }
      }
    }
    if (with_ones) {
      uint32_t r[16];
      wt_tmem_ld16(tbase + (uint32_t)a.n_tile, r);
      if (o < a.O) a.ws_r[(int64_t)part * a.O + o] = __uint_as_float(r[0]);
    }
  }
  wt_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)a.tmem_cols));
}

// dW[o][c] = sum_b ( s_bc * sum_splits G + t_bc * sum_splits r ),  db[o] = sum r      (fixed order: deterministic)
__global__ void __launch_bounds__(256) wgrad_tc_combine_kernel(ConvArgs a, WtArgs w, float* __restrict__ dW, float* __restrict__ db) {
  __shared__ float stat[64][2];                                      // per-sample mean / rstd (GroupNorm prologue)
  if (a.gn_sums) {
    for (int b = threadIdx.x >> 5; b < a.B && b < 64; b += blockDim.x >> 5) {
      float mu, rstd;
      gn_mean_rstd(a.gn_sums, b, (double)a.C0 * (double)a.P_in, a.gn_eps, mu, rstd);
      if ((threadIdx.x & 31) == 0) { stat[b][0] = mu; stat[b][1] = rstd; }
    }
  }
  __syncthreads();
  const int64_t n = (int64_t)a.O * a.Cin;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int o = (int)(i / a.Cin), c = (int)(i - (int64_t)o * a.Cin);
  const float gamma = a.gn_sums ? a.gn_gamma[c] : 1.f, beta = a.gn_sums ? a.gn_beta[c] : 0.f;
  float acc = 0.f, racc = 0.f;
  for (int b = 0; b < a.B; ++b) {
    float g = 0.f, r = 0.f;
    for (int s = 0; s < w.splits; ++s) {
      const int64_t part = (int64_t)b * w.splits + s;
      g += w.ws_g[(part * a.O + o) * a.Cin + c];
      r += w.ws_r[part * a.O + o];
    }
    float sc = 1.f, sh = 0.f;
    if (a.gn_sums) {
      sc = stat[b][1] * gamma;
      sh = fmaf(-stat[b][0], sc, beta);
    } else if (a.table) {
      const float4 t = reinterpret_cast<const float4*>(a.table)[(int64_t)b * a.Cin + c];
      sc = t.x; sh = t.y;
    }
    acc = fmaf(sc, g, fmaf(sh, r, acc));
    racc += r;
  }
  dW[i] = acc;
  if (db && c == 0) db[o] = racc;
}

bool plan(const ConvArgs& a, int dy_dtype, WtArgs& w, int& smem_bytes) {
  const char* knob = getenv("VRCOC_WGRAD_TC");                       // "0": A/B switch (tests, tools)
  if (knob && knob[0] == '0') return false;
  if (dy_dtype != VRCOC_BF16 || a.src0_dtype != VRCOC_BF16 || a.C1 != 0 || a.chan_src || a.has_gate) return false;
  if (a.kh != 1 || a.kw != 1 || a.stride != 1 || a.pad != 0 || a.P_in != a.P_out) return false;
  if ((a.P_out % 8) != 0 || (a.src0_bstride % 8) != 0 || (reinterpret_cast<uintptr_t>(a.src0) & 15) != 0 || a.B > 64) return false;
  if (tma_encode_fn() == nullptr) return false;
  w.B = a.B; w.O = a.O; w.C = a.Cin; w.P = a.P_out;
  const int c_tiles = (int)cdiv(a.Cin, 256);
  w.n_tile = (int)cdiv(cdiv(a.Cin, c_tiles), 16) * 16;
  w.tmem_cols = 32;
  while (w.tmem_cols < w.n_tile + 16) w.tmem_cols *= 2;
  w.slabs = (int)cdiv(a.P_out, WT_BK);
  // point ranges per sample: enough CTAs for one wave, but never ranges shorter than 8 slabs (the set-up has to amortise)
  const int64_t base = cdiv(a.O, WT_BM) * cdiv(a.Cin, w.n_tile) * a.B;
  int splits = (int)cdiv(sm_count(), base);
  const int max_splits = w.slabs / 8 > 0 ? w.slabs / 8 : 1;
  if (splits > max_splits) splits = max_splits;
  w.slabs_per_split = (int)cdiv(w.slabs, splits);
  w.splits = (int)cdiv(w.slabs, w.slabs_per_split);
  const int stage_bytes = WT_A_BYTES + w.n_tile * 128;
  int stages = (200 * 1024) / stage_bytes;
  if (stages > WT_MAX_STAGES) stages = WT_MAX_STAGES;
  if (stages > w.slabs_per_split) stages = w.slabs_per_split;
  if (stages < 1) stages = 1;
  w.stages = stages;
  smem_bytes = stages * stage_bytes + WT_ONES_BYTES + (2 * WT_MAX_STAGES + 1) * 8 + 16 + 1024;
  return true;
}

}  // namespace

// floats of workspace the tensor-core path needs, or -1 when it does not cover the problem
int64_t wgrad_tc_workspace(const ConvArgs& a, int dy_dtype) {
  WtArgs w{};
  int smem;
  if (!plan(a, dy_dtype, w, smem)) return -1;
  return (int64_t)a.B * w.splits * ((int64_t)a.O * a.Cin + a.O);
}

// VRCOC_OK when launched, 1 when not covered
int launch_wgrad_tc(const ConvArgs& a, const void* dy, int dy_dtype, float* dW, float* db, float* ws, cudaStream_t st) {
  WtArgs w{};
  int smem;
  if (!plan(a, dy_dtype, w, smem) || (reinterpret_cast<uintptr_t>(dy) & 15) != 0) return 1;
  const int64_t parts = (int64_t)a.B * w.splits;
  w.ws_g = ws;
  w.ws_r = ws + parts * a.O * a.Cin;
  CUtensorMap tmDY, tmX;
  {
    cuuint64_t dims[3] = {(cuuint64_t)a.P_out, (cuuint64_t)a.O, (cuuint64_t)a.B};
    cuuint64_t strides[2] = {(cuuint64_t)a.P_out * 2, (cuuint64_t)a.O * a.P_out * 2};
    cuuint32_t box[3] = {(cuuint32_t)WT_BK, (cuuint32_t)WT_BM, 1};
    int rc = tma_encode(&tmDY, VRCOC_BF16, dy, 3, dims, strides, box, true);
    if (rc) return rc;
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)a.P_in, (cuuint64_t)a.Cin, (cuuint64_t)a.B};
    cuuint64_t strides[2] = {(cuuint64_t)a.P_in * 2, (cuuint64_t)a.src0_bstride * 2};
    cuuint32_t box[3] = {(cuuint32_t)WT_BK, (cuuint32_t)w.n_tile, 1};
    int rc = tma_encode(&tmX, VRCOC_BF16, a.src0, 3, dims, strides, box, true);
    if (rc) return rc;
  }
  cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  dim3 grid((unsigned)cdiv(a.O, WT_BM), (unsigned)cdiv(a.Cin, w.n_tile), (unsigned)parts);
  wgrad_tc_kernel<<<grid, WT_THREADS, smem, st>>>(w, tmDY, tmX);
  int rc = check_launch("wgrad_tc");
  if (rc) return rc;
  const int64_t n = (int64_t)a.O * a.Cin;
  wgrad_tc_combine_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(a, w, dW, db);
  return check_launch("wgrad_tc.combine");
}

}  // namespace vrcoc
