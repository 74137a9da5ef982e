// Convolution-as-GEMM engine, tcgen05 path (5th-gen tensor cores, bf16 operands, fp32 accumulation in TMEM).
// Same operator contract as the CUDA-core path (include/vrcoc.h, conv_common.cuh); used whenever the weights are bf16.
//
//   D[128 points x N outs] (TMEM, fp32)  +=  A[128 points x 64 k] (smem, bf16)  *  B[N outs x 64 k]^T (smem, bf16)
//
// * M = points.  In NCHW the points of one channel are contiguous, so A is an MN-major operand: each K-row of the
//   smem tile holds 64 consecutive points (128 B) of one logical input channel, two 64-point blocks per tile,
//   128-byte swizzled — the canonical UMMA layout  Sw<3,4,3> o ((8,2),(8,k)) : ((1,LBO),(8,SBO))  in 16-byte units
//   with LBO = 8192 B (next 64-point block) and SBO = 1024 B (next group of 8 k-rows).
// * A is produced by the CTA's own threads ("transform on load"): 128-bit coalesced global loads, the prologue
//   (GroupNorm apply / attention gate / ECA scale / channel shuffle / im2col gather) applied in registers in fp32, one
//   rounding to bf16, one 16-byte swizzled st.shared.  The normalised / gated / gathered activation never exists in HBM,
//   and no weight folding (with its cancellation problem) is needed.
// * B = weights [O][K] row-major = K-major operand: 8-row x 128-byte swizzle atoms, SBO = 1024 B.
// * One elected thread issues tcgen05.mma (M=128, N=N_tile, K=16) per 16 k; completion is tracked with tcgen05.commit on
//   mbarriers: one per smem stage (frees the stage for the next slab) and one for the finished accumulator.
// * Epilogue: 8 warps read the accumulator with tcgen05.ld (32 lanes x 16 columns per instruction; warp w owns TMEM lanes
//   32*(w%4).. and column half w/4), apply bias/BN/activation/layer-scale/residual/BN and the side statistics in
//   registers and store NCHW directly: lane = point, so every output channel is a coalesced 64/128-byte store per warp.
#include "conv_common.cuh"

namespace vrcoc {

constexpr int TC_THREADS = 256;
constexpr int TC_BM = 128;            // points per CTA
constexpr int TC_BK = 64;             // k per smem slab (128 B of bf16 per B row)
constexpr int TC_A_BYTES = TC_BM * TC_BK * 2;       // 16 KB
constexpr int TC_A_LBO = 64 * TC_BK * 2;            // 8192 B: second 64-point block
constexpr int TC_STAGES = 2;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

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

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ uint4 pack8_bf16(const float (&v)[8]) {
  uint4 r;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  return r;
}

struct TcLayout {
  int n_tile;       // output channels per CTA (multiple of 32, <= 256)
  int tmem_cols;    // power of two >= 32
  int b_bytes;      // n_tile * 128
  int off_b, off_tab, off_bar, total;
};

__global__ void __launch_bounds__(TC_THREADS) conv_tc_kernel(ConvArgs a, TcLayout L) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // 1024-byte alignment is required by the 128-byte swizzle atoms
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* sA = smem;                                   // [STAGES][16 KB]
  unsigned char* sB = smem + L.off_b;                         // [STAGES][b_bytes]
  float4* tab = reinterpret_cast<float4*>(smem + L.off_tab);  // [Cin]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.off_bar);   // [STAGES] stage-free + [1] accumulator-done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + TC_STAGES + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.z;
  const int p0 = blockIdx.x * TC_BM;
  const int n0 = blockIdx.y * L.n_tile;
  const int P = a.P_out;
  const int taps = a.kh * a.kw;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)L.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 32) {
    for (int i = 0; i < TC_STAGES + 1; ++i) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  build_prologue_table(a, b, tab);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int nk = (a.K + TC_BK - 1) / TC_BK;
  const uint32_t idesc = make_idesc(L.n_tile);

  // A loader geometry: thread -> (8-point chunk, k-rows krow0 + 16*i)
  const int a_chunk = tid & 15;            // points a_chunk*8 .. +8 of the tile
  const int a_krow0 = tid >> 4;            // 0..15
  const int a_blk = a_chunk >> 3;          // 64-point block
  const int a_c = a_chunk & 7;             // 16-byte chunk inside the 128-byte row
  const int q0 = p0 + a_chunk * 8;

  for (int kc = 0; kc < nk; ++kc) {
    const int s = kc % TC_STAGES;
    if (kc >= TC_STAGES) mbar_wait(&bars[s], (uint32_t)(((kc / TC_STAGES) - 1) & 1));
    const int kvalid = min(TC_BK, a.K - kc * TC_BK);
    const int ksteps = (kvalid + 15) >> 4;            // MMAs (K=16) for this slab
    unsigned char* As = sA + s * TC_A_BYTES;
    unsigned char* Bs = sB + s * L.b_bytes;

    // ---- A slab: transform on load ---------------------------------------------------------------------------
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k = a_krow0 + 16 * i;
      if (k >= ksteps * 16) break;
      const int kk = kc * TC_BK + k;
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = 0.f;
      if (kk < a.K && q0 < P) {
        const int c = kk / taps;
        const int tap = kk - c * taps;
        const int sc = a.chan_src ? a.chan_src[c] : c;
        const void* src; int dt; int64_t base;
        if (sc < a.C0) { src = a.src0; dt = a.src0_dtype; base = (int64_t)b * a.src0_bstride + (int64_t)sc * a.P_in; }
        else           { src = a.src1; dt = a.src1_dtype; base = (int64_t)b * a.src1_bstride + (int64_t)(sc - a.C0) * a.P_in; }
        const float4 t = tab[c];
        if (a.fast1x1) {
          if (dt == VRCOC_F32) ld8<float>(reinterpret_cast<const float*>(src) + base + q0, v);
          else ld8<__nv_bfloat16>(reinterpret_cast<const __nv_bfloat16*>(src) + base + q0, v);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float x = v[j];
            float y = fmaf(x, t.x, t.y);
            if (a.has_gate) y *= sigmoidf_exact(fmaf(t.z, x, t.w));
            v[j] = y;
          }
        } else {
          const int ky = tap / a.kw, kx = tap - ky * a.kw;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int q = q0 + j;
            if (q < P) {
              const int oy = q / a.W_out, ox = q - oy * a.W_out;
              const int iy = oy * a.stride - a.pad + ky, ix = ox * a.stride - a.pad + kx;
              if (iy >= 0 && iy < a.H_in && ix >= 0 && ix < a.W_in) {
                float x = ld_any(src, base + (int64_t)iy * a.W_in + ix, dt);
                float y = fmaf(x, t.x, t.y);
                if (a.has_gate) y *= sigmoidf_exact(fmaf(t.z, x, t.w));
                v[j] = y;
              }
            }
          }
        }
      }
      *reinterpret_cast<uint4*>(As + a_blk * TC_A_LBO + k * 128 + ((a_c ^ (k & 7)) << 4)) = pack8_bf16(v);
    }

    // ---- B slab: weights, K-major ---------------------------------------------------------------------------------
    {
      const int units = L.n_tile * (ksteps * 2);      // 16-byte chunks: n_tile rows x (ksteps*16/8) chunks
      const int cpr = ksteps * 2;
      const bool wvec = (a.K % 8 == 0) && ((reinterpret_cast<uintptr_t>(a.weight) & 15) == 0);
      const __nv_bfloat16* W = reinterpret_cast<const __nv_bfloat16*>(a.weight);
      for (int u = tid; u < units; u += TC_THREADS) {
        const int n = u / cpr, c = u - n * cpr;
        const int o = n0 + n;
        const int kk = kc * TC_BK + c * 8;
        uint4 w = make_uint4(0u, 0u, 0u, 0u);
        if (o < a.O && kk < a.K) {
          const __nv_bfloat16* wp = W + (int64_t)o * a.K + kk;
          if (wvec) {
            w = __ldg(reinterpret_cast<const uint4*>(wp));
          } else {
            __nv_bfloat16 tmp[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) tmp[j] = (kk + j < a.K) ? wp[j] : __float2bfloat16_rn(0.f);
            w = *reinterpret_cast<uint4*>(tmp);
          }
        }
        *reinterpret_cast<uint4*>(Bs + (n >> 3) * 1024 + (n & 7) * 128 + ((c ^ (n & 7)) << 4)) = w;
      }
    }

    fence_async_smem();          // generic-proxy smem writes -> visible to the tensor core (async proxy)
    __syncthreads();

    if (tid == 0) {
      tc_fence_after();
      const uint32_t a_addr = smem_u32(As), b_addr = smem_u32(Bs);
      for (int j = 0; j < ksteps; ++j) {
        const uint64_t ad = make_desc(a_addr + j * 2048, TC_A_LBO, 1024);   // 16 k-rows = two 8-row groups
        const uint64_t bd = make_desc(b_addr + j * 32, 16, 1024);           // 16 k = 32 B inside the 128 B row
        tc_mma(tmem_base, ad, bd, idesc, (kc > 0 || j > 0) ? 1u : 0u);
      }
      tc_commit(&bars[s]);                       // stage s reusable once these MMAs have read it
      if (kc == nk - 1) tc_commit(&bars[TC_STAGES]);
    }
  }

  // ---- epilogue ---------------------------------------------------------------------------------------------------------
  mbar_wait(&bars[TC_STAGES], 0);
  tc_fence_after();
  const int lq = warp & 3, chalf = warp >> 2;
  const int m = lq * 32 + lane;
  const int q = p0 + m;
  const bool valid = q < P;
  const int ncols = L.n_tile >> 1;
  float ssum = 0.f, ssq = 0.f, vmax = 0.f, vmin = __int_as_float(0x7f800000);
  for (int c0 = chalf * ncols; c0 < (chalf + 1) * ncols; c0 += 16) {
    if (n0 + c0 >= a.O) break;                   // warp-uniform
    float acc[16];
    tmem_ld16(tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)c0, acc);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int o = n0 + c0 + j;
      if (o < a.O && valid) {
        const EpiCoef ec = load_epi(a, o);
        const float r = a.res ? ld_any(a.res, ((int64_t)b * a.O + o) * P + q, a.res_dtype) : 0.f;
        const float y = epilogue_value(acc[j], ec, a.act, r);
        ssum += y; ssq = fmaf(y, y, ssq);
        vmax = fmaxf(vmax, y); vmin = fminf(vmin, y);
        if (o < a.O_split) st_any(a.out, ((int64_t)b * a.O_split + o) * P + q, a.out_dtype, y);
        else st_any(a.out2, ((int64_t)b * (a.O - a.O_split) + (o - a.O_split)) * P + q, a.out2_dtype, y);
      }
    }
  }
  emit_side_stats(a, b, ssum, ssq, vmax, vmin);

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)L.tmem_cols));
  }
}

static TcLayout tc_layout(const ConvArgs& a) {
  TcLayout L{};
  int ntiles = (int)cdiv(a.O, 256);
  int per = (int)cdiv(a.O, ntiles);
  L.n_tile = (int)cdiv(per, 32) * 32;
  if (L.n_tile > 256) L.n_tile = 256;
  L.tmem_cols = 32;
  while (L.tmem_cols < L.n_tile) L.tmem_cols *= 2;
  L.b_bytes = L.n_tile * 128;
  L.off_b = TC_STAGES * TC_A_BYTES;
  L.off_tab = L.off_b + TC_STAGES * L.b_bytes;
  L.off_bar = L.off_tab + a.Cin * 16;
  L.off_bar = (L.off_bar + 15) & ~15;
  L.total = L.off_bar + (TC_STAGES + 1) * 8 + 16 + 1024;   // + alignment slack
  return L;
}

bool conv_tc_supported(const ConvArgs& a) {
  if (a.weight_dtype != VRCOC_BF16) return false;
  TcLayout L = tc_layout(a);
  return L.total <= 220 * 1024;
}

int launch_conv_tc(const ConvArgs& a, cudaStream_t st) {
  TcLayout L = tc_layout(a);
  VRCOC_REQUIRE(L.total <= 220 * 1024, "conv(tcgen05): shared memory budget exceeded (%d bytes)", L.total);
  cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total);
  dim3 grid((unsigned)cdiv(a.P_out, TC_BM), (unsigned)cdiv(a.O, L.n_tile), (unsigned)a.B);
  conv_tc_kernel<<<grid, TC_THREADS, L.total, st>>>(a, L);
  return check_launch("conv_tc");
}

}  // namespace vrcoc
