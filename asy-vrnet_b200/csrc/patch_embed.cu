// Patch embedding of VRCoC (reference vr_coc.py:83-102 PointRecuder with patch 4 / stride 4 / pad 0, called from :575-587 on
// cat([x, pos])): a 4x4 stride-4 convolution of a FEW channels (3 or 4 map channels + the 2 position channels) at 512x512 into 64
// channels at 128x128.
//
// It is an HBM-bound gather (B*Cin*H*W*2 bytes in, B*64*P*2 bytes out: 42 MB at batch 8, ~7 us at the measured copy bandwidth) with a
// small contraction (K = 16*Cin <= 128) attached.  On the general tcgen05 engine it ran as 8192 CTAs of one 128-point tile each
// (60 us: CTA set-up, TMEM allocation and the scalar gather dominate); here one warp owns 16 consecutive output points, reads its
// 4 x 128-byte input rows per channel straight into mma.sync.m16n8k16 A fragments (non-overlapping patches: every input element
// is read exactly once, one k16 step = the 16 taps of one channel), the weights sit in shared memory in B-fragment order, and
// the CTA's 64 x 128 output tile leaves through shared memory as 256-byte rows.  The per-sample sum / sum of squares of the
// output (the GroupNorm statistics of the first ClusterBlock) leave with it, as in the engine's epilogue.
#include <cuda_bf16.h>

#include "common.cuh"

namespace vrcoc {

constexpr int PE_O = 64;            // output channels (embed_dims[0] of every coc_* factory that the detector uses)
constexpr int PE_MAX_CIN = 8;
constexpr int PE_PITCH = 136;       // staging row pitch in elements (272 B: 16-byte aligned, breaks the 256 B bank period)

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(256)
patch_embed4_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ extra, int64_t extra_bstride,
                    const __nv_bfloat16* __restrict__ w, const float* __restrict__ bias, __nv_bfloat16* __restrict__ out,
                    double* __restrict__ sums, int C0, int C1, int H, int W, int Wo, int Po) {
  __shared__ __align__(16) uint2 wf[PE_MAX_CIN * 8 * 32];                 // [c][n-tile j][lane] = {b0, b1}
  __shared__ __align__(16) __nv_bfloat16 tile[PE_O * PE_PITCH];           // [o][point]
  __shared__ float red[16];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int b = blockIdx.y, p0 = blockIdx.x * 128;
  const int Cin = C0 + C1;
  // weights [64][Cin][4][4] -> B fragments of m16n8k16 (k = the 16 taps of channel c, n = 8j + g)
  for (int i = tid; i < Cin * 8 * 32; i += 256) {
    const int l = i & 31, j = (i >> 5) & 7, c = i >> 8;
    const int n = 8 * j + (l >> 2), k0 = 2 * (l & 3);
    const uint32_t* wp = reinterpret_cast<const uint32_t*>(w + ((int64_t)n * Cin + c) * 16);
    wf[i] = make_uint2(__ldg(wp + (k0 >> 1)), __ldg(wp + ((k0 + 8) >> 1)));
  }
  // A fragments: point g (and g + 8) of the warp's 16, taps 2t, 2t+1 (row ky = t/2, columns 2*(t%2) ..) and 2t+8, 2t+9 (row 2 + t/2)
  const int p = p0 + 16 * warp;
  const int oy = p / Wo, ox0 = p - oy * Wo;
  uint32_t a[PE_MAX_CIN][4];
#pragma unroll
  for (int c = 0; c < PE_MAX_CIN; ++c) {
    if (c < Cin) {
      const __nv_bfloat16* plane = c < C0 ? x + ((int64_t)b * C0 + c) * H * W : extra + (int64_t)b * extra_bstride + (int64_t)(c - C0) * H * W;
      const __nv_bfloat16* r0 = plane + (int64_t)(4 * oy + (t >> 1)) * W + 4 * (ox0 + g) + 2 * (t & 1);
      a[c][0] = __ldg(reinterpret_cast<const uint32_t*>(r0));
      a[c][1] = __ldg(reinterpret_cast<const uint32_t*>(r0 + 32));
      a[c][2] = __ldg(reinterpret_cast<const uint32_t*>(r0 + 2 * (int64_t)W));
      a[c][3] = __ldg(reinterpret_cast<const uint32_t*>(r0 + 2 * (int64_t)W + 32));
    }
  }
  __syncthreads();                                                         // wf complete
  float acc[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float b0 = __ldg(bias + 8 * j + 2 * t), b1 = __ldg(bias + 8 * j + 2 * t + 1);
    acc[j][0] = b0; acc[j][1] = b1; acc[j][2] = b0; acc[j][3] = b1;
  }
#pragma unroll
  for (int c = 0; c < PE_MAX_CIN; ++c) {
    if (c < Cin) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint2 bf = wf[(c * 8 + j) * 32 + lane];
        mma_bf16_16816(acc[j], a[c], bf.x, bf.y);
      }
    }
  }
  // statistics on the fp32 values, then the tile through shared memory: c0,c1 = (point g, channels 8j+2t, +1), c2,c3 = point g+8
  float ssum = 0.f, ssq = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
#pragma unroll
    for (int e = 0; e < 4; ++e) { ssum += acc[j][e]; ssq = fmaf(acc[j][e], acc[j][e], ssq); }
    const int o = 8 * j + 2 * t, q = 16 * warp + g;
    tile[o * PE_PITCH + q] = __float2bfloat16_rn(acc[j][0]);
    tile[(o + 1) * PE_PITCH + q] = __float2bfloat16_rn(acc[j][1]);
    tile[o * PE_PITCH + q + 8] = __float2bfloat16_rn(acc[j][2]);
    tile[(o + 1) * PE_PITCH + q + 8] = __float2bfloat16_rn(acc[j][3]);
  }
  if (sums) {
    ssum = warp_sum(ssum);
    ssq = warp_sum(ssq);
    if (lane == 0) { red[warp] = ssum; red[8 + warp] = ssq; }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int o = (tid >> 4) + 16 * i, ch = tid & 15;
    *reinterpret_cast<uint4*>(out + ((int64_t)b * PE_O + o) * Po + p0 + 8 * ch) = *reinterpret_cast<const uint4*>(tile + o * PE_PITCH + 8 * ch);
  }
  if (sums && tid == 0) {
    double s = 0.0, s2 = 0.0;
    for (int i = 0; i < 8; ++i) { s += (double)red[i]; s2 += (double)red[8 + i]; }
    double* dst = sums + ((int64_t)b * VRCOC_STAT_SLOTS + (blockIdx.x & (VRCOC_STAT_SLOTS - 1))) * 2;
    atomicAdd(dst, s);
    atomicAdd(dst + 1, s2);
  }
}

}  // namespace vrcoc

extern "C" int vrcoc_patch_embed_supported(int dtype, int C0, int C1, int H, int W, int O, int patch) {
  if (dtype != VRCOC_BF16 || patch != 4 || O != vrcoc::PE_O || C0 <= 0 || C1 < 0 || C0 + C1 > vrcoc::PE_MAX_CIN) return 0;
  if (H % 4 != 0 || W % 64 != 0) return 0;                  // a warp's 16 points lie in one output row; 4-byte aligned fragment loads
  return ((int64_t)(H / 4) * (W / 4)) % 128 == 0 ? 1 : 0;   // whole 128-point tiles per sample
}

extern "C" int vrcoc_patch_embed(const void* x, const void* extra, int64_t extra_bstride, const void* weight, const float* bias, void* out,
                                 double* out_sample_sums, int dtype, int B, int C0, int C1, int H, int W, int O, int patch, void* stream) {
  using namespace vrcoc;
  VRCOC_REQUIRE(x && weight && bias && out && B > 0 && (C1 == 0 || extra), "patch_embed: bad argument");
  VRCOC_REQUIRE(vrcoc_patch_embed_supported(dtype, C0, C1, H, W, O, patch), "patch_embed: unsupported problem (bf16, patch 4, O = 64, "
                "C0 + C1 <= 8, W %% 64 == 0, (H/4)*(W/4) %% 128 == 0): C0=%d C1=%d H=%d W=%d O=%d patch=%d", C0, C1, H, W, O, patch);
  auto al = [](const void* q, uintptr_t m) { return (reinterpret_cast<uintptr_t>(q) & m) == 0; };
  VRCOC_REQUIRE(al(x, 3) && al(extra, 3) && al(weight, 3) && al(out, 15) && (extra_bstride % 2) == 0, "patch_embed: misaligned pointer");
  const int Wo = W / 4, Po = (H / 4) * Wo;
  dim3 grid((unsigned)(Po / 128), (unsigned)B);
  patch_embed4_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)extra, extra_bstride,
                                                             (const __nv_bfloat16*)weight, bias, (__nv_bfloat16*)out, out_sample_sums, C0, C1,
                                                             H, W, Wo, Po);
  return check_launch("patch_embed");
}
