// Stand-alone probe: tcgen05.mma issue rate for the operand layouts the projection kernels use.
//   D[128 x N] (TMEM) += A[128 x 16] * B[N x 16]^T, bf16, fp32 accumulate; operands stay in shared memory (no loads in the
//   timed loop), one thread issues `iters` x 4 MMAs (one 64-deep slab) and commits once per slab, as the kernels do.
//   variants: A K-major / B K-major (wgrad_tc), A K-major / B MN-major (channel-major kernel), A MN-major / B K-major
//   (point-major kernels); N = 64, 128, 256.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I asy-vrnet_b200/csrc tools/mma_rate_probe.cu -o tools/mma_rate_probe -lcuda
#include "tma.cuh"
namespace vrcoc { char* err_buf() { static char b[256]; return b; } int fail(int c, const char* f, ...) { printf("fail: %s\n", f); return c; } int check_launch(const char*) { return 0; } }
using namespace vrcoc;

__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b),
               "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// amaj / bmaj: 0 = K-major, 1 = MN-major.  mode: what the other warps do meanwhile (the projection kernels' epilogue / producer):
//   0 nothing; 1 warps 1-8 read the OTHER accumulator buffer (TMEM columns 128..255) with tcgen05.ld.x16 in a loop;
//   2 warps 1-8 stream 16-byte shared-memory stores + loads over a private 32 KB region; 3 warp 9 keeps 16 KB TMA boxes (L2
//   hits) landing in a 4-slot ring; 4 = 1 + 2 + 3; 5: tcgen05.fence::after_thread_sync before every slab (as the kernels do
//   after their full-barrier wait); 6: that fence + a wait on an already completed mbarrier phase; 7: the wait alone
__global__ void __launch_bounds__(320) mma_kernel(int n, int amaj, int bmaj, int iters, long long* cycles, int mode,
                                                 const __grid_constant__ CUtensorMap tm) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint64_t ring[4];
  __shared__ uint32_t slot;
  __shared__ volatile int done;
  unsigned char* sA = smem;                 // 16 KB: 128 rows x 64 k (either layout)
  unsigned char* sB = smem + 16384;         // up to 32 KB: 256 rows x 64 k
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(256u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (threadIdx.x == 0) { mbar_init(&bar, (uint32_t)iters + 1u); for (int i = 0; i < 4; ++i) mbar_init(&ring[i], 1); done = 0; mbar_fence_init(); }
  fence_async_smem();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)amaj << 15) | ((uint32_t)bmaj << 16) | ((uint32_t)(n >> 3) << 17) | (8u << 24);
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (mode == 6 || mode == 7) mbar_wait(&ring[0], 1);               // an already-completed phase (never armed: parity 1 passes)
      if (mode == 5 || mode == 6) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // modes 8, 9: operands walk over a large footprint like the kernels' (X resident 80 KB + a 64 KB weight ring)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint64_t ad = amaj ? desc(a0 + j * 2048, 8192, 1024) : desc(a0 + j * 32, 16, 1024);
        const uint64_t bd = bmaj ? desc(b0 + j * 2048, 8192, 1024) : desc(b0 + j * 32, 16, 1024);
        mma(tmem, ad, bd, idesc, (it | j) ? 1u : 0u);
      }
      commit(&bar);                                      // one arrival per slab; the barrier expects iters + 1 of them
    }
    commit(&bar);
    mbar_wait(&bar, 0);
    cycles[blockIdx.x] = clock64() - t0;
    done = 1;
  }
  const int warp = threadIdx.x >> 5;
  if (warp >= 1 && warp <= 8) {
    uint32_t sink = 0;
    unsigned char* priv = smem + 49152 + (warp - 1) * 4096;             // 8 x 4 KB private staging regions
    while (!done) {
      if (mode == 1 || mode == 4) {
        uint32_t r[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                       "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(tmem + ((uint32_t)((warp & 3) * 32) << 16) + 128u + (uint32_t)(((warp - 1) >> 2) * 64)));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        sink += r[0] + r[15];
      }
      if (mode == 2 || mode == 4) {
        uint4 v;
        const uint32_t pa = smem_u32(priv) + (threadIdx.x & 31) * 128;
        asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(pa + ((sink & 7) << 4)));
        v.x += 1;
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(pa + (((sink + 3) & 7) << 4)), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
        sink += v.y;
      }
      if (mode == 0 || mode == 3) __nanosleep(200);
    }
    if (sink == 0x12345678u) cycles[0] = 0;
  }
  if (warp == 9 && (threadIdx.x & 31) == 0 && (mode == 3 || mode == 4)) {
    unsigned char* dst = smem + 49152 + 32768;                              // 4 x 16 KB ring
    for (int it = 0; !done; ++it) {
      const int s2 = it & 3;
      if (it >= 4) mbar_wait(&ring[s2], (uint32_t)((it >> 2) - 1) & 1);
      mbar_expect_tx(&ring[s2], 16384u);
      tma_load_2d(dst + s2 * 16384, &tm, (it & 31) * 64, ((it >> 5) & 63) * 128, &ring[s2]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u));
}

int main() {
  long long* d;
  cudaMalloc(&d, 148 * 8);
    uint16_t* w;
  cudaMalloc(&w, (size_t)2048 * 8192 * 2);
  cudaMemset(w, 0, (size_t)2048 * 8192 * 2);
  CUtensorMap tm;
  {
    cuuint64_t dims[2] = {2048, 8192};
    cuuint64_t strides[1] = {2048 * 2};
    cuuint32_t box[2] = {64, 128};
    if (tma_encode(&tm, VRCOC_BF16, w, 2, dims, strides, box, true)) return 1;
  }
  cudaFuncSetAttribute(mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  printf("%4s %6s %6s %5s | %12s %14s %10s\n", "N", "A", "B", "mode", "clk / MMA", "flop/clk/SM", "of 8192");
  for (int n : {128})
    for (int mode : {0, 5, 6, 7})
    for (int v = 0; v < 2; ++v) {
      const int amaj = 0, bmaj = v == 1, iters = 4000;
      mma_kernel<<<148, 320, 150 * 1024>>>(n, amaj, bmaj, 64, d, mode, tm);
      mma_kernel<<<148, 320, 150 * 1024>>>(n, amaj, bmaj, iters, d, mode, tm);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
      long long h[148];
      cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      double c = 0;
      for (int i = 0; i < 148; ++i) c += (double)h[i];
      c /= 148.0 * iters * 4;
      const double fl = 2.0 * 128 * n * 16 / c;
      printf("%4d %6s %6s %5d | %12.1f %14.0f %9.1f%%\n", n, amaj ? "MN" : "K", bmaj ? "MN" : "K", mode, c, fl, 100.0 * fl / 8192.0);
    }
  return 0;
}
