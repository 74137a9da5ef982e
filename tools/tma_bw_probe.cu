// Stand-alone probe: how fast can the SMs RECEIVE TMA boxes from L2-resident data?  (the deep-stage GEMMs stream ~16 KB weight
// slabs at ~30 GB/s per SM whatever the ring depth - DESIGN.md 4)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I asy-vrnet_b200/csrc tools/tma_bw_probe.cu -o tools/tma_bw_probe -lcuda
//   ./tools/tma_bw_probe            -> table: footprint x ring depth x box rows
#include <vector>
#include "tma.cuh"
namespace vrcoc { char* err_buf() { static char b[256]; return b; } int fail(int c, const char* f, ...) { printf("fail: %s\n", f); return c; } int check_launch(const char*) { return 0; } }
using namespace vrcoc;

// one CTA per SM; thread 0 keeps `stages` boxes of [rows x 64 bf16] (rows * 128 B) in flight, walking `nboxes` distinct boxes of
// the tensor starting at a per-CTA offset (spread = 1) or all CTAs the same sequence (spread = 0); nobody reads the data
__global__ void __launch_bounds__(128) bw_kernel(const __grid_constant__ CUtensorMap tm, int stages, int rows, int iters, int nboxes, int spread,
                                                 int kslabs, int issuers, int by_lane) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bars[4][16];
  const int bytes = rows * 128;
  const int w = by_lane ? (threadIdx.x < 4 ? threadIdx.x : 99) : ((threadIdx.x & 31) == 0 ? (threadIdx.x >> 5) : 99);   // issuer id
  if (threadIdx.x == 0) {
    for (int q = 0; q < 4; ++q)
      for (int i = 0; i < stages; ++i) mbar_init(&bars[q][i], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (w < issuers) {
    uint64_t* bar = bars[w];
    smem += w * stages * bytes;
    int box = spread ? (int)(((blockIdx.x * 4 + w) * 7919u) % (unsigned)nboxes) : 0;
    for (int it = 0; it < iters + stages; ++it) {
      const int s = it % stages;
      if (it >= stages) mbar_wait(&bar[s], (uint32_t)((it / stages) - 1) & 1);
      if (it < iters) {
        mbar_expect_tx(&bar[s], (uint32_t)bytes);
        tma_load_2d(smem + s * bytes, &tm, (box % kslabs) * 64, (box / kslabs) * rows, &bar[s]);
        box = (box + 1) % nboxes;
      }
    }
  }
}

int main() {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int K = 2048;                                   // bf16 columns (32 slabs of 64)
  const int R = 16384;                                  // rows: 64 MB tensor; sub-ranges give the footprints
  uint16_t* d;
  cudaMalloc(&d, (size_t)K * R * 2);
  cudaMemset(d, 0, (size_t)K * R * 2);
  cudaFuncSetAttribute(bw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  printf("%8s %6s %6s %7s | %10s %12s\n", "foot MB", "rows", "depth", "issuers", "GB/s/SM", "TB/s total");
  for (int rows : {128, 64}) {
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)R};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)rows};
    if (tma_encode(&tm, VRCOC_BF16, d, 2, dims, strides, box, true)) return 1;
    const double foot_mb = 8.0;
    const int kslabs = K / 64;
    const int nboxes = (int)(foot_mb * 1e6 / (rows * 128));
    for (int by_lane : {0, 1})
    for (int issuers : {1, 2, 4}) {
      for (int depth : {2}) {
        if (issuers * depth * rows * 128 > 190 * 1024) continue;
        const int spread = 1, iters = 2000;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        const int smem = issuers * depth * rows * 128 + 1024;
        bw_kernel<<<sms, 128, smem>>>(tm, depth, rows, 200, nboxes, spread, kslabs, issuers, by_lane);
        cudaEventRecord(e0);
        bw_kernel<<<sms, 128, smem>>>(tm, depth, rows, iters, nboxes, spread, kslabs, issuers, by_lane);
        cudaEventRecord(e1);
        cudaError_t err = cudaDeviceSynchronize();
        if (err != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(err)); return 1; }
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double per_sm = (double)issuers * iters * rows * 128 / (ms * 1e-3) / 1e9;
        printf("%8.1f %6d %6d %7d %s | %10.1f %12.2f   (%.3f us per box per issuer)\n", foot_mb, rows, depth, issuers, by_lane ? "lanes" : "warps", per_sm, per_sm * sms / 1e3,
               ms * 1e3 / iters);
      }
    }
  }
  return 0;
}
