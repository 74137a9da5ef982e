// Probe: how does a SWIZZLE_128B TMA box whose inner extent is < 128 bytes land in shared memory?
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/tma_box_probe tools/tma_box_probe.cu && tools/tma_box_probe
// Tensor [W | H | C] of uint16 ids (c*256 + h*32 + w); one CTA loads the box [bw | bh | 16] at (0, h0, 0) and dumps 8 KB of smem.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <vector>

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void probe(const __grid_constant__ CUtensorMap tm, int h0, uint32_t tx_bytes, uint16_t* out, int* status) {
  extern __shared__ __align__(1024) unsigned char raw[];
  unsigned char* sm = raw + ((1024u - ((uint32_t)__cvta_generic_to_shared(raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  uint16_t* s16 = reinterpret_cast<uint16_t*>(sm);
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) s16[i] = 0xFFFF;
  const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(tx_bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(sm)),
                 "l"(reinterpret_cast<uint64_t>(&tm)), "r"(0), "r"(h0), "r"(0), "r"(b)
                 : "memory");
    uint32_t done = 0;
    const long long t0 = clock64();
    while (!done && clock64() - t0 < 20000000ll)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(b) : "memory");
    *status = (int)done;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) out[i] = s16[i];
}

int main() {
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess || !fp) { printf("no encode fn\n"); return 1; }
  EncodeFn enc = (EncodeFn)fp;
  const int W = 32, H = 8, C = 16;
  std::vector<uint16_t> h(W * H * C);
  for (int c = 0; c < C; ++c) for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) h[(c * H + y) * W + x] = (uint16_t)(c * 256 + y * 32 + x);
  uint16_t *dx, *dout; int* dst;
  cudaMalloc(&dx, h.size() * 2); cudaMalloc(&dout, 8192 * 2); cudaMalloc(&dst, 4);
  cudaMemcpy(dx, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  const int cases[4][3] = {{32, 2, 1}, {16, 4, 1}, {32, 2, -1}, {16, 4, 6}};      // bw, bh, h0
  for (int k = 0; k < 4; ++k) {
    const int bw = cases[k][0], bh = cases[k][1], h0 = cases[k][2];
    CUtensorMap tm;
    cuuint64_t dims[3] = {W, H, C}, strides[2] = {W * 2, (cuuint64_t)W * H * 2};
    cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 16}, estr[3] = {1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, dx, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("== box [%d | %d | 16] at h0 = %d: encode rc %d\n", bw, bh, h0, (int)r);
    if (r != CUDA_SUCCESS) continue;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024);
    probe<<<1, 128, 40 * 1024>>>(tm, h0, (uint32_t)(bw * bh * 16 * 2), dout, dst);
    cudaError_t e = cudaDeviceSynchronize();
    printf("   launch: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 2;
    std::vector<uint16_t> o(8192); int st = 0;
    cudaMemcpy(o.data(), dout, 8192 * 2, cudaMemcpyDeviceToHost); cudaMemcpy(&st, dst, 4, cudaMemcpyDeviceToHost);
    int written = 0, last = -1;
    for (int i = 0; i < 8192; ++i) if (o[i] != 0xFFFF) { ++written; last = i; }
    printf("   barrier completed with the dense byte count: %d; elements written %d (dense %d); last written element index %d\n", st, written,
           bw * bh * 16, last);
    // first 4 smem rows of 128 B (64 elements), as (c,h,w) of the first element of each 16-byte chunk
    for (int row = 0; row < 10; ++row) {
      printf("   smem row %2d:", row);
      for (int ch = 0; ch < 8; ++ch) {
        uint16_t v = o[row * 64 + ch * 8];
        if (v == 0xFFFF) printf("  ........"); else printf("  c%02dh%dw%02d", v >> 8, (v >> 5) & 7, v & 31);
      }
      printf("\n");
    }
  }
  return 0;
}
