// Stand-alone probe: which 4-D tiled TMA loads does this GPU accept?  (box shapes / swizzle / negative coordinates)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I asy-vrnet_b200/csrc tools/tma_probe.cu -o tools/tma_probe -lcuda
#include <vector>
#include "tma.cuh"
namespace vrcoc { char* err_buf() { static char b[256]; return b; } int fail(int c, const char* f, ...) { printf("fail: %s\n", f); return c; } int check_launch(const char*) { return 0; } }
using namespace vrcoc;

__global__ void probe(const __grid_constant__ CUtensorMap tm, int x, int y, int c, int b, int bytes, uint16_t* out) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  __syncthreads();
  if (threadIdx.x == 0) { mbar_expect_tx(&bar, bytes); tma_load_4d(smem, &tm, x, y, c, b, &bar); }
  mbar_wait(&bar, 0);
  for (int i = threadIdx.x; i < bytes / 2; i += blockDim.x) out[i] = reinterpret_cast<uint16_t*>(smem)[i];
}

int main(int argc, char** argv) {
  const int only = argc > 1 ? atoi(argv[1]) : -1;
  int ci = -1;
  const int W = 32, H = 32, C = 64, B = 2;
  std::vector<uint16_t> h((size_t)W * H * C * B);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (uint16_t)(i & 0x7fff);      // bf16 bit patterns = linear index
  uint16_t *d, *o;
  cudaMalloc(&d, h.size() * 2); cudaMalloc(&o, 65536);
  cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  struct Case { int bw, bh, bc, sw, x, y; } cases[] = {
      {32, 2, 64, 0, 0, 0}, {32, 2, 64, 1, 0, 0}, {32, 2, 64, 1, -1, -1}, {16, 4, 64, 1, 0, 0}, {16, 4, 64, 1, -1, 3},
      {32, 1, 64, 1, 0, 0}, {32, 2, 32, 1, 0, 0}, {32, 2, 64, 0, -1, -1},
      {32, 2, 64, 0, 1, 0}, {32, 2, 64, 0, 8, 0}, {32, 2, 64, 0, -8, 0}, {32, 2, 64, 0, 0, -1}, {32, 2, 64, 0, 0, 31}, {32, 2, 64, 0, 24, 0},
      {32, 2, 64, 0, -1, 0}, {32, 2, 64, 0, 7, 5}};
  for (auto cs : cases) {
    if (++ci != only && only >= 0) continue;
    CUtensorMap tm;
    cuuint64_t dims[4] = {W, H, C, B};
    cuuint64_t strides[3] = {W * 2, (cuuint64_t)W * H * 2, (cuuint64_t)W * H * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)cs.bw, (cuuint32_t)cs.bh, (cuuint32_t)cs.bc, 1};
    int rc = tma_encode(&tm, VRCOC_BF16, d, 4, dims, strides, box, cs.sw != 0);
    const int bytes = cs.bw * cs.bh * cs.bc * 2;
    cudaMemset(o, 0xff, 65536);
    probe<<<1, 128, 34000>>>(tm, cs.x, cs.y, 0, 1, bytes, o);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<uint16_t> r(bytes / 2);
    if (e == cudaSuccess) cudaMemcpy(r.data(), o, bytes, cudaMemcpyDeviceToHost);
    printf("box {%d,%d,%d,1} swizzle=%d at (%d,%d): encode rc=%d, run: %s", cs.bw, cs.bh, cs.bc, cs.sw, cs.x, cs.y, rc, cudaGetErrorString(e));
    if (e == cudaSuccess) {
      printf("  first 16-byte chunks of channel rows 0,1:");
      for (int row = 0; row < 2; ++row) { printf(" |"); for (int k = 0; k < 8; ++k) printf(" %d", r[row * (cs.bw * cs.bh) + 8 * k] - (int)((size_t)W * H * C % 32768 + (size_t)row * W * H) % 32768); }
    }
    printf("\n");
    if (e != cudaSuccess) { printf("sticky error - stopping\n"); return 0; }
  }
  return 0;
}
