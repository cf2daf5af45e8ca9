// Standalone timing of the interpreter's F8 tensor-core row kernel (mma_rows_f8, csrc/dsk_mega.cuh) on data already in
// shared memory: cycles per 64-column group per warp, for 1..8 consumer warps per CTA and two row pitches.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -Ideepseek.cpp_b200/csrc -o tools/mmarows_bench tools/mmarows_bench.cu
#include <cstdio>
#include "dsk_mega.cuh"
using namespace dsk;

__global__ void __launch_bounds__(288, 1) k(int n, int pitch, int warps, int reps, int use_scale, long long* cyc, float* sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (200 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u ^ (i * 2654435761u & 0x03000300u);
  __syncthreads();
  if (tid == 0) { for (int i = 0; i < 8; i++) reinterpret_cast<uint32_t*>(smem + kHdrZero)[i] = 0u; for (int i = 0; i < 4; i++) reinterpret_cast<float*>(smem + kHdrOne)[i] = 1.0f; }
  __syncthreads();
  const uint32_t base = smem_u32(smem);
  const X16 x = carve_x16(smem + 8192, n);
  const uint32_t tile = base + 8192 + (uint32_t)x16_bytes(n) + (uint32_t)warp * 16u * (uint32_t)pitch % (150u * 1024u);
  const int gid = (tid & 31) >> 2;
  float acc = 0.f;
  long long t0 = 0, t1 = 0;
  if (warp < warps) {
    t0 = clock64();
    for (int r = 0; r < reps; r++) {
      const float2 v = mma_rows_f8(tile + gid * pitch, tile + (gid + 8) * pitch, use_scale ? base + 1024 : 0u, use_scale ? base + 1024 : 0u, 7, 0, n, x.hi, x.lo, x.gs);
      acc += v.x + v.y;
    }
    t1 = clock64();
  }
  if (acc == 123.f) sink[0] = acc;
  if (tid == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
  long long* d; float* sink; cudaMalloc(&d, 8); cudaMalloc(&sink, 4);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  const int n = 2048, reps = 64;
  for (int pitch : {2048 + 16, 2048 + 64}) for (int sc : {0, 1}) for (int w : {1, 2, 4, 6, 8}) {
    k<<<148, 288, 220 * 1024>>>(n, pitch, w, reps, sc, d, sink);
    k<<<148, 288, 220 * 1024>>>(n, pitch, w, reps, sc, d, sink);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("pitch %d scale %d warps %d: %7.1f cycles per 64-col group per warp (%s)\n", pitch, sc, w, (double)h / (reps * (n / 64)), cudaGetErrorString(e));
  }
  return 0;
}
