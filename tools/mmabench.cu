// Legacy tensor path (mma.sync.m16n8k16 f16 -> f32, SASS HMMA.16816.F32) on sm_100a: latency and issue rate per SM
// sub-partition, as a function of independent accumulator chains (ILP) and warps per sub-partition.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/mmabench tools/mmabench.cu && tools/mmabench
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

__device__ __forceinline__ void mma(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int ILP>
__global__ void k(int iters, long long* cyc, float* sink) {
  float c[ILP][4];
#pragma unroll
  for (int i = 0; i < ILP; i++) for (int j = 0; j < 4; j++) c[i][j] = 0.f;
  uint32_t a0 = threadIdx.x * 0x3c003c00u, a1 = 0x3c003800u, a2 = 0x38003c00u, a3 = 0x3c003c00u, b0 = 0x3c003c00u, b1 = 0x38003800u;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) mma(c[i], a0, a1, a2, a3, b0, b1);
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ILP; i++) for (int j = 0; j < 4; j++) s += c[i][j];
  if (s == 123.456f) sink[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int ILP>
void run(int warps) {
  long long* d; float* sink; cudaMalloc(&d, 8); cudaMalloc(&sink, 4);
  const int iters = 4096;
  k<ILP><<<148, warps * 32>>>(iters, d, sink);
  k<ILP><<<148, warps * 32>>>(iters, d, sink);
  long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  const double per_warp = (double)h / ((double)iters * ILP);
  const double per_smsp = per_warp / ((warps + 3) / 4);
  printf("warps/CTA %2d  ILP %2d : %7.2f cycles per HMMA per warp, %6.2f cycles per HMMA per sub-partition -> %7.1f TFLOP/s chip @1.965 GHz\n",
         warps, ILP, per_warp, warps >= 4 ? per_smsp : per_warp, 4096.0 / (warps >= 4 ? per_smsp : per_warp) * (warps >= 4 ? 4 : warps) * 148 * 1.965e9 / 1e12);
  cudaFree(d); cudaFree(sink);
}

int main() {
  for (int w : {1, 4, 8, 16}) { run<1>(w); run<2>(w); run<4>(w); run<8>(w); }
  return 0;
}
