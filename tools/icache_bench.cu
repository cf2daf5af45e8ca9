// Instruction-fetch capacity probe for sm_100a: a straight-line block of KB kilobytes of SASS is executed R times by one
// warp (or several) per SM; cycles per instruction vs block size show where the L0 / L1 / L1.5 instruction caches end.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/icache_bench tools/icache_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

#define STEP { a = a * 1664525u + 1013904223u; b ^= a >> 7; c += b * 22695477u; d = (d << 1) ^ c; }

template <int STEPS>
__global__ void k(int reps, unsigned seed, long long* cyc, unsigned* sink) {
  unsigned a = seed + threadIdx.x, b = a * 3u, c = a ^ 5u, d = a + 7u;
  long long t0 = clock64();
#pragma unroll 1
  for (int r = 0; r < reps; r++) {
#pragma unroll
    for (int i = 0; i < STEPS; i++) STEP
  }
  long long t1 = clock64();
  if ((a ^ b ^ c ^ d) == 0x1234567u) sink[0] = a;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int STEPS>
void run(int warps, long long* d, unsigned* sink) {
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k<STEPS>);
  const int reps = 64;
  k<STEPS><<<148, warps * 32>>>(reps, 1u, d, sink);
  k<STEPS><<<148, warps * 32>>>(reps, 2u, d, sink);
  cudaDeviceSynchronize();
  long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("steps %6d (~%5.1f KB of code)  warps/CTA %d : %8.2f cycles per step (a step = 6-7 dependent instrs)\n", STEPS, STEPS * 6.5 * 16 / 1024.0, warps, (double)h / ((double)reps * STEPS));
}

int main() {
  long long* d; unsigned* sink; cudaMalloc(&d, 8); cudaMalloc(&sink, 4);
  for (int w : {1, 4}) {
    run<40>(w, d, sink); run<80>(w, d, sink); run<160>(w, d, sink); run<240>(w, d, sink); run<320>(w, d, sink); run<480>(w, d, sink);
    run<640>(w, d, sink); run<960>(w, d, sink); run<1280>(w, d, sink); run<1920>(w, d, sink); run<2560>(w, d, sink); run<3840>(w, d, sink);
  }
  return 0;
}
