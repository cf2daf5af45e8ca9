// How long does it take 148 CTAs to read the SAME small buffer out of L2 at the same moment (the activation staging of a
// decode stage), versus R replicas of it (CTA i reads replica i % R)?  Each thread issues `per` float4 loads per round.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/hotread_bench tools/hotread_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256, 1) k(const float4* buf, int nf, int replicas, size_t rstride, int reps, long long* cyc, float* sink,
                                            unsigned* counter) {
  const int tid = threadIdx.x;
  float acc = 0.f;
  long long total = 0;
  for (int r = 0; r < reps; r++) {
    // grid barrier so that every CTA starts its reads together (as after a stage barrier)
    __syncthreads();
    if (tid == 0) {
      atomicAdd(counter, 1u);
      while (atomicAdd(counter, 0u) < (unsigned)(r + 1) * gridDim.x) {}
    }
    __syncthreads();
    const float4* src = buf + (size_t)(blockIdx.x % replicas) * rstride + (size_t)(r & 1) * 0;   // same data every rep: L2 resident
    const long long t0 = clock64();
    for (int f0 = tid; f0 < nf; f0 += 256 * 8) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; u++) { const int f = f0 + u * 256; v[u] = f < nf ? __ldcg(src + f) : make_float4(0, 0, 0, 0); }
#pragma unroll
      for (int u = 0; u < 8; u++) acc += v[u].x + v[u].y + v[u].z + v[u].w;
    }
    __syncthreads();
    total += clock64() - t0;
  }
  if (acc == 123.456f) sink[0] = acc;
  if (tid == 0) atomicMax((unsigned long long*)cyc, (unsigned long long)(total / reps));
}

int main() {
  const int max_rep = 16;
  const size_t nf_max = 4096;   // float4 per replica (64 KB)
  float4* buf; cudaMalloc(&buf, max_rep * nf_max * sizeof(float4)); cudaMemset(buf, 0, max_rep * nf_max * sizeof(float4));
  long long* d; float* sink; unsigned* counter; cudaMalloc(&d, 8); cudaMalloc(&sink, 4); cudaMalloc(&counter, 4);
  for (int nf : {512, 2816}) for (int R : {1, 2, 4, 8, 16}) {
    cudaMemset(d, 0, 8); cudaMemset(counter, 0, 4);
    k<<<148, 256>>>(buf, nf, R, nf_max, 20, d, sink, counter);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("%5d float4 (%5.1f KB) per CTA, %2d replicas: slowest CTA %7.2f us per round (%s)\n", nf, nf * 16 / 1024.0, R, h / 1965.0, cudaGetErrorString(e));
  }
  return 0;
}
