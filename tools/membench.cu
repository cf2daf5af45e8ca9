// tools/membench.cu — how fast can ONE resident CTA per SM stream HBM into shared memory on B200?
// Mechanisms: (a) cp.async.bulk (1-D TMA) with `depth` copies of `bytes` in flight, (b) LDG.128 into registers,
// (c) cp.async 16 B (LDGSTS).  Prints GB/s for the whole chip.  nvcc -arch=sm_100a -O3 -o membench membench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(c) : "memory"); }
__device__ __forceinline__ void expect_tx(uint32_t b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(n) : "memory"); }
__device__ __forceinline__ void bulk(uint32_t d, const void* s, uint32_t n, uint32_t b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(s), "r"(n), "r"(b) : "memory");
}
__device__ __forceinline__ void wait(uint32_t b, uint32_t ph) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0,1,0,p;\n\t}" : "=r"(ok) : "r"(b), "r"(ph) : "memory");
}
// each CTA streams `per_cta` bytes starting at base + blockIdx*per_cta; ring of `depth` slots of `bytes`; `split` copies per slot
__global__ void k_tma(const uint8_t* base, size_t per_cta, int bytes, int depth, int split, unsigned* sink) {
  extern __shared__ __align__(128) uint8_t sm[];
  uint32_t bars = s32(sm);
  uint32_t ring = s32(sm + 1024);
  if (threadIdx.x == 0) { for (int i = 0; i < depth; i++) mbar_init(bars + 8 * i, 1); asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
  __syncthreads();
  const uint8_t* src = base + (size_t)blockIdx.x * per_cta;
  const int n = (int)(per_cta / bytes);
  unsigned acc = 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < n + depth; i++) {
      if (i >= depth) {   // consume slot (i-depth): wait, touch one word
        const int j = i - depth, sl = j % depth;
        wait(bars + 8 * sl, (j / depth) & 1);
        acc += *(volatile unsigned*)(sm + 1024 + (size_t)sl * bytes);
      }
      if (i < n) {
        const int sl = i % depth;
        expect_tx(bars + 8 * sl, bytes);
        const int part = bytes / split;
        for (int p = 0; p < split; p++) bulk(ring + sl * bytes + p * part, src + (size_t)i * bytes + (size_t)p * part, part, bars + 8 * sl);
      }
    }
    sink[blockIdx.x] = acc;
  }
}
__global__ void k_ldg(const uint4* base, size_t per_cta_vec, int unroll, unsigned* sink) {
  const uint4* src = base + (size_t)blockIdx.x * per_cta_vec;
  unsigned acc = 0;
  for (size_t i = threadIdx.x; i + (size_t)(unroll - 1) * blockDim.x < per_cta_vec; i += (size_t)unroll * blockDim.x) {
    uint4 v[8];
#pragma unroll
    for (int u = 0; u < 8; u++) if (u < unroll) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(src + i + (size_t)u * blockDim.x));
#pragma unroll
    for (int u = 0; u < 8; u++) if (u < unroll) acc += v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
  }
  if (acc == 0x12345678u) sink[blockIdx.x] = acc;
}
__global__ void k_cpasync(const uint4* base, size_t per_cta_vec, int stages, unsigned* sink) {
  extern __shared__ __align__(128) uint8_t sm[];
  const uint4* src = base + (size_t)blockIdx.x * per_cta_vec;
  const int chunk = blockDim.x;   // vec per stage
  const int n = (int)(per_cta_vec / chunk);
  unsigned acc = 0;
  for (int i = 0; i < n + stages - 1; i++) {
    if (i < n) {
      uint32_t d = s32(sm) + ((i % stages) * chunk + threadIdx.x) * 16;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src + (size_t)i * chunk + threadIdx.x));
    }
    asm volatile("cp.async.commit_group;");
    if (i >= stages - 1) {
      asm volatile("cp.async.wait_group %0;" ::"n"(3));
      acc += ((volatile unsigned*)sm)[(((i - stages + 1) % stages) * chunk + threadIdx.x) * 4];
    }
  }
  if (acc == 0x12345678u) sink[blockIdx.x] = acc;
}
int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const size_t per_cta = 24u << 20;           // 24 MB per CTA -> 3.5 GB total, far beyond L2
  uint8_t* buf; cudaMalloc(&buf, per_cta * sms); cudaMemset(buf, 1, per_cta * sms);
  unsigned* sink; cudaMalloc(&sink, 4096 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaFuncSetAttribute(k_cpasync, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  auto run = [&](const char* name, auto launch, int ctas) {
    launch(); cudaDeviceSynchronize();
    cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t er = cudaGetLastError();
    printf("%-44s %8.1f GB/s  (%.3f ms)%s\n", name, (double)per_cta * ctas / ms / 1e6, ms, er ? cudaGetErrorString(er) : "");
  };
  char nm[128];
  for (int bytes : {4096, 8192, 16384, 32768, 65536})
    for (int depth : {1, 2, 4, 6}) {
      if ((size_t)bytes * depth > 200 * 1024) continue;
      snprintf(nm, sizeof nm, "tma bulk %6d B x depth %d (1 CTA/SM)", bytes, depth);
      run(nm, [&] { k_tma<<<sms, 32, 1024 + bytes * depth>>>(buf, per_cta, bytes, depth, 1, sink); }, sms);
    }
  for (int split : {2, 4, 8, 16}) {
    snprintf(nm, sizeof nm, "tma bulk 32768 B x depth 4, %2d copies/slot", split);
    run(nm, [&] { k_tma<<<sms, 32, 1024 + 32768 * 4>>>(buf, per_cta, 32768, 4, split, sink); }, sms);
  }
  for (int cps : {2, 4}) {   // several CTAs per SM, each its own ring
    snprintf(nm, sizeof nm, "tma bulk 16384 B x depth 3, %d CTAs/SM", cps);
    run(nm, [&] { k_tma<<<sms * cps, 32, 1024 + 16384 * 3>>>(buf, per_cta / cps, 16384, 3, 1, sink); }, sms);
  }
  for (int threads : {256, 512, 1024})
    for (int unroll : {1, 2, 4, 8}) {
      snprintf(nm, sizeof nm, "ldg.128 %4d thr x unroll %d (1 CTA/SM)", threads, unroll);
      run(nm, [&] { k_ldg<<<sms, threads>>>((const uint4*)buf, per_cta / 16, unroll, sink); }, sms);
    }
  for (int cps : {2, 4, 8}) {
    snprintf(nm, sizeof nm, "ldg.128 256 thr x unroll 4, %d CTAs/SM", cps);
    run(nm, [&] { k_ldg<<<sms * cps, 256>>>((const uint4*)buf, per_cta / 16 / cps, 4, sink); }, sms);
  }
  for (int stages : {4, 8}) {
    snprintf(nm, sizeof nm, "cp.async 16B 512 thr x %d stages", stages);
    run(nm, [&] { k_cpasync<<<sms, 512, 512 * 16 * stages>>>((const uint4*)buf, per_cta / 16, stages, sink); }, sms);
  }
  return 0;
}
