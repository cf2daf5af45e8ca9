// dsk_kernels.cuh — shared device helpers of the sm_100a decode path (async-copy / mbarrier wrappers, quant traits,
// activation layouts) plus the two upload/fallback utility kernels.  The decode path itself is the persistent interpreter
// in dsk_mega.cuh; the first-generation per-stage kernels that used to live here were removed in round 2 so that every
// C-ABI test hook exercises the code that produces the benchmark numbers.
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace dsk {

enum { Q_F32 = 0, Q_F16 = 1, Q_F8 = 2, Q_Q2K = 3, Q_Q3K = 4 };
enum { EPI_STORE = 0, EPI_RESID = 1, EPI_GLU = 2, EPI_KVB = 3, EPI_LOGITS = 4, EPI_PARTIAL = 5 };   // PARTIAL: tensor-parallel partial sum -> exchange buffer

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kQ2Bytes = 84;    // block_q2_K on disk and on device (src/quant.h:41-52)
constexpr int kQ3Disk = 110;    // block_q3_K on disk (src/quant.h:70-76)
constexpr int kQ3Bytes = 112;   // device repack: 110 B + 2 B pad -> 16-byte aligned blocks
constexpr int kMaxJobs = 9;
// F8E5M2 rows are stored on the device with a pitch of roundup(n, 128) + 64 bytes (re-pitched at upload; disk format
// untouched): a tile of consecutive rows copied by ONE TMA bulk copy then has a shared-memory row pitch of 64 mod 128, so
// the 16-byte A-fragment loads of a quarter warp (2 rows x 4 chunks) cover all 32 banks exactly once (measured with
// tools/mmarows_bench.cu: 8 warps, 174 vs 235 cycles per 64-column group against a pitch of n + 16).
__host__ __device__ constexpr inline size_t f8_pitch(size_t n) { return ((n + 127) & ~(size_t)127) + 64; }

// control words living in device memory so one CUDA graph serves every token
struct Ctrl {
  int token, pos, kv_sink, kv_pos, kv_len;
  int pad[3];
  unsigned long long argmax_key;  // (orderable(logit) << 32) | (0xFFFFFFFF - index)
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// ---- sm_100a async machinery: 1-D TMA bulk copies (UBLKCP) completing on an mbarrier, PDL, explicit LDS ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// cp.async.bulk global -> shared, completion counted in bytes on `bar` (src/dst 16-byte aligned, size % 16 == 0)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a));
  return r;
}
// 16-byte shared load executed only by lanes with p set (the others get zeros and generate no shared-memory wavefronts)
__device__ __forceinline__ uint4 lds128_pred(uint32_t a, uint32_t p) {
  uint4 r;
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %5, 0;\n\tmov.u32 %0, 0;\n\tmov.u32 %1, 0;\n\tmov.u32 %2, 0;\n\tmov.u32 %3, 0;\n\t"
               "@q ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];\n\t}"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a), "r"(p));
  return r;
}
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
  uint32_t r;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(a));
  return r;
}
__device__ __forceinline__ float h2f(uint16_t h) { return __half2float(__ushort_as_half(h)); }
// f8e5m2 byte = top byte of an fp16 (src/codec.h:40-48): two bytes -> half2 -> float2
__device__ __forceinline__ float2 f8x2_lo(uint32_t w) {
  uint32_t u = __byte_perm(w, 0, 0x1404);
  return __half22float2(*reinterpret_cast<__half2*>(&u));
}
__device__ __forceinline__ float2 f8x2_hi(uint32_t w) {
  uint32_t u = __byte_perm(w, 0, 0x3424);
  return __half22float2(*reinterpret_cast<__half2*>(&u));
}
__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + expf(-x)); }                      // infer.cpp:640
__device__ __forceinline__ float gelu_f(float x) {                                                      // infer.cpp:636
  return 0.5f * x * (1.0f + tanhf(0.797885f * (x + 0.044715f * x * x * x)));
}
__device__ __forceinline__ unsigned int orderable(float f) {
  unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}


// Q8_K activation blocks in shared memory (quantize_row_q8_K_ref, src/quant.cpp:616-653; staged by stage_q8 in dsk_mega.cuh)
struct Q8Smem {
  int8_t* qs;    // n int8 (16-byte aligned)
  float* d;      // n/256
  short* bsums;  // n/256*16
};

template <int Q> struct QTraits;
template <> struct QTraits<Q_F32> { static constexpr bool kq = false; static constexpr int epc = 4;  static __host__ __device__ size_t row_bytes(int n) { return (size_t)n * 4; } };
template <> struct QTraits<Q_F16> { static constexpr bool kq = false; static constexpr int epc = 8;  static __host__ __device__ size_t row_bytes(int n) { return (size_t)n * 2; } };
template <> struct QTraits<Q_F8>  { static constexpr bool kq = false; static constexpr int epc = 16; static __host__ __device__ size_t row_bytes(int n) { return f8_pitch((size_t)n); } };
template <> struct QTraits<Q_Q2K> { static constexpr bool kq = true;  static constexpr int epc = 0;  static __host__ __device__ size_t row_bytes(int n) { return (size_t)(n >> 8) * kQ2Bytes; } };
template <> struct QTraits<Q_Q3K> { static constexpr bool kq = true;  static constexpr int epc = 0;  static __host__ __device__ size_t row_bytes(int n) { return (size_t)(n >> 8) * kQ3Bytes; } };

// 16 weight bytes x EPC activations -> partial dot (F32: src/infer.cpp:121-157, F16: 161-233, F8E5M2: 238-313)
template <int Q>
__device__ __forceinline__ float chunk_dot(const uint4& wv, const float* xv) {
  float p = 0.f;
  if constexpr (Q == Q_F8) {
    float2 a;
    a = f8x2_lo(wv.x); p = fmaf(a.x, xv[0], p);  p = fmaf(a.y, xv[1], p);
    a = f8x2_hi(wv.x); p = fmaf(a.x, xv[2], p);  p = fmaf(a.y, xv[3], p);
    a = f8x2_lo(wv.y); p = fmaf(a.x, xv[4], p);  p = fmaf(a.y, xv[5], p);
    a = f8x2_hi(wv.y); p = fmaf(a.x, xv[6], p);  p = fmaf(a.y, xv[7], p);
    a = f8x2_lo(wv.z); p = fmaf(a.x, xv[8], p);  p = fmaf(a.y, xv[9], p);
    a = f8x2_hi(wv.z); p = fmaf(a.x, xv[10], p); p = fmaf(a.y, xv[11], p);
    a = f8x2_lo(wv.w); p = fmaf(a.x, xv[12], p); p = fmaf(a.y, xv[13], p);
    a = f8x2_hi(wv.w); p = fmaf(a.x, xv[14], p); p = fmaf(a.y, xv[15], p);
  } else if constexpr (Q == Q_F16) {
    float2 a;
    a = __half22float2(*reinterpret_cast<const __half2*>(&wv.x)); p = fmaf(a.x, xv[0], p); p = fmaf(a.y, xv[1], p);
    a = __half22float2(*reinterpret_cast<const __half2*>(&wv.y)); p = fmaf(a.x, xv[2], p); p = fmaf(a.y, xv[3], p);
    a = __half22float2(*reinterpret_cast<const __half2*>(&wv.z)); p = fmaf(a.x, xv[4], p); p = fmaf(a.y, xv[5], p);
    a = __half22float2(*reinterpret_cast<const __half2*>(&wv.w)); p = fmaf(a.x, xv[6], p); p = fmaf(a.y, xv[7], p);
  } else {
    p = __uint_as_float(wv.x) * xv[0];
    p = fmaf(__uint_as_float(wv.y), xv[1], p);
    p = fmaf(__uint_as_float(wv.z), xv[2], p);
    p = fmaf(__uint_as_float(wv.w), xv[3], p);
  }
  return p;
}

__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
// bytes of the staged activation vector (fp32, or Q8_K: n int8 + n/256 floats + n/256*16 shorts)
template <int Q>
__host__ __device__ inline size_t xvec_bytes(int n) {
  if (QTraits<Q>::kq) return align_up((size_t)n + (size_t)(n >> 8) * 36, 128);
  return align_up((size_t)n * 4, 128);
}
template <int Q>
__device__ __forceinline__ void carve_x(unsigned char* p, int n, float*& xs, Q8Smem& q8) {
  if constexpr (QTraits<Q>::kq) {
    q8.qs = reinterpret_cast<int8_t*>(p);
    q8.d = reinterpret_cast<float*>(p + n);
    q8.bsums = reinterpret_cast<short*>(p + n + (n >> 8) * 4);
    xs = nullptr;
  } else {
    xs = reinterpret_cast<float*>(p);
    q8.qs = nullptr; q8.d = nullptr; q8.bsums = nullptr;
  }
}

// x += partial (NCCL fallback path: after the all-reduce between kernel segments)
__global__ void add_vec_kernel(float* x, const float* p, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] += p[i];
}

// ------------------------------------------------------------------------------------------------
// RoPE helpers (src/infer.cpp:648-724).  freq[t] = 1/powf(theta, 2t/rot) is tabulated on the HOST with libm's
// powf — the same function the reference calls — because the angle pos*freq amplifies a 1-ulp difference in freq
// by `pos` (CUDA's powf is only 4-ulp accurate).  cosf/sinf here are CUDA's full-range-reduction versions.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void rope_cs(const float* __restrict__ freq, int t, int pos, float& c, float& s) {
  const float val = (float)pos * freq[t];
  c = cosf(val);
  s = sinf(val);
}

// upload-time repack of Q3_K: 110-byte disk blocks -> 112-byte device blocks
__global__ void q3k_repack_kernel(const unsigned char* src, unsigned char* dst, size_t nblocks) {
  const size_t b = (size_t)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  if (b >= nblocks) return;
  const int lane = threadIdx.x & 31;
  const unsigned char* s = src + b * kQ3Disk;
  unsigned char* d = dst + b * kQ3Bytes;
  for (int i = lane; i < kQ3Bytes; i += 32) d[i] = i < kQ3Disk ? s[i] : 0;
}

}  // namespace dsk