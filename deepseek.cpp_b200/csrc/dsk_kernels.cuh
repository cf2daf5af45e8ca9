// dsk_kernels.cuh — sm_100a kernels of the single-batch DeepSeek decode path.
//
// Everything here is batch-1 GEMV-shaped, HBM-bound integer/fp32 work (no tensor cores on this path).
// Reference functions replaced (all /root/reference @ 8db9e56):
//   _matmul F32/F16/F8E5M2/Q2_K/Q3_K   src/infer.cpp:121-379   -> gemv_kernel<Q>
//   quantize_row_q8_K_ref              src/quant.cpp:616-653    -> stage_input_q8 (fused GEMV prologue)
//   ggml_vec_dot_q{2,3}_K_q8_K         src/quant.cpp:434-783    -> dot_q2k / dot_q3k (dp4a)
//   rmsnorm                            src/infer.cpp:601-611    -> fused GEMV prologue (IN_RMSNORM)
//   silu/gelu * up                     src/infer.cpp:636-642,866-870 -> EPI_GLU
//   residual adds / expert accumulate  src/infer.cpp:832-834,874-877,901-903 -> EPI_RESID / moe_down_kernel
//   KV-cache fp16 write                src/infer.cpp:979-1002   -> EPI_KVB + attn_kernel prologue
//   rope / rope_v3 (+fp16 sink path)   src/infer.cpp:648-724    -> attn_kernel prologue
//   attn                               src/infer.cpp:728-762    -> attn_kernel
//   softmax/sigmoid + moe_gate         src/infer.cpp:472-599    -> gate_topk_kernel
//   _copy_embedding                    src/infer.cpp:1217-1263  -> embed_kernel
//   Sampler::sample_argmax             src/sampler.cpp:28-39    -> EPI_LOGITS + token feed in embed_kernel
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace dsk {

enum { Q_F32 = 0, Q_F16 = 1, Q_F8 = 2, Q_Q2K = 3, Q_Q3K = 4 };
enum { EPI_STORE = 0, EPI_RESID = 1, EPI_GLU = 2, EPI_KVB = 3, EPI_LOGITS = 4 };

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
// sum over the whole CTA; `red` is >= 33 floats of shared memory; result broadcast to all threads
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}
__device__ __forceinline__ float block_max(float v, float* red) {
  v = warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : -3.402823466e38f;
    t = warp_max(t);
    if (threadIdx.x == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}

__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint32_t ldg_stream32(const void* p) {
  uint32_t r;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
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
// Programmatic dependent launch: let the next kernel of the graph start (and prefetch its weights) while this one
// runs; griddepcontrol.wait blocks until every prerequisite grid has completed and flushed its writes.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
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

// ------------------------------------------------------------------------------------------------
// Input staging (GEMV prologue).  Every CTA builds its own copy of the activation vector in shared
// memory — optionally RMS-normalised (rmsnorm fused into the consumer) and, for K-quant weights,
// quantised to Q8_K bit-exactly like quantize_row_q8_K_ref.
// ------------------------------------------------------------------------------------------------
struct Q8Smem {
  int8_t* qs;    // n int8 (16-byte aligned)
  float* d;      // n/256
  short* bsums;  // n/256*16
};

__device__ __forceinline__ float rms_scale(const float* __restrict__ in, int n, float eps, float* red) {
  float ss = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) { float v = in[i]; ss = fmaf(v, v, ss); }
  ss = block_sum(ss, red);
  return 1.0f / sqrtf(ss / (float)n + eps);
}

__device__ __forceinline__ void stage_input_f32(const float* __restrict__ in, int n, const float* __restrict__ norm_w,
                                                float eps, float* xs, float* red) {
  if (norm_w) {
    float sc = rms_scale(in, n, eps, red);
    for (int i = threadIdx.x; i < n; i += blockDim.x) xs[i] = __fmul_rn(__fmul_rn(in[i], sc), norm_w[i]);
  } else {
    for (int i = threadIdx.x; i < n; i += blockDim.x) xs[i] = in[i];
  }
  __syncthreads();
}

// quantize_row_q8_K_ref (src/quant.cpp:616-653), one warp per 256-block, lane owns 8 consecutive values.
// iscale = -127/max (true division), q = min(127, RNE(iscale*x)), d = max*(-1/127.f) — the compiled
// form of the reference (see oracle/dsk_oracle.c).  __f*_rn intrinsics keep nvcc from contracting.
__device__ __forceinline__ void stage_input_q8(const float* __restrict__ in, int n, const float* __restrict__ norm_w,
                                               float eps, Q8Smem q, float* red) {
  float sc = 1.0f;
  if (norm_w) sc = rms_scale(in, n, eps, red);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int nb = n >> 8;
  for (int b = warp; b < nb; b += nwarps) {
    const int base = (b << 8) + lane * 8;
    float v[8];
    const float4 a0 = *reinterpret_cast<const float4*>(in + base);
    const float4 a1 = *reinterpret_cast<const float4*>(in + base + 4);
    v[0] = a0.x; v[1] = a0.y; v[2] = a0.z; v[3] = a0.w; v[4] = a1.x; v[5] = a1.y; v[6] = a1.z; v[7] = a1.w;
    if (norm_w) {
      const float4 w0 = *reinterpret_cast<const float4*>(norm_w + base);
      const float4 w1 = *reinterpret_cast<const float4*>(norm_w + base + 4);
      const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int j = 0; j < 8; j++) v[j] = __fmul_rn(__fmul_rn(v[j], sc), ww[j]);
    }
    // first element (lowest index) holding the maximum |x|
    float amax = 0.f, mx = 0.f;
    int idx = 0x7fffffff;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      float ax = fabsf(v[j]);
      if (ax > amax) { amax = ax; mx = v[j]; idx = lane * 8 + j; }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      float oa = __shfl_xor_sync(0xffffffffu, amax, o);
      float om = __shfl_xor_sync(0xffffffffu, mx, o);
      int oi = __shfl_xor_sync(0xffffffffu, idx, o);
      if (oa > amax || (oa == amax && oi < idx)) { amax = oa; mx = om; idx = oi; }
    }
    int qv[8];
    if (amax == 0.f) {
#pragma unroll
      for (int j = 0; j < 8; j++) qv[j] = 0;
      if (lane == 0) q.d[b] = 0.f;
    } else {
      const float iscale = __fdiv_rn(-127.f, mx);
#pragma unroll
      for (int j = 0; j < 8; j++) qv[j] = min(127, __float2int_rn(__fmul_rn(iscale, v[j])));
      if (lane == 0) q.d[b] = __fmul_rn(mx, -1.0f / 127.0f);
    }
    int s = qv[0] + qv[1] + qv[2] + qv[3] + qv[4] + qv[5] + qv[6] + qv[7];
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    if ((lane & 1) == 0) q.bsums[b * 16 + (lane >> 1)] = (short)s;
    uint32_t p0 = (qv[0] & 0xff) | ((qv[1] & 0xff) << 8) | ((qv[2] & 0xff) << 16) | ((uint32_t)(qv[3] & 0xff) << 24);
    uint32_t p1 = (qv[4] & 0xff) | ((qv[5] & 0xff) << 8) | ((qv[6] & 0xff) << 16) | ((uint32_t)(qv[7] & 0xff) << 24);
    *reinterpret_cast<uint2*>(q.qs + base) = make_uint2(p0, p1);
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// Row dot products over weight tiles staged in SHARED memory by TMA.  One warp owns NACC rows that share
// one activation vector: the activation chunk is read once into registers and reused for every row.
// ------------------------------------------------------------------------------------------------
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

// wa[i]: shared address of row i; sr[i]: its f8 scale row (global, nullable); xs: shared address of the fp32 activations
template <int Q, int NACC>
__device__ __forceinline__ void dot_dense(const uint32_t (&wa)[NACC], const float* const (&sr)[NACC], int bs1, int n,
                                          uint32_t xs, int lane, float (&acc)[NACC]) {
  constexpr int EPC = QTraits<Q>::epc;
  const int nch = n / EPC;
#pragma unroll 2
  for (int c = lane; c < nch; c += 32) {
    float xv[EPC];
#pragma unroll
    for (int q = 0; q < EPC / 4; q++) {
      const uint4 t = lds128(xs + (uint32_t)(c * EPC + q * 4) * 4u);
      xv[4 * q] = __uint_as_float(t.x); xv[4 * q + 1] = __uint_as_float(t.y);
      xv[4 * q + 2] = __uint_as_float(t.z); xv[4 * q + 3] = __uint_as_float(t.w);
    }
    const int sidx = (c * EPC) / bs1;
#pragma unroll
    for (int r = 0; r < NACC; r++) {
      const uint4 wv = lds128(wa[r] + (uint32_t)c * 16u);
      const float p = chunk_dot<Q>(wv, xv);
      const float s = sr[r] ? __ldg(sr[r] + sidx) : 1.0f;
      acc[r] = fmaf(p, s, acc[r]);
    }
  }
}

// Q2_K x Q8_K row (ggml_vec_dot_q2_K_q8_K, src/quant.cpp:666-783) from a shared-memory tile.  A lane owns a quarter
// block (h = 128-half, c = 16-byte half of the 32 qs bytes): 4 sub-blocks j = 8h+2s+c, s = 0..3.  Integer part
// exact (dp4a); per-block fp32 combine as the reference: d_y*d*isum - d_y*dmin*summs.
__device__ __forceinline__ float dot_q2k(uint32_t wrow, int nb, const Q8Smem& q8, int lane) {
  float acc = 0.f;
  const int nqb = nb * 4;
  for (int base = 0; base < nqb; base += 32) {
    const int qb = base + lane;
    const bool act = qb < nqb;
    const int b = qb >> 2, h = (qb >> 1) & 1, c = qb & 1;
    int isum = 0, summs = 0;
    const uint32_t blk = wrow + (uint32_t)b * kQ2Bytes;
    if (act) {
      const uint32_t qp = blk + 16 + 32 * h + 16 * c;
      const uint32_t q0 = lds32(qp), q1 = lds32(qp + 4), q2 = lds32(qp + 8), q3 = lds32(qp + 12);
      const uint32_t sA = lds32(blk + 8 * h), sB = lds32(blk + 8 * h + 4);
      const int8_t* y = q8.qs + b * 256 + 128 * h + 16 * c;
      const short* bs = q8.bsums + b * 16 + 8 * h + c;
#pragma unroll
      for (int s = 0; s < 4; s++) {
        const int4 yv = *reinterpret_cast<const int4*>(y + 32 * s);
        int dp = __dp4a((int)((q0 >> (2 * s)) & 0x03030303u), yv.x, 0);
        dp = __dp4a((int)((q1 >> (2 * s)) & 0x03030303u), yv.y, dp);
        dp = __dp4a((int)((q2 >> (2 * s)) & 0x03030303u), yv.z, dp);
        dp = __dp4a((int)((q3 >> (2 * s)) & 0x03030303u), yv.w, dp);
        const uint32_t sw = (s < 2) ? sA : sB;
        const int sc = (sw >> (8 * ((2 * s + c) & 3))) & 0xff;
        isum += (sc & 0xF) * dp;
        summs += (sc >> 4) * (int)bs[2 * s];
      }
    }
    isum += __shfl_xor_sync(0xffffffffu, isum, 1);
    isum += __shfl_xor_sync(0xffffffffu, isum, 2);
    summs += __shfl_xor_sync(0xffffffffu, summs, 1);
    summs += __shfl_xor_sync(0xffffffffu, summs, 2);
    if (act && (lane & 3) == 0) {
      const uint32_t dm = lds32(blk + 80);
      const float yd = q8.d[b];
      const float dall = yd * h2f((uint16_t)(dm & 0xffff));
      const float dmin = yd * h2f((uint16_t)(dm >> 16));
      acc += dall * (float)isum - dmin * (float)summs;
    }
  }
  return acc;
}

// Q3_K x Q8_K row (ggml_vec_dot_q3_K_q8_K, src/quant.cpp:434-614) on 112-byte repacked blocks
// [hmask 32 | qs 64 | scales 12 | d 2 | pad 2].  q = (low2 | hbit<<2) - 4  =>  dot = dp4a(low2|hbit<<2, y) - 4*bsum.
__device__ __forceinline__ float dot_q3k(uint32_t wrow, int nb, const Q8Smem& q8, int lane) {
  float acc = 0.f;
  const int nqb = nb * 4;
  for (int base = 0; base < nqb; base += 32) {
    const int qb = base + lane;
    const bool act = qb < nqb;
    const int b = qb >> 2, h = (qb >> 1) & 1, c = qb & 1;
    int isum = 0;
    const uint32_t blk = wrow + (uint32_t)b * kQ3Bytes;
    if (act) {
      const uint4 hm = lds128(blk + 16 * c);
      const uint4 qq = lds128(blk + 32 + 32 * h + 16 * c);
      const uint32_t s0 = lds32(blk + 96), s1 = lds32(blk + 100), s2 = lds32(blk + 104);
      const int8_t* y = q8.qs + b * 256 + 128 * h + 16 * c;
      const short* bs = q8.bsums + b * 16 + 8 * h + c;
#pragma unroll
      for (int s = 0; s < 4; s++) {
        const int bit = 4 * h + s;
        const int4 yv = *reinterpret_cast<const int4*>(y + 32 * s);
        int dp = __dp4a((int)(((qq.x >> (2 * s)) & 0x03030303u) | (((hm.x >> bit) & 0x01010101u) << 2)), yv.x, 0);
        dp = __dp4a((int)(((qq.y >> (2 * s)) & 0x03030303u) | (((hm.y >> bit) & 0x01010101u) << 2)), yv.y, dp);
        dp = __dp4a((int)(((qq.z >> (2 * s)) & 0x03030303u) | (((hm.z >> bit) & 0x01010101u) << 2)), yv.z, dp);
        dp = __dp4a((int)(((qq.w >> (2 * s)) & 0x03030303u) | (((hm.w >> bit) & 0x01010101u) << 2)), yv.w, dp);
        dp -= 4 * (int)bs[2 * s];
        // 6-bit scale j = 8h+2s+c (src/quant.cpp:346-349): lo4 from scales[(2s+c)], hi2 from scales[8+(2s+c)%4]
        const int t = 2 * s + c;                       // 0..7
        const uint32_t lw = (t < 4) ? s0 : s1;
        const int lob = (lw >> (8 * (t & 3))) & 0xff;
        const int lo4 = h ? (lob >> 4) : (lob & 0xF);
        const int hib = (s2 >> (8 * (t & 3))) & 0xff;
        const int hi2 = (hib >> (2 * (2 * h + (t >> 2)))) & 3;
        isum += ((lo4 | (hi2 << 4)) - 32) * dp;
      }
    }
    isum += __shfl_xor_sync(0xffffffffu, isum, 1);
    isum += __shfl_xor_sync(0xffffffffu, isum, 2);
    if (act && (lane & 3) == 0) {
      const uint32_t dw = lds32(blk + 108);
      acc += (h2f((uint16_t)(dw & 0xffff)) * q8.d[b]) * (float)isum;
    }
  }
  return acc;
}

// NACC rows sharing one activation vector -> NACC warp-reduced dot products (valid in every lane)
template <int Q, int NACC>
__device__ __forceinline__ void rows_dot(const uint32_t (&wa)[NACC], const float* const (&sr)[NACC], int bs1, int n,
                                         uint32_t xs, const Q8Smem& q8, int lane, float (&out)[NACC]) {
#pragma unroll
  for (int r = 0; r < NACC; r++) out[r] = 0.f;
  if constexpr (QTraits<Q>::kq) {
#pragma unroll
    for (int r = 0; r < NACC; r++) out[r] = (Q == Q_Q2K) ? dot_q2k(wa[r], n >> 8, q8, lane) : dot_q3k(wa[r], n >> 8, q8, lane);
  } else {
    dot_dense<Q, NACC>(wa, sr, bs1, n, xs, lane, out);
  }
#pragma unroll
  for (int r = 0; r < NACC; r++) out[r] = warp_sum(out[r]);
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
constexpr int kSmemHdr = 384;  // [0,16) two mbarriers, [64,320) reduction scratch

// ------------------------------------------------------------------------------------------------
// gemv_kernel: up to kMaxJobs weight matrices sharing ONE input vector.  Each CTA owns `rows_per_cta` consecutive
// rows of one job: that slice of the weight matrix (and of the paired `up` matrix for EPI_GLU) is pulled into
// shared memory with 1-D TMA bulk copies — issued BEFORE the programmatic-dependency wait whenever the address
// does not depend on the previous kernel (everything except routed experts), so weight streaming overlaps the
// tail of the producer kernel.  After the wait the CTA stages the activation vector (RMSNorm / Q8_K fused),
// waits for its tile, and each warp reduces `rpass` rows at a time out of shared memory.
// ------------------------------------------------------------------------------------------------
struct GemvJob {
  const uint8_t* w;      // (rows, cols) row-major payload; for expert stacks: base of the local slice
  const float* scale;    // f8e5m2 block scales or null
  const uint8_t* w_b;    // EPI_GLU: the `up` matrix (w3), same shape as w
  const float* scale_b;
  float* out;            // output vector of this job
  int rows;
  int expert_slot;       // >= 0: weights of expert active_experts[expert_slot]; -1: plain matrix
  long long w_stride;    // bytes per expert
  long long s_stride;    // scale floats per expert
};

struct GemvArgs {
  const float* in;       // input vector (n floats)
  const float* norm_w;   // fused RMSNorm weight (nullable)
  float eps;
  int n;
  int njobs;
  int epi;
  int rows_per_cta;
  int rpass;             // rows a warp reduces at once (1, 2 or 4)
  int bs0, bs1;
  int act_silu;
  const int* active_experts;  // device list of routed expert ids (jobs with expert_slot >= 0)
  int expert_first, expert_count;  // this rank's expert range [first, first+count)
  const Ctrl* ctrl;
  // EPI_KVB: kv_b rows -> fp16 K(nope)/V cache row kv_pos  (src/infer.cpp:979-1002)
  __half* kcache; __half* vcache; int n_heads, nope, vh, hd;
  // EPI_LOGITS
  Ctrl* ctrl_rw;
  int cta_begin[kMaxJobs + 1];
  GemvJob job[kMaxJobs];
};

template <int Q, int R, bool GLU>
__device__ __forceinline__ void gemv_rows(const GemvArgs& a, const GemvJob& jb, int r0, int nrows, uint32_t tile,
                                          uint32_t part_stride, const float* sc, const float* scb, uint32_t xs,
                                          const Q8Smem& q8, unsigned long long& best) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t rb = (uint32_t)QTraits<Q>::row_bytes(a.n);
  const int ncb = (a.n + a.bs1 - 1) / a.bs1;
  constexpr int NACC = R * (GLU ? 2 : 1);
  for (int g = warp * R; g < nrows; g += kWarps * R) {
    uint32_t wa[NACC];
    const float* sr[NACC];
#pragma unroll
    for (int i = 0; i < R; i++) {
      const int lr = min(g + i, nrows - 1);  // ragged tail: recompute the last row, never store it twice
      wa[i] = tile + (uint32_t)lr * rb;
      sr[i] = sc ? sc + (size_t)((r0 + lr) / a.bs0) * ncb : nullptr;
      if constexpr (GLU) {
        wa[R + i] = tile + part_stride + (uint32_t)lr * rb;
        sr[R + i] = scb ? scb + (size_t)((r0 + lr) / a.bs0) * ncb : nullptr;
      }
    }
    float v[NACC];
    rows_dot<Q, NACC>(wa, sr, a.bs1, a.n, xs, q8, lane, v);
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < R; i++) {
        if (g + i >= nrows) break;
        const int r = r0 + g + i;
        float val = v[i];
        if constexpr (GLU) val = (a.act_silu ? silu_f(val) : gelu_f(val)) * v[R + i];
        switch (a.epi) {
          case EPI_RESID: jb.out[r] = jb.out[r] + val; break;
          case EPI_KVB: {
            jb.out[r] = val;
            const int per = a.nope + a.vh, hh = r / per, ii = r - hh * per;
            const int kv_pos = a.ctrl->kv_pos;
            if (ii < a.nope) a.kcache[(size_t)kv_pos * a.n_heads * a.hd + hh * a.hd + ii] = __float2half_rn(val);
            else a.vcache[(size_t)kv_pos * a.n_heads * a.vh + hh * a.vh + (ii - a.nope)] = __float2half_rn(val);
            break;
          }
          case EPI_LOGITS: {
            jb.out[r] = val;
            const unsigned long long key = ((unsigned long long)orderable(val) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)r);
            if (key > best) best = key;
            break;
          }
          default: jb.out[r] = val; break;
        }
      }
    }
  }
}

template <int Q>
__global__ void __launch_bounds__(kThreads) gemv_kernel(const __grid_constant__ GemvArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t bar = smem_u32(smem);
  float* red = reinterpret_cast<float*>(smem + 64);
  float* xsp;
  Q8Smem q8;
  carve_x<Q>(smem + kSmemHdr, a.n, xsp, q8);
  const uint32_t tile = smem_u32(smem + kSmemHdr + xvec_bytes<Q>(a.n));
  const bool glu = a.epi == EPI_GLU;

  int j = 0;
  while (j + 1 < a.njobs && (int)blockIdx.x >= a.cta_begin[j + 1]) j++;
  const GemvJob& jb = a.job[j];
  const int r0 = ((int)blockIdx.x - a.cta_begin[j]) * a.rows_per_cta;
  const int nrows = min(a.rows_per_cta, jb.rows - r0);
  const size_t rb = QTraits<Q>::row_bytes(a.n);
  const uint32_t part_bytes = (uint32_t)align_up((size_t)nrows * rb, 16);
  const uint32_t part_stride = (uint32_t)align_up((size_t)a.rows_per_cta * rb, 128);

  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_proxy_async(); }
  __syncthreads();
  const bool dyn = jb.expert_slot >= 0;
  if (!dyn && threadIdx.x == 0) {   // static weights: stream them in before the dependency resolves
    mbar_expect_tx(bar, part_bytes * (glu ? 2u : 1u));
    bulk_g2s(tile, jb.w + (size_t)r0 * rb, part_bytes, bar);
    if (glu) bulk_g2s(tile + part_stride, jb.w_b + (size_t)r0 * rb, part_bytes, bar);
  }
  pdl_launch_dependents();
  pdl_wait();

  const float* sc = jb.scale;
  const float* scb = jb.scale_b;
  if (dyn) {
    const int e = a.active_experts[jb.expert_slot] - a.expert_first;
    if (e < 0 || e >= a.expert_count) return;  // expert lives on another rank
    if (sc) sc += (size_t)e * jb.s_stride;
    if (scb) scb += (size_t)e * jb.s_stride;
    if (threadIdx.x == 0) {
      mbar_expect_tx(bar, part_bytes * (glu ? 2u : 1u));
      bulk_g2s(tile, jb.w + (size_t)e * jb.w_stride + (size_t)r0 * rb, part_bytes, bar);
      if (glu) bulk_g2s(tile + part_stride, jb.w_b + (size_t)e * jb.w_stride + (size_t)r0 * rb, part_bytes, bar);
    }
  }
  if constexpr (QTraits<Q>::kq) stage_input_q8(a.in, a.n, a.norm_w, a.eps, q8, red);
  else stage_input_f32(a.in, a.n, a.norm_w, a.eps, xsp, red);
  mbar_wait(bar, 0);

  const uint32_t xs = QTraits<Q>::kq ? 0u : smem_u32(xsp);
  unsigned long long best = 0ull;
  if (glu) {
    if (a.rpass >= 2) gemv_rows<Q, 2, true>(a, jb, r0, nrows, tile, part_stride, sc, scb, xs, q8, best);
    else gemv_rows<Q, 1, true>(a, jb, r0, nrows, tile, part_stride, sc, scb, xs, q8, best);
  } else {
    if (a.rpass >= 4) gemv_rows<Q, 4, false>(a, jb, r0, nrows, tile, part_stride, sc, scb, xs, q8, best);
    else if (a.rpass >= 2) gemv_rows<Q, 2, false>(a, jb, r0, nrows, tile, part_stride, sc, scb, xs, q8, best);
    else gemv_rows<Q, 1, false>(a, jb, r0, nrows, tile, part_stride, sc, scb, xs, q8, best);
  }
  if (a.epi == EPI_LOGITS && (threadIdx.x & 31) == 0 && best) atomicMax(&a.ctrl_rw->argmax_key, best);
}

// ------------------------------------------------------------------------------------------------
// moe_down_kernel: x[i] += sum_k w_k * (w2[e_k][i,:] . hb_k) + shared_w2[i,:] . hb_shared
// (src/infer.cpp:873-877, 899-903; dense layers: K = 0 and the "shared" matrix is the dense w2, 926-930).
// Each CTA owns `rows_per_cta` output rows: the shared/dense slice is TMA-prefetched before the dependency wait,
// the K routed slices right after it (their address needs the gate's top-K).  One warp per output row walks the
// K+1 segments in the reference's accumulation order.
// ------------------------------------------------------------------------------------------------
struct DownArgs {
  const uint8_t* w2; const float* s2; long long w_stride, s_stride;  // routed stack (local slice)
  const uint8_t* sw2; const float* ss2;                              // shared / dense down projection (nullable)
  const float* hb;         // K x mi GLU outputs
  const float* hb_shared;  // sh
  const int* active; const float* weights;
  int K, mi, sh, dim;
  int bs0, bs1;
  int expert_first, expert_count;
  float* x;       // residual stream, updated in place when partial == null
  float* partial; // multi-GPU: write the local partial sum here instead (then all-reduce + add)
  int add_shared; // multi-GPU: only rank 0 adds the shared expert
  int rows_per_cta;
};

template <int Q>
__global__ void __launch_bounds__(kThreads) moe_down_kernel(const __grid_constant__ DownArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t bar_s = smem_u32(smem), bar_r = smem_u32(smem) + 8;
  float* red = reinterpret_cast<float*>(smem + 64);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row0 = blockIdx.x * a.rows_per_cta;
  const int nrows = min(a.rows_per_cta, a.dim - row0);
  const size_t rb_mi = QTraits<Q>::row_bytes(a.mi), rb_sh = QTraits<Q>::row_bytes(a.sh);
  const bool use_shared = a.sw2 != nullptr && a.add_shared;
  // shared memory: [hdr][x_0 .. x_{K-1}][x_shared][tile_0 .. tile_{K-1}][tile_shared]
  unsigned char* p = smem + kSmemHdr;
  float* xs[kMaxJobs];
  Q8Smem q8[kMaxJobs];
  for (int k = 0; k <= a.K; k++) {
    const int n = k < a.K ? a.mi : a.sh;
    carve_x<Q>(p, n, xs[k], q8[k]);
    p += n ? xvec_bytes<Q>(n) : 0;
  }
  const uint32_t tiles = smem_u32(p);
  const uint32_t stride_mi = (uint32_t)align_up((size_t)a.rows_per_cta * rb_mi, 128);
  const uint32_t tile_sh = tiles + (uint32_t)a.K * stride_mi;

  if (threadIdx.x == 0) { mbar_init(bar_s, 1); mbar_init(bar_r, 1); fence_proxy_async(); }
  __syncthreads();
  if (use_shared && threadIdx.x == 0) {
    const uint32_t bytes = (uint32_t)align_up((size_t)nrows * rb_sh, 16);
    mbar_expect_tx(bar_s, bytes);
    bulk_g2s(tile_sh, a.sw2 + (size_t)row0 * rb_sh, bytes, bar_s);
  }
  pdl_launch_dependents();
  pdl_wait();

  int nlocal = 0;
  for (int k = 0; k < a.K; k++) {
    const int e = a.active[k] - a.expert_first;
    if (e >= 0 && e < a.expert_count) nlocal++;
  }
  if (nlocal && threadIdx.x == 0) {
    const uint32_t bytes = (uint32_t)align_up((size_t)nrows * rb_mi, 16);
    mbar_expect_tx(bar_r, bytes * (uint32_t)nlocal);
    for (int k = 0; k < a.K; k++) {
      const int e = a.active[k] - a.expert_first;
      if (e < 0 || e >= a.expert_count) continue;
      bulk_g2s(tiles + (uint32_t)k * stride_mi, a.w2 + (size_t)e * a.w_stride + (size_t)row0 * rb_mi, bytes, bar_r);
    }
  }
  for (int k = 0; k <= a.K; k++) {
    const int n = k < a.K ? a.mi : a.sh;
    if (n == 0) continue;
    if (k < a.K) {
      const int e = a.active[k] - a.expert_first;
      if (e < 0 || e >= a.expert_count) continue;
    } else if (!use_shared) continue;
    const float* src = k < a.K ? a.hb + (size_t)k * a.mi : a.hb_shared;
    if constexpr (QTraits<Q>::kq) stage_input_q8(src, n, nullptr, 0.f, q8[k], red);
    else stage_input_f32(src, n, nullptr, 0.f, xs[k], red);
  }
  __syncthreads();
  if (nlocal) mbar_wait(bar_r, 0);
  if (use_shared) mbar_wait(bar_s, 0);

  const int ncb_mi = (a.mi + a.bs1 - 1) / a.bs1, ncb_sh = (a.sh + a.bs1 - 1) / a.bs1;
  for (int li = warp; li < nrows; li += kWarps) {
    const int i = row0 + li;
    float acc = a.partial ? 0.f : a.x[i];
    for (int k = 0; k < a.K; k++) {
      const int e = a.active[k] - a.expert_first;
      if (e < 0 || e >= a.expert_count) continue;
      const uint32_t wa[1] = {tiles + (uint32_t)k * stride_mi + (uint32_t)li * (uint32_t)rb_mi};
      const float* const sr[1] = {a.s2 ? a.s2 + (size_t)e * a.s_stride + (size_t)(i / a.bs0) * ncb_mi : nullptr};
      float v[1];
      rows_dot<Q, 1>(wa, sr, a.bs1, a.mi, QTraits<Q>::kq ? 0u : smem_u32(xs[k]), q8[k], lane, v);
      acc = fmaf(v[0], a.weights[k], acc);
    }
    if (use_shared) {
      const uint32_t wa[1] = {tile_sh + (uint32_t)li * (uint32_t)rb_sh};
      const float* const sr[1] = {a.ss2 ? a.ss2 + (size_t)(i / a.bs0) * ncb_sh : nullptr};
      float v[1];
      rows_dot<Q, 1>(wa, sr, a.bs1, a.sh, QTraits<Q>::kq ? 0u : smem_u32(xs[a.K]), q8[a.K], lane, v);
      acc += v[0];
    }
    if (lane == 0) { if (a.partial) a.partial[i] = acc; else a.x[i] = acc; }
  }
}

// x += partial (after the all-reduce)
__global__ void add_vec_kernel(float* x, const float* p, int n) {
  pdl_launch_dependents();
  pdl_wait();
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

// ------------------------------------------------------------------------------------------------
// attn_kernel: one CTA per head (BlockMHA::_attention_impl tail, src/infer.cpp:956-1045).
//   prologue: RoPE q_pe (in place in `q`), RoPE k_pe -> fp16 K cache row kv_pos, re-rotate sink keys
//   body: scores = q.K/sqrt(head_dim) over kv_len, softmax, out = att.V   (attn, src/infer.cpp:728-762)
// ------------------------------------------------------------------------------------------------
struct AttnArgs {
  float* q;            // n_heads*hd (q_pe rotated in place)
  const float* kv_a;   // kv_lora + rope (k_pe un-rotated at [kv_lora:])
  __half* kcache; __half* vcache;
  float* out;          // n_heads*vh
  const Ctrl* ctrl;
  int n_heads, hd, nope, rope, vh, kv_lora;
  const float* rope_freq;  // rope/2 host-tabulated frequencies
  int is_v3;
  int max_seq;
  int do_prologue;     // 0 for the dsk_attn test hook (plain attn over a given cache)
  int kv_len_fixed;    // used when ctrl == null
};

__global__ void __launch_bounds__(kThreads) attn_kernel(const __grid_constant__ AttnArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* red = reinterpret_cast<float*>(smem);
  float* stage = reinterpret_cast<float*>(smem + 256);        // 512 floats: sink re-rotation staging
  float* qs = reinterpret_cast<float*>(smem + 256 + 2048);    // hd
  float* att = qs + ((a.hd + 3) & ~3);                        // kv_len, then kThreads partials
  const int h = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  pdl_launch_dependents();
  pdl_wait();
  const int pos = a.ctrl ? a.ctrl->pos : 0;
  const int kv_pos = a.ctrl ? a.ctrl->kv_pos : 0;
  const int kv_len = a.ctrl ? a.ctrl->kv_len : a.kv_len_fixed;
  const int kv_sink = a.ctrl ? a.ctrl->kv_sink : 0;
  float* qh = a.q + (size_t)h * a.hd;
  const size_t kstride = (size_t)a.n_heads * a.hd, vstride = (size_t)a.n_heads * a.vh;

  for (int i = tid; i < a.nope; i += blockDim.x) qs[i] = qh[i];
  if (a.do_prologue) {
    const int half_r = a.rope >> 1;
    if (tid < half_r) {                       // q_pe
      float c, s; rope_cs(a.rope_freq, tid, pos, c, s);
      const float v0 = qh[a.nope + 2 * tid], v1 = qh[a.nope + 2 * tid + 1];
      const float r0 = v0 * c - v1 * s, r1 = v0 * s + v1 * c;
      if (a.is_v3) { qs[a.nope + 2 * tid] = r0; qs[a.nope + 2 * tid + 1] = r1; }
      else { qs[a.nope + tid] = r0; qs[a.nope + tid + half_r] = r1; }
    } else if (tid >= 64 && tid < 64 + half_r) {  // k_pe -> cache row kv_pos of this head
      const int t = tid - 64;
      float c, s; rope_cs(a.rope_freq, t, pos, c, s);
      const float v0 = a.kv_a[a.kv_lora + 2 * t], v1 = a.kv_a[a.kv_lora + 2 * t + 1];
      const float r0 = v0 * c - v1 * s, r1 = v0 * s + v1 * c;
      __half* kr = a.kcache + (size_t)kv_pos * kstride + (size_t)h * a.hd + a.nope;
      if (a.is_v3) { kr[2 * t] = __float2half_rn(r0); kr[2 * t + 1] = __float2half_rn(r1); }
      else { kr[t] = __float2half_rn(r0); kr[t + half_r] = __float2half_rn(r1); }
    } else if (tid >= 128 && tid < 128 + half_r * kv_sink && kv_sink > 0) {  // sink re-rotation by one position
      const int t = (tid - 128) % half_r, r = (tid - 128) / half_r;
      float c, s; rope_cs(a.rope_freq, t, 1, c, s);
      __half* kr = a.kcache + (size_t)r * kstride + (size_t)h * a.hd + a.nope;
      const float v0 = __half2float(kr[2 * t]), v1 = __half2float(kr[2 * t + 1]);
      stage[2 * (r * half_r + t)] = v0 * c - v1 * s;     // staged: the V2 layout permutes, so read all first
      stage[2 * (r * half_r + t) + 1] = v0 * s + v1 * c;
    }
    __syncthreads();
    if (tid >= 128 && tid < 128 + half_r * kv_sink && kv_sink > 0) {
      const int t = (tid - 128) % half_r, r = (tid - 128) / half_r;
      __half* kr = a.kcache + (size_t)r * kstride + (size_t)h * a.hd + a.nope;
      const float r0 = stage[2 * (r * half_r + t)], r1 = stage[2 * (r * half_r + t) + 1];
      if (a.is_v3) { kr[2 * t] = __float2half_rn(r0); kr[2 * t + 1] = __float2half_rn(r1); }
      else { kr[t] = __float2half_rn(r0); kr[t + half_r] = __float2half_rn(r1); }
    }
    if (tid < a.rope) qh[a.nope + tid] = qs[a.nope + tid];  // keep the state buffer like the reference's s.q()
  } else {
    for (int i = tid; i < a.rope; i += blockDim.x) qs[a.nope + i] = qh[a.nope + i];
  }
  __syncthreads();
  __threadfence_block();

  // scores
  const float inv = sqrtf((float)a.hd);
  for (int t = warp; t < kv_len; t += kWarps) {
    const __half* kr = a.kcache + (size_t)t * kstride + (size_t)h * a.hd;
    float s = 0.f;
    for (int i = lane * 2; i < a.hd; i += 64) {
      const float2 kk = __half22float2(*reinterpret_cast<const __half2*>(kr + i));
      s = fmaf(qs[i], kk.x, s);
      s = fmaf(qs[i + 1], kk.y, s);
    }
    s = warp_sum(s);
    if (lane == 0) att[t] = s / inv;
  }
  __syncthreads();
  // softmax (src/infer.cpp:472-487)
  float m = -3.402823466e38f;
  for (int t = tid; t < kv_len; t += blockDim.x) m = fmaxf(m, att[t]);
  m = block_max(m, red);
  float sum = 0.f;
  for (int t = tid; t < kv_len; t += blockDim.x) { const float e = expf(att[t] - m); att[t] = e; sum += e; }
  sum = block_sum(sum, red);
  for (int t = tid; t < kv_len; t += blockDim.x) att[t] = att[t] / sum;
  __syncthreads();
  // out = att . V  — thread (i, g): value dim i, token group g; groups reduced through shared memory
  const int groups = blockDim.x / a.vh > 0 ? blockDim.x / a.vh : 1;
  if (a.vh <= (int)blockDim.x) {
    const int i = tid % a.vh, g = tid / a.vh;
    float acc = 0.f;
    if (g < groups) {
      const __half* vb = a.vcache + (size_t)h * a.vh + i;
      for (int t = g; t < kv_len; t += groups) acc = fmaf(att[t], __half2float(vb[(size_t)t * vstride]), acc);
    }
    float* part = att + ((kv_len + 3) & ~3);
    if (g < groups) part[g * a.vh + i] = acc;
    __syncthreads();
    if (tid < a.vh) {
      float o = 0.f;
      for (int g2 = 0; g2 < groups; g2++) o += part[g2 * a.vh + tid];
      a.out[(size_t)h * a.vh + tid] = o;
    }
  } else {
    for (int i = tid; i < a.vh; i += blockDim.x) {
      const __half* vb = a.vcache + (size_t)h * a.vh + i;
      float acc = 0.f;
      for (int t = 0; t < kv_len; t++) acc = fmaf(att[t], __half2float(vb[(size_t)t * vstride]), acc);
      a.out[(size_t)h * a.vh + i] = acc;
    }
  }
}
inline size_t attn_smem_bytes(int hd, int vh, int max_kv) {
  return 256 + 2048 + (size_t)((hd + 3) & ~3) * 4 + (size_t)((max_kv + 3) & ~3) * 4 + (size_t)kThreads * 4 + 64;
}

// ------------------------------------------------------------------------------------------------
// gate_topk_kernel: softmax | sigmoid (+bias), greedy / group-limited-greedy top-K  (src/infer.cpp:493-599).
// One CTA, E <= 256.  Ties: lowest index wins (strict `>` scans).  The group-limited first pass compares the
// first candidate against 0.0f — the value the reference's out-of-bounds x[-1] read yields.
// ------------------------------------------------------------------------------------------------
struct GateArgs {
  float* x;            // E logits in, scores out (like the reference's in-place update)
  const float* bias;   // nullable
  int* active; float* weights;
  int E, K, norm_topk_prob, sigmoid, method, n_group, topk_group;
  float scale;
};

__device__ __forceinline__ void argmax_pair(float& v, int& i) {
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, i, o);
    if (oi >= 0 && (i < 0 || ov > v || (ov == v && oi < i))) { v = ov; i = oi; }
  }
}

__global__ void __launch_bounds__(256) gate_topk_kernel(const __grid_constant__ GateArgs a) {
  __shared__ float sx[256];
  __shared__ float red[40];
  __shared__ unsigned char mask[256];  // 1 = not selectable
  __shared__ int sel[16];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  pdl_launch_dependents();
  pdl_wait();
  float v = tid < a.E ? a.x[tid] : -3.402823466e38f;
  if (a.sigmoid) {
    v = 1.0f / (1.0f + expf(-v));
  } else {
    const float m = block_max(v, red);
    const float e = tid < a.E ? expf(v - m) : 0.f;
    const float s = block_sum(e, red);
    v = e / s;
  }
  if (a.bias && tid < a.E) v += a.bias[tid];
  if (tid < a.E) { sx[tid] = v; a.x[tid] = v; }
  mask[tid] = tid < a.E ? 0 : 1;
  __syncthreads();
  if (a.method == 1) {
    const int gs = a.E / a.n_group;
    for (int g = warp; g < a.n_group; g += 8) {
      for (int k = 0; k < a.topk_group; k++) {
        float bv = 0.f; int bi = -1;
        for (int j = g * gs + lane; j < (g + 1) * gs; j += 32) {
          if (!mask[j] && sx[j] > 0.0f && (bi < 0 || sx[j] > bv)) { bv = sx[j]; bi = j; }
        }
        argmax_pair(bv, bi);
        if (lane == 0 && bi >= 0) mask[bi] = 2;  // 2 = candidate
        __syncwarp();
      }
    }
    __syncthreads();
    if (tid < a.E) mask[tid] = (mask[tid] == 2) ? 0 : 1;
    __syncthreads();
  }
  if (warp == 0) {
    for (int k = 0; k < a.K; k++) {
      float bv = 0.f; int bi = -1;
      for (int j = lane; j < a.E; j += 32) {
        if (!mask[j] && (bi < 0 || sx[j] > bv)) { bv = sx[j]; bi = j; }
      }
      argmax_pair(bv, bi);
      if (lane == 0) { sel[k] = bi; if (bi >= 0) mask[bi] = 1; }
      __syncwarp();
    }
    if (lane == 0) {
      float wsum = 0.f;
      for (int k = 0; k < a.K; k++) wsum += sel[k] >= 0 ? sx[sel[k]] : 0.f;
      if (!a.norm_topk_prob) wsum = 1.0f;
      for (int k = 0; k < a.K; k++) {
        a.active[k] = sel[k];
        a.weights[k] = sel[k] >= 0 ? sx[sel[k]] / wsum * a.scale : 0.f;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// embed_kernel: x = dequant(embedding row `token`)  (Model::_copy_embedding, src/infer.cpp:1217-1263).
// In device-loop mode the token is decoded from the previous step's argmax key, and the step counters
// (pos, kv_sink, kv_pos, kv_len — src/infer.cpp:1274-1277) are advanced here, so the whole token is one graph.
// ------------------------------------------------------------------------------------------------
struct EmbedArgs {
  const uint8_t* table; const float* scale;
  float* x;
  Ctrl* ctrl;
  int quant, dim, bs0, bs1;
  int from_argmax;     // 1: token = argmax of the previous logits, then pos++ (device-resident decode loop)
  int original_max;    // rope_scaling_original_max_position_embeddings
  int* token_log;      // nullable: generated tokens, indexed by step
  int* step;           // nullable device step counter
};

__global__ void __launch_bounds__(256) embed_kernel(const __grid_constant__ EmbedArgs a) {
  __shared__ int s_token;
  pdl_launch_dependents();
  pdl_wait();
  if (threadIdx.x == 0) {
    Ctrl* c = a.ctrl;
    int token = c->token;
    if (a.from_argmax) {
      token = (int)(0xFFFFFFFFu - (unsigned)(c->argmax_key & 0xFFFFFFFFull));
      const int pos = c->pos + 1;
      const int sink = pos >= a.original_max ? 2 : 0;  // KV_SINKS = 2 (src/model.h:14)
      c->token = token;
      c->pos = pos;
      c->kv_sink = sink;
      c->kv_pos = sink + (pos - sink) % (a.original_max - sink);
      c->kv_len = pos >= a.original_max ? a.original_max : pos + 1;
      if (a.token_log && a.step) { a.token_log[*a.step] = token; *a.step = *a.step + 1; }
    }
    c->argmax_key = 0ull;
    s_token = token;
  }
  __syncthreads();
  const int token = s_token;
  const int dim = a.dim;
  switch (a.quant) {
    case Q_F32: {
      const float* t = reinterpret_cast<const float*>(a.table) + (size_t)token * dim;
      for (int i = threadIdx.x; i < dim; i += blockDim.x) a.x[i] = t[i];
      break;
    }
    case Q_F16: {
      const __half* t = reinterpret_cast<const __half*>(a.table) + (size_t)token * dim;
      for (int i = threadIdx.x; i < dim; i += blockDim.x) a.x[i] = __half2float(t[i]);
      break;
    }
    case Q_F8: {
      const uint8_t* t = a.table + (size_t)token * f8_pitch((size_t)dim);
      const int ncb = (dim + a.bs1 - 1) / a.bs1;
      for (int i = threadIdx.x; i < dim; i += blockDim.x) {
        const float sc = a.scale[(size_t)(token / a.bs0) * ncb + i / a.bs1];
        a.x[i] = h2f((uint16_t)((uint16_t)t[i] << 8)) * sc;
      }
      break;
    }
    case Q_Q2K: {  // dequantize_row_q2_K src/quant.cpp:217-247: weight idx -> (block, 128-half, s, l)
      const int nb = dim >> 8;
      const uint8_t* row = a.table + (size_t)token * nb * kQ2Bytes;
      for (int i = threadIdx.x; i < dim; i += blockDim.x) {
        const int b = i >> 8, w = i & 255, hh = w >> 7, s = (w >> 5) & 3, l = w & 31;
        const uint8_t* blk = row + (size_t)b * kQ2Bytes;
        const int sc = blk[8 * hh + 2 * s + (l >> 4)];
        const int qv = (blk[16 + 32 * hh + l] >> (2 * s)) & 3;
        const float d = h2f(*reinterpret_cast<const uint16_t*>(blk + 80));
        const float mn = h2f(*reinterpret_cast<const uint16_t*>(blk + 82));
        a.x[i] = (d * (float)(sc & 0xF)) * (float)qv - mn * (float)(sc >> 4);
      }
      break;
    }
    default: {     // dequantize_row_q3_K src/quant.cpp:384-432 on the 112-byte repacked blocks
      const int nb = dim >> 8;
      const uint8_t* row = a.table + (size_t)token * nb * kQ3Bytes;
      for (int i = threadIdx.x; i < dim; i += blockDim.x) {
        const int b = i >> 8, w = i & 255, hh = w >> 7, s = (w >> 5) & 3, l = w & 31;
        const uint8_t* blk = row + (size_t)b * kQ3Bytes;
        const int j = 8 * hh + 2 * s + (l >> 4);
        const int lob = blk[96 + (j & 7)];
        const int lo4 = j < 8 ? (lob & 0xF) : (lob >> 4);
        const int hi2 = (blk[96 + 8 + (j & 3)] >> (2 * (j >> 2))) & 3;
        const int sc = (lo4 | (hi2 << 4)) - 32;
        const int hb = (blk[l] >> (4 * hh + s)) & 1;
        const int qv = ((blk[32 + 32 * hh + l] >> (2 * s)) & 3) - (hb ? 0 : 4);
        const float d = h2f(*reinterpret_cast<const uint16_t*>(blk + 108));
        a.x[i] = (d * (float)sc) * (float)qv;
      }
      break;
    }
  }
}

// final RMSNorm in place is fused into the LM-head GEMV prologue, but the reference also leaves the normed
// vector in s.x() (src/infer.cpp:1292); this tiny kernel keeps that contract for taps.
__global__ void __launch_bounds__(256) rmsnorm_kernel(float* out, const float* in, const float* w, int n, float eps) {
  __shared__ float red[40];
  const float sc = rms_scale(in, n, eps, red);
  for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = __fmul_rn(__fmul_rn(in[i], sc), w[i]);
}

// standalone Q8_K quantiser for the dsk_quantize_q8k test hook: writes reference-layout block_q8_K (292 B)
__global__ void __launch_bounds__(kThreads) q8k_export_kernel(const float* in, int n, unsigned char* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* red = reinterpret_cast<float*>(smem);
  Q8Smem q8;
  q8.qs = reinterpret_cast<int8_t*>(smem + 256);
  q8.d = reinterpret_cast<float*>(smem + 256 + n);
  q8.bsums = reinterpret_cast<short*>(smem + 256 + n + (n >> 8) * 4);
  stage_input_q8(in, n, nullptr, 0.f, q8, red);
  const int nb = n >> 8;
  for (int b = 0; b < nb; b++) {
    unsigned char* o = out + (size_t)b * 292;
    if (threadIdx.x == 0) *reinterpret_cast<float*>(o) = q8.d[b];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) o[4 + i] = (unsigned char)q8.qs[b * 256 + i];
    if (threadIdx.x < 16) *reinterpret_cast<short*>(o + 260 + 2 * threadIdx.x) = q8.bsums[b * 16 + threadIdx.x];
  }
}

// standalone rope for the dsk_rope test hook
__global__ void rope_test_kernel(float* vec, int d, int head_dim, int pos, const float* freq, int v3) {
  extern __shared__ float buf[];
  const int t = threadIdx.x;
  if (2 * t < d) {
    float c, s; rope_cs(freq, ((2 * t) % head_dim) / 2, pos, c, s);
    const float v0 = vec[2 * t], v1 = vec[2 * t + 1];
    if (v3) { buf[2 * t] = v0 * c - v1 * s; buf[2 * t + 1] = v0 * s + v1 * c; }
    else { buf[t] = v0 * c - v1 * s; buf[t + d / 2] = v0 * s + v1 * c; }
  }
  __syncthreads();
  for (int i = t; i < d; i += blockDim.x) vec[i] = buf[i];
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
