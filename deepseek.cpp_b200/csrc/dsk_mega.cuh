// dsk_mega.cuh — the persistent decode interpreter ("one resident grid runs the whole token").
//
// A token is a PROGRAM: an array of Stage descriptors in device memory (built once per dsk_state).  Stages are the
// dependent phases of Model::_forward_cpu / Block::_block_cpu / BlockMHA::_attention_impl
// (/root/reference/src/infer.cpp:810-1049, 1265-1317):
//     EMBED | per layer: S1 q(+q_a)+kv_a | [S1b wq_b] | S2 kv_b->KV cache | S3 rope+attention | S4 wo+residual |
//     S5 gate logits | S56 routing + shared/routed w1,w3 with act(h1)*h3 | S7 w2 + weighted accumulate | LM head+argmax
// decode_kernel<Q> executes stages [s_begin, s_end) with ONE CTA per SM:
//   * warp 8 is the TMA producer: it walks this CTA's tiles of the current stage — and runs ahead into the NEXT
//     stage's static weights — issuing cp.async.bulk copies (weight rows + their f8 scale row) into a ring of
//     shared-memory slots guarded by full/empty mbarriers.  Only routed-expert tiles wait for the routing.
//   * warps 0-7 are consumers: at stage start they pass the grid barrier, stage the activation vector (RMSNorm /
//     bit-exact Q8_K fused), compute the routing if the stage needs it, then reduce tiles straight out of shared
//     memory (warp task = R rows x one column piece), combine pieces in a fixed order and run the epilogue.
//   * stages are separated by a grid barrier (one release-add + acquire-spin on a monotonically increasing
//     counter); launched one stage at a time (s_end = s_begin+1) the same code needs no barrier — the fallback mode.
// HBM traffic is the algorithmic weight bytes only; every byte crosses global->shared exactly once via TMA.
#pragma once

#include <cstddef>

#include "dsk_kernels.cuh"

namespace dsk {

enum { ST_EMBED = 0, ST_GEMV = 1, ST_DOWN = 2, ST_ATTN = 3, ST_XCHG = 4, ST_AMAX = 5, ST_MLA_CACHE = 6, ST_ATTN_MLA = 7 };
constexpr int kMaxRanks = 8;

constexpr int kConsumers = 256;            // warps 0..7
constexpr int kMegaThreads = 288;          // + producer warp 8
constexpr int kSlotScale = 2048;           // bytes reserved per ring slot for f8 scale rows
constexpr int kSlotData = 32 * 1024;       // weight bytes per ring slot
constexpr int kSlotBytes = kSlotScale + kSlotData;
constexpr int kRingEntries = 16;          // tiles in flight per CTA (mbarrier pairs); their bytes come out of one circular buffer
constexpr int kMaxPieces = 24;
constexpr int kMegaHdr = 8192;             // barriers, scratch, partial results, smem copies of Program + 2 Stage descriptors
constexpr int kStageSlot = 1536;           // bytes reserved per cached Stage descriptor

struct MJob {                              // one weight matrix of a GEMV stage
  const uint8_t* w; const float* scale;
  const uint8_t* w_b; const float* scale_b;  // EPI_GLU: the `up` matrix
  float* out;
  long long w_stride, s_stride;            // per-expert strides (bytes / floats)
  int rows, expert_slot, tile_begin, row_base;   // row_base: global index of local row 0 (row-sharded LM head)
};

struct Piece { int seg, g0, g1, pad; };    // column piece: granules [g0,g1) of segment `seg` (granule = 16 B or a 256-block)

struct Stage {
  int kind, quant, epi, njobs;
  int n, rows_per_tile, rpass, ntiles;
  int need_topk, npieces, layer, has_dyn;
  int use_mma, wp, down_rows, down_nbuf;   // down_nbuf: partial-sum buffers of the warp-per-tile DOWN stage (row groups in flight)   // wp: warp-per-tile stage (tensor-core tiles owned by single warps)          // F8E5M2 tiles through mma.sync (pieces in 64-column units)
  const float* in; const float* norm_w;
  MJob job[kMaxJobs];
  Piece piece[kMaxPieces];
  // ST_GEMV extras
  __half* kcache; __half* vcache;          // EPI_KVB / ST_ATTN
  float* gate_logits; const float* gate_bias;
  // ST_DOWN
  const uint8_t* w2; const float* s2; long long w2_stride, s2_stride;
  const uint8_t* sw2; const float* ss2;
  int K, mi, sh, add_shared;
  int seg_stride;                          // bytes between segments inside a ring slot (ST_DOWN)
  int xchg_ord;                            // peer-memory mode: ordinal (within a token) of the exchange this DOWN / XCHG stage belongs to
  int max_inflight;                        // tiles the producer may have outstanding in this stage (<= kRingEntries)
  int peer_stores;                         // multi-GPU: the epilogues of this stage store into peer memory (one system fence per thread at stage end)
} __attribute__((aligned(16)));

static_assert(sizeof(Stage) <= kStageSlot, "Stage descriptor must fit its shared-memory cache slot");
static_assert(sizeof(Stage) % 16 == 0, "Stage is copied in 16-byte words");

struct Program {
  // model constants
  int dim, n_heads, hd, nope, rope, vh, kv_lora, is_v3;
  int bs0, bs1, act_silu, max_seq;
  int bs1_shift, pad_i[3];
  int E, K, norm_topk_prob, sigmoid, topk_method, n_group, topk_group, original_max;
  float eps, routed_scale;
  int expert_first, expert_count;
  int embed_quant, n_stages;
  // buffers
  const uint8_t* embed_w; const float* embed_scale;
  const float* rope_freq;
  float *x, *q, *q_a, *kv_a, *kv_b, *xb2, *hbk, *hbs, *moe_logits, *moe_scores, *act_w, *logits, *partial, *att_scratch;
  float* q_c;                              // true-MLA blocks: absorbed query (n_heads x kv_lora_rank)
  int* act;
  Ctrl* ctrl;
  unsigned int* sync_counter;              // grid barrier arrivals (monotonic)
  unsigned int* sync_base;                 // value of the counter when this launch started
  int* token_log; int* step;
  int ring_bytes, xregion_bytes;           // circular tile buffer after the activation region
  int slot_data, slot_scale, slot_bytes, pad_s;   // ring slot geometry (bytes): [scale rows | weight tile]
  // multi-GPU, peer-memory mode: every rank stores its MoE partial sums straight into every peer's exchange buffer over
  // NVLink (plain stores to IPC-mapped memory), then raises a flag there; an ST_XCHG stage waits for the N flags and adds the
  // N partials in rank order — no kernel boundary, no NCCL call on the token path
  int n_ranks, rank, n_xchg, tp;           // tp: tensor-parallel program (heads / FFN slices / LM-head rows sharded)
  float* xchg_peer[kMaxRanks];             // [2][n_ranks][dim] buffer of rank q (q == rank: the local one)
  unsigned* xflag_peer[kMaxRanks];         // n_ranks flags of rank q: flag[r] = sequence number of rank r's last finished store
  float* logits_peer[kMaxRanks];           // tp: full-vocabulary logits of rank q (every rank stores its rows into every copy)
  unsigned long long* amax_peer[kMaxRanks];   // tp: [2][n_ranks] arg-max keys of rank q
  long long* route_prof;                   // profiling: 4 phase durations of the last routing (cycles)
  unsigned long long* tstamp;              // [n_stages][8] globaltimer stamps of CTA 0: start, inputs staged, tiles done, arrived, ...
  // test taps (null in production): CTA 0 dumps the activation vector exactly as the tile loop will read it
  unsigned char* dbg_q8;                   // K-quant stages: block_q8_K records (292 B each, src/quant.h:104-109)
  float* dbg_x;                            // fp32 staging: the staged vector; F8 tensor-core staging: (hi + lo) * 2^-e
  Stage stage[1];                          // n_stages entries follow
};
constexpr int kProgHdrBytes = (int)offsetof(Program, stage);
static_assert(kProgHdrBytes <= 1024 && kProgHdrBytes % 16 == 0, "Program header is cached in 1 KB of shared memory");

// ---- consumer-only synchronisation (the producer warp never joins) ---------------------------------------------
__device__ __forceinline__ void csync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ float csum(float v, float* red) {
  v = warp_sum(v);
  csync();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  csync();
  float t = red[0] + red[1] + red[2] + red[3] + red[4] + red[5] + red[6] + red[7];
  return t;
}
__device__ __forceinline__ float cmax(float v, float* red) {
  v = warp_max(v);
  csync();
  if ((threadIdx.x & 31) == 0) red[8 + (threadIdx.x >> 5)] = v;
  csync();
  float t = fmaxf(fmaxf(fmaxf(red[8], red[9]), fmaxf(red[10], red[11])), fmaxf(fmaxf(red[12], red[13]), fmaxf(red[14], red[15])));
  return t;
}
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// bounded waits are TIME based (globaltimer, checked every 4096 polls): a protocol bug must trap, not hang the GPU
constexpr unsigned long long kSpinLimitNs = 20ull * 1000ull * 1000ull * 1000ull;
__device__ __noinline__ void spin_check(unsigned long long& t0) {
  const unsigned long long now = gtime();
  if (t0 == 0ull) t0 = now;
  else if (now - t0 > kSpinLimitNs) __trap();
}
__device__ __forceinline__ void mbar_wait_guard(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  unsigned long long t0 = 0ull;
  for (unsigned spins = 0; !ok; spins++) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if ((spins & 4095u) == 4095u) spin_check(t0);
  }
}
// producer side (waiting for a ring slot to be released): same, with a back-off between polls
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  unsigned long long t0 = 0ull;
  for (unsigned spins = 0; !ok; spins++) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (!ok) __nanosleep(40);
    if ((spins & 4095u) == 4095u) spin_check(t0);
  }
}
// producer <- consumers: "inputs (and routing) of stage k are ready" as a MONOTONIC stage count in shared memory
// (an mbarrier phase bit could alias if the consumers ever got two signals ahead of the producer's wait)
__device__ __forceinline__ void dep_signal(uint32_t addr, int stage_count) {
  asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(addr), "r"(stage_count) : "memory");
}
__device__ __forceinline__ void dep_wait(uint32_t addr, int stage_count) {
  int v = 0;
  unsigned long long t0 = 0ull;
  for (unsigned spins = 0;; spins++) {
    asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    if (v >= stage_count) break;
    __nanosleep(40);   // the producer shares a sub-partition with two consumer warps: do not spin in their issue slots
    if ((spins & 4095u) == 4095u) spin_check(t0);
  }
}
__device__ __forceinline__ unsigned int ld_acquire(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add(unsigned int* p, unsigned int v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Shared-memory layout of the fp32 activations.  An F8 lane-chunk needs 16 consecutive floats (4 float4), so lanes
// stride 64 B and a straight layout is 4-way bank conflicted; float4 (chunk c, quarter q) is stored at
// 4c + (q ^ ((c >> 1) & 3)) instead, which spreads 8 consecutive lanes over all 8 sixteen-byte bank groups.
template <int Q>
__device__ __forceinline__ int xswz(int f) {   // f = logical float4 index
  if constexpr (Q == Q_F8) return (f & ~3) | ((f & 3) ^ ((f >> 3) & 3));
  else return f;
}

// ---- activation staging with consumer-only barriers -------------------------------------------------------------
__device__ __forceinline__ float c_rms_scale(const float* __restrict__ in, int n, float eps, float* red) {
  float ss = 0.f;
  for (int i = threadIdx.x; i < n; i += kConsumers) { const float v = in[i]; ss = fmaf(v, v, ss); }
  ss = csum(ss, red);
  return 1.0f / sqrtf(ss / (float)n + eps);
}
// one 256-block of quantize_row_q8_K_ref (src/quant.cpp:616-653) by one warp; v[] already scaled/normalised
__device__ __noinline__ void q8_block_nf(float4 va, float4 vb, int b, int8_t* q_qs, float* q_d, short* q_bsums) {
  const int lane = threadIdx.x & 31;
  const float v[8] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w};
  Q8Smem q; q.qs = q_qs; q.d = q_d; q.bsums = q_bsums;
  float amax = 0.f, mx = 0.f;
  int idx = 0x7fffffff;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const float ax = fabsf(v[j]);
    if (ax > amax) { amax = ax; mx = v[j]; idx = lane * 8 + j; }
  }
  {  // block arg-max |v| with REDUX (non-negative floats order like their bit patterns); ties -> lowest index, as the scan above
    const unsigned best = __reduce_max_sync(0xffffffffu, __float_as_uint(amax));
    const int cand = (__float_as_uint(amax) == best) ? idx : 0x7fffffff;
    const int bi = __reduce_min_sync(0xffffffffu, cand);
    mx = __shfl_sync(0xffffffffu, mx, (bi == 0x7fffffff ? 0 : bi) >> 3);
    amax = __uint_as_float(best);
  }
  int qv[8];
  if (amax == 0.f) {
#pragma unroll
    for (int j = 0; j < 8; j++) qv[j] = 0;
    if (lane == 0) q.d[b] = 0.f;
  } else {
    const float iscale = __fdiv_rn(-127.f, mx);
#pragma unroll
    for (int j = 0; j < 8; j++)   // round-to-nearest-even through the fp32 adder (exact for |x| < 2^22): same value as cvt.rni, no XU pipe
      qv[j] = min(127, __float_as_int(__fadd_rn(__fmul_rn(iscale, v[j]), 12582912.f)) - 0x4B400000);
    if (lane == 0) q.d[b] = __fmul_rn(mx, -1.0f / 127.0f);
  }
  int s = qv[0] + qv[1] + qv[2] + qv[3] + qv[4] + qv[5] + qv[6] + qv[7];
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  if ((lane & 1) == 0) q.bsums[b * 16 + (lane >> 1)] = (short)s;
  const uint32_t p0 = (qv[0] & 0xff) | ((qv[1] & 0xff) << 8) | ((qv[2] & 0xff) << 16) | ((uint32_t)(qv[3] & 0xff) << 24);
  const uint32_t p1 = (qv[4] & 0xff) | ((qv[5] & 0xff) << 8) | ((qv[6] & 0xff) << 16) | ((uint32_t)(qv[7] & 0xff) << 24);
  *reinterpret_cast<uint2*>(q.qs + (b << 8) + lane * 8) = make_uint2(p0, p1);
}
// out-of-line on purpose: one copy for every staging site (instruction-cache footprint)
__device__ __forceinline__ void q8_block(const float (&v)[8], int b, int lane, const Q8Smem& q) {
  (void)lane;
  q8_block_nf(make_float4(v[0], v[1], v[2], v[3]), make_float4(v[4], v[5], v[6], v[7]), b, q.qs, q.d, q.bsums);
}
template <int Q>
__device__ __forceinline__ void c_stage_vec(const float* __restrict__ in, int n, const float* __restrict__ norm_w, float sc,
                                            float* xs, const Q8Smem& q8) {
  constexpr bool KQ = QTraits<Q>::kq;
  if constexpr (KQ) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int b = warp; b < (n >> 8); b += 8) {
      const int base = (b << 8) + lane * 8;
      const float4 a0 = *reinterpret_cast<const float4*>(in + base);
      const float4 a1 = *reinterpret_cast<const float4*>(in + base + 4);
      float v[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      if (norm_w) {
        const float4 w0 = *reinterpret_cast<const float4*>(norm_w + base);
        const float4 w1 = *reinterpret_cast<const float4*>(norm_w + base + 4);
        const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = __fmul_rn(__fmul_rn(v[j], sc), ww[j]);
      }
      q8_block(v, b, lane, q8);
    }
  } else {
    for (int f = threadIdx.x; f < (n >> 2); f += kConsumers) {
      float4 o = reinterpret_cast<const float4*>(in)[f];
      if (norm_w) {
        const float4 w = reinterpret_cast<const float4*>(norm_w)[f];
        o.x = __fmul_rn(__fmul_rn(o.x, sc), w.x); o.y = __fmul_rn(__fmul_rn(o.y, sc), w.y);
        o.z = __fmul_rn(__fmul_rn(o.z, sc), w.z); o.w = __fmul_rn(__fmul_rn(o.w, sc), w.w);
      }
      reinterpret_cast<float4*>(xs)[xswz<Q>(f)] = o;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Tensor-core path for F8E5M2 weights.  Batch-1 GEMV is bandwidth-bound on paper, but dequantising every weight on the
// CUDA cores costs ~2.7 instructions per byte and one resident CTA per SM cannot issue that fast.  Instead:
//   * f8e5m2 -> f16 is exact (the byte is the top half of an fp16): one PRMT per two weights builds the A fragment;
//   * the fp32 activation x is split EXACTLY into two fp16 values (hi = rn(x*2^e), lo = rn(x*2^e - hi)), with a
//     power-of-two normalisation 2^e per 64-column group so the split keeps ~22 significant bits for any magnitude;
//   * mma.sync.m16n8k16 (f16 x f16 -> f32): A = 16 weight rows x 16 columns, B columns 0/1 = hi/lo -> every product is
//     exact in fp32 and the accumulation is fp32; the 128x128 block scale and 2^-e are applied per 64-column group.
// The K index inside one mma is permuted (thread-in-group t owns 4 consecutive columns) identically for A and B.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_f16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
struct X16 { uint32_t hi, lo, gs; };   // shared addresses: hi halves [n], lo halves [n], per-64-column 2^-e floats [n/64]
__host__ __device__ inline size_t x16_bytes(int n) { return align_up((size_t)n * 4 + (size_t)(n / 64) * 4, 128); }
__device__ __forceinline__ X16 carve_x16(unsigned char* p, int n) {
  X16 x; x.hi = smem_u32(p); x.lo = x.hi + (uint32_t)n * 2u; x.gs = x.hi + (uint32_t)n * 4u; return x;
}
// one float4 (4 consecutive columns) -> hi/lo halves; 16 lanes (= 64 columns) share the normalisation
__device__ __forceinline__ void x16_store(const X16& x, int f, float4 v);
__device__ __noinline__ void x16_store_nf(uint32_t xhi, uint32_t xlo, uint32_t xgs, int f, float4 v) {
  X16 x; x.hi = xhi; x.lo = xlo; x.gs = xgs;
  float am = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));
  // maximum over the 16-lane group (= 64 columns) with one REDUX; groups are always fully in or out (n % 64 == 0)
  am = __uint_as_float(__reduce_max_sync((threadIdx.x & 16) ? 0xffff0000u : 0x0000ffffu, __float_as_uint(am)));
  const unsigned eb = (__float_as_uint(am) >> 23) & 0xffu;                 // biased exponent of the group maximum
  const unsigned sb = am > 0.f ? min(max(267u - eb, 1u), 254u) : 127u;     // 2^(13 - floor(log2 amax))
  const float sc = __uint_as_float(sb << 23), inv = __uint_as_float((254u - sb) << 23);
  const float a0 = v.x * sc, a1 = v.y * sc, a2 = v.z * sc, a3 = v.w * sc;  // exact (power of two)
  // packed conversions (F2FP.PACK_AB / HADD2.F32): same roundings as scalar cvt.rn, a third of the instructions, no XU pipe
  const __half2 h01 = __floats2half2_rn(a0, a1), h23 = __floats2half2_rn(a2, a3);
  const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
  const __half2 l01 = __floats2half2_rn(a0 - f01.x, a1 - f01.y), l23 = __floats2half2_rn(a2 - f23.x, a3 - f23.y);
  const uint32_t hw0 = *reinterpret_cast<const uint32_t*>(&h01), hw1 = *reinterpret_cast<const uint32_t*>(&h23);
  const uint32_t lw0 = *reinterpret_cast<const uint32_t*>(&l01), lw1 = *reinterpret_cast<const uint32_t*>(&l23);
  asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(x.hi + (uint32_t)f * 8u), "r"(hw0), "r"(hw1) : "memory");
  asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(x.lo + (uint32_t)f * 8u), "r"(lw0), "r"(lw1) : "memory");
  if ((f & 15) == 0) asm volatile("st.shared.f32 [%0], %1;" ::"r"(x.gs + (uint32_t)(f >> 4) * 4u), "f"(inv) : "memory");
}
__device__ __forceinline__ void x16_store(const X16& x, int f, float4 v) { x16_store_nf(x.hi, x.lo, x.gs, f, v); }
// rows (gid) and (gid+8) of a 16-row group over columns [col0, col1) (multiples of 64).  a_lo/a_hi: shared addresses of the
// two weight rows (a_hi == 0: no upper rows); s_lo/s_hi: f8 scale rows (0 = none; block width 2^sshift columns).
// ONE out-of-line copy shared by every stage kind (instruction-cache footprint), written as a single basic block per
// 64-column group so ptxas can overlap the shared-memory loads, PRMTs and HMMAs of four unrolled groups: no predicated
// loads (idle B lanes read a zero block with stride 0, missing rows/scales alias valid memory), no division.
// Returns {rows gid, rows gid+8}, valid in lanes with (lane & 3) == 0.
constexpr uint32_t kHdrZero = 576;   // 32 zero bytes in the CTA header
constexpr uint32_t kHdrOne = 608;    // 1.0f
__device__ __forceinline__ void mma_f16_zero(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
               : "=f"(c[0]), "=f"(c[1]), "=f"(c[2]), "=f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "f"(0.f));
}
// A fragments of one mma from word k of the two 16-byte row chunks: bytes (0,1) -> k-pair 0, bytes (2,3) -> k-pair 1
#define DSK_A4(wl, wh) __byte_perm(wl, 0, 0x1404), __byte_perm(wh, 0, 0x1404), __byte_perm(wl, 0, 0x3424), __byte_perm(wh, 0, 0x3424)
__device__ __noinline__ float2 mma_rows_f8(uint32_t a_lo, uint32_t a_hi, uint32_t s_lo, uint32_t s_hi, int sshift, int col0, int col1,
                                           uint32_t xhi, uint32_t xlo, uint32_t xgs) {
  extern __shared__ __align__(128) unsigned char dsk_dyn_smem[];
  const uint32_t hdr = smem_u32(dsk_dyn_smem);
  const int lane = threadIdx.x & 31, gid = lane >> 2, tig = lane & 3;
  uint32_t pa_lo = a_lo + (uint32_t)(col0 + 16 * tig);
  uint32_t pa_hi = (a_hi ? a_hi : a_lo) + (uint32_t)(col0 + 16 * tig);
  // B columns 0/1 = hi/lo halves of x: only lanes gid 0 and 1 load (predicated, no divergence), columns 2..7 are zero
  const uint32_t bp = gid < 2 ? 1u : 0u;
  uint32_t pb = (gid == 1 ? xlo : xhi) + (uint32_t)(col0 + 16 * tig) * 2u;
  const uint32_t pb_step = 128u;
  uint32_t pg = xgs + (uint32_t)(col0 >> 6) * 4u;
  const uint32_t smask = s_lo ? 0xffffffffu : 0u;
  const uint32_t ps_lo = s_lo ? s_lo : hdr + kHdrOne, ps_hi = s_hi ? s_hi : hdr + kHdrOne;
  float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
  int cb = col0;
  // Two 64-column groups (128 columns) per iteration as four 2-deep HMMA chains; the operands of iteration i+1 are loaded
  // while iteration i computes (register double buffering: ptxas does not move loads across the loop edge by itself, and a
  // single warp per sub-partition is otherwise exposed to the full shared-memory latency every iteration).
  struct Frag { uint4 wa0, wa1, ba0, ba1, wb0, wb1, bb0, bb1; float ga, gb, sla, sha, slb, shb; };
  auto load_frag = [&](Frag& F, int c) {
    F.wa0 = lds128(pa_lo); F.wa1 = lds128(pa_hi); F.ba0 = lds128_pred(pb, bp); F.ba1 = lds128_pred(pb + 16u, bp);
    F.wb0 = lds128(pa_lo + 64u); F.wb1 = lds128(pa_hi + 64u); F.bb0 = lds128_pred(pb + pb_step, bp); F.bb1 = lds128_pred(pb + pb_step + 16u, bp);
    F.ga = __uint_as_float(lds32(pg)); F.gb = __uint_as_float(lds32(pg + 4u));
    const uint32_t sia = (((uint32_t)c >> sshift) << 2) & smask, sib = (((uint32_t)(c + 64) >> sshift) << 2) & smask;
    F.sla = __uint_as_float(lds32(ps_lo + sia)); F.sha = __uint_as_float(lds32(ps_hi + sia));
    F.slb = __uint_as_float(lds32(ps_lo + sib)); F.shb = __uint_as_float(lds32(ps_hi + sib));
    pa_lo += 128u; pa_hi += 128u; pb += 2u * pb_step; pg += 8u;
  };
  auto compute_frag = [&](const Frag& F) {
    float ca0[4], ca1[4], cb0[4], cb1[4];
    mma_f16_zero(ca0, DSK_A4(F.wa0.x, F.wa1.x), F.ba0.x, F.ba0.y);
    mma_f16_zero(ca1, DSK_A4(F.wa0.z, F.wa1.z), F.ba1.x, F.ba1.y);
    mma_f16_zero(cb0, DSK_A4(F.wb0.x, F.wb1.x), F.bb0.x, F.bb0.y);
    mma_f16_zero(cb1, DSK_A4(F.wb0.z, F.wb1.z), F.bb1.x, F.bb1.y);
    mma_f16(ca0, DSK_A4(F.wa0.y, F.wa1.y), F.ba0.z, F.ba0.w);
    mma_f16(ca1, DSK_A4(F.wa0.w, F.wa1.w), F.ba1.z, F.ba1.w);
    mma_f16(cb0, DSK_A4(F.wb0.y, F.wb1.y), F.bb0.z, F.bb0.w);
    mma_f16(cb1, DSK_A4(F.wb0.w, F.wb1.w), F.bb1.z, F.bb1.w);
    const float fla = F.ga * F.sla, fha = F.ga * F.sha, flb = F.gb * F.slb, fhb = F.gb * F.shb;
    t0 = fmaf(ca0[0] + ca1[0], fla, t0);
    t1 = fmaf(ca0[1] + ca1[1], fla, t1);
    t2 = fmaf(ca0[2] + ca1[2], fha, t2);
    t3 = fmaf(ca0[3] + ca1[3], fha, t3);
    t0 = fmaf(cb0[0] + cb1[0], flb, t0);
    t1 = fmaf(cb0[1] + cb1[1], flb, t1);
    t2 = fmaf(cb0[2] + cb1[2], fhb, t2);
    t3 = fmaf(cb0[3] + cb1[3], fhb, t3);
  };
  if (cb + 128 <= col1) {
    Frag cur, nxt;
    load_frag(cur, cb);
    for (; cb + 256 <= col1; cb += 128) {
      load_frag(nxt, cb + 128);
      compute_frag(cur);
      cur = nxt;
    }
    compute_frag(cur);
    cb += 128;
  }
  if (cb < col1) {   // odd group count: one more 64-column group
    const uint4 wa0 = lds128(pa_lo), wa1 = lds128(pa_hi), ba0 = lds128_pred(pb, bp), ba1 = lds128_pred(pb + 16u, bp);
    const float ga = __uint_as_float(lds32(pg));
    const uint32_t sia = (((uint32_t)cb >> sshift) << 2) & smask;
    const float sla = __uint_as_float(lds32(ps_lo + sia)), sha = __uint_as_float(lds32(ps_hi + sia));
    float ca0[4], ca1[4];
    mma_f16_zero(ca0, DSK_A4(wa0.x, wa1.x), ba0.x, ba0.y);
    mma_f16_zero(ca1, DSK_A4(wa0.z, wa1.z), ba1.x, ba1.y);
    mma_f16(ca0, DSK_A4(wa0.y, wa1.y), ba0.z, ba0.w);
    mma_f16(ca1, DSK_A4(wa0.w, wa1.w), ba1.z, ba1.w);
    const float fla = ga * sla, fha = ga * sha;
    t0 = fmaf(ca0[0] + ca1[0], fla, t0);
    t1 = fmaf(ca0[1] + ca1[1], fla, t1);
    t2 = fmaf(ca0[2] + ca1[2], fha, t2);
    t3 = fmaf(ca0[3] + ca1[3], fha, t3);
  }
  return make_float2(t0 + t1, t2 + t3);   // hi-part + lo-part contributions
}

// ---- dots over a column piece, weights and f8 scale row both in shared memory -----------------------------------
// 16 weight bytes x EPC activations with TWO interleaved partial sums (halves the dependent FMA chain)
template <int Q>
__device__ __forceinline__ float chunk_dot2(const uint4& wv, const float* xv) {
  float p0 = 0.f, p1 = 0.f;
  if constexpr (Q == Q_F8) {
    float2 a, b;
    a = f8x2_lo(wv.x); b = f8x2_hi(wv.x); p0 = fmaf(a.x, xv[0], p0); p1 = fmaf(a.y, xv[1], p1); p0 = fmaf(b.x, xv[2], p0); p1 = fmaf(b.y, xv[3], p1);
    a = f8x2_lo(wv.y); b = f8x2_hi(wv.y); p0 = fmaf(a.x, xv[4], p0); p1 = fmaf(a.y, xv[5], p1); p0 = fmaf(b.x, xv[6], p0); p1 = fmaf(b.y, xv[7], p1);
    a = f8x2_lo(wv.z); b = f8x2_hi(wv.z); p0 = fmaf(a.x, xv[8], p0); p1 = fmaf(a.y, xv[9], p1); p0 = fmaf(b.x, xv[10], p0); p1 = fmaf(b.y, xv[11], p1);
    a = f8x2_lo(wv.w); b = f8x2_hi(wv.w); p0 = fmaf(a.x, xv[12], p0); p1 = fmaf(a.y, xv[13], p1); p0 = fmaf(b.x, xv[14], p0); p1 = fmaf(b.y, xv[15], p1);
    return p0 + p1;
  } else {
    return chunk_dot<Q>(wv, xv);
  }
}

// sshift >= 0: scale block index = (c * EPC) >> sshift (power-of-two block width); < 0: generic division by bs1
template <int Q, int NACC>
__device__ __forceinline__ void piece_dot_dense(const uint32_t (&wa)[NACC], const uint32_t (&ssm)[NACC], int bs1, int sshift, int g0, int g1,
                                                uint32_t xs, int lane, float (&acc)[NACC]) {
  constexpr int EPC = QTraits<Q>::epc;
#pragma unroll 2
  for (int c = g0 + lane; c < g1; c += 32) {
    float xv[EPC];
#pragma unroll
    for (int q = 0; q < EPC / 4; q++) {
      const uint4 t = lds128(xs + (uint32_t)xswz<Q>(c * (EPC / 4) + q) * 16u);
      xv[4 * q] = __uint_as_float(t.x); xv[4 * q + 1] = __uint_as_float(t.y);
      xv[4 * q + 2] = __uint_as_float(t.z); xv[4 * q + 3] = __uint_as_float(t.w);
    }
    const uint32_t sidx = (sshift >= 0 ? (uint32_t)(c * EPC) >> sshift : (uint32_t)((c * EPC) / bs1)) * 4u;
#pragma unroll
    for (int r = 0; r < NACC; r++) {
      const uint4 wv = lds128(wa[r] + (uint32_t)c * 16u);
      const float p = chunk_dot2<Q>(wv, xv);
      const float s = ssm[r] ? __uint_as_float(lds32(ssm[r] + sidx)) : 1.0f;
      acc[r] = fmaf(p, s, acc[r]);
    }
  }
}
// K-quant rows over blocks [b0,b1): same arithmetic as dot_q2k/dot_q3k, block range instead of whole row
template <int Q>
__device__ __forceinline__ float piece_dot_kq(uint32_t wrow, int b0, int b1, const Q8Smem& q8, int lane) {
  float acc = 0.f;
  const int q0 = b0 * 4, q1 = b1 * 4;
  for (int base = q0; base < q1; base += 32) {
    const int qb = base + lane;
    const bool act = qb < q1;
    const int b = qb >> 2, h = (qb >> 1) & 1, c = qb & 1;
    int isum = 0, summs = 0;
    if constexpr (Q == Q_Q2K) {
      const uint32_t blk = wrow + (uint32_t)b * kQ2Bytes;
      if (act) {
        const uint32_t qp = blk + 16 + 32 * h + 16 * c;
        const uint32_t w0 = lds32(qp), w1 = lds32(qp + 4), w2 = lds32(qp + 8), w3 = lds32(qp + 12);
        const uint32_t sA = lds32(blk + 8 * h), sB = lds32(blk + 8 * h + 4);
        const int8_t* y = q8.qs + b * 256 + 128 * h + 16 * c;
        const short* bs = q8.bsums + b * 16 + 8 * h + c;
#pragma unroll
        for (int s = 0; s < 4; s++) {
          const int4 yv = *reinterpret_cast<const int4*>(y + 32 * s);
          int dp = __dp4a((int)((w0 >> (2 * s)) & 0x03030303u), yv.x, 0);
          dp = __dp4a((int)((w1 >> (2 * s)) & 0x03030303u), yv.y, dp);
          dp = __dp4a((int)((w2 >> (2 * s)) & 0x03030303u), yv.z, dp);
          dp = __dp4a((int)((w3 >> (2 * s)) & 0x03030303u), yv.w, dp);
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
        acc += (yd * h2f((uint16_t)(dm & 0xffff))) * (float)isum - (yd * h2f((uint16_t)(dm >> 16))) * (float)summs;
      }
    } else {
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
          const int t = 2 * s + c;
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
  }
  return acc;
}

// NACC rows x one piece -> NACC warp-reduced partial dots
template <int Q, int NACC>
__device__ __forceinline__ void piece_rows(const uint32_t (&wa)[NACC], const uint32_t (&ssm)[NACC], int bs1, int sshift, int g0, int g1,
                                           uint32_t xs, const Q8Smem& q8, int lane, float (&out)[NACC]) {
#pragma unroll
  for (int r = 0; r < NACC; r++) out[r] = 0.f;
  if constexpr (QTraits<Q>::kq) {
#pragma unroll
    for (int r = 0; r < NACC; r++) out[r] = piece_dot_kq<Q>(wa[r], g0, g1, q8, lane);
  } else {
    piece_dot_dense<Q, NACC>(wa, ssm, bs1, sshift, g0, g1, xs, lane, out);
  }
#pragma unroll
  for (int r = 0; r < NACC; r++) out[r] = warp_sum(out[r]);
}

// ---- shared-memory map of the interpreter -----------------------------------------------------------------------
// The TMA ring.  Tiles are numbered per CTA in issue order (`it`, the same sequence on the producer and consumer side, running
// across stages).  Tile j uses mbarrier pair (full, empty)[j mod 16]; its BYTES are carved out of one circular buffer by the
// producer (variable size per tile: small DOWN pieces do not occupy a whole 20 KB slot, so up to 16 of them are in flight and
// keep the HBM pipe full; large tiles simply use more of the buffer), which publishes the offset in ent_off[] before issuing
// the copies.  Consumers CLAIM tiles dynamically (shared-memory counter): a warp that finishes early takes the next tile, so
// the stage ends when the last tile does, not when the most loaded warp of a static assignment does.  A claimer of use k of
// an entry first waits until use k-1 has been consumed (consumed[] counter) — only then does the full-barrier parity k & 1
// refer to the right phase (an mbarrier parity wait cannot tell phase k-2 from phase k).
struct MegaSmem {
  uint32_t full0, empty0, dep, claim;   // shared addresses: mbarrier arrays (8 B apart), producer release word, claim counter
  uint32_t ent_off, consumed, start_abs;   // 16 x u32 each: tile offset in the ring / completed uses per entry / producer bookkeeping
  float* red;          // 32 floats
  int* act; float* actw;   // routing (K <= 16)
  float* res;          // 512 floats of partial results
  float* sx;           // 256 gate scores
  int* sel;            // 16 ints: warp-per-tile DOWN completion counters
  int* marker;         // "routing of layer l is in this CTA's shared memory"
  unsigned char* xregion;
  uint32_t ring;       // shared address of the circular tile buffer
  Stage* st_c;         // consumers' copy of the current stage descriptor
  Stage* st_p;         // producer's copy (it may be one stage ahead)
  Program* prog;       // header copy (no stage array)
};
__device__ __forceinline__ MegaSmem carve_mega(unsigned char* smem, int xregion_bytes) {
  MegaSmem m;
  const uint32_t b = smem_u32(smem);
  m.full0 = b; m.empty0 = b + 128;                         // 16 + 16 mbarriers -> 256
  m.red = reinterpret_cast<float*>(smem + 256);            // 32 floats -> 384
  m.act = reinterpret_cast<int*>(smem + 384);              // 16 ints -> 448
  m.actw = reinterpret_cast<float*>(smem + 448);           // 16 floats -> 512
  m.sel = reinterpret_cast<int*>(smem + 512);              // 16 ints -> 576
  //                                                          576..624: kHdrZero / kHdrOne
  m.dep = b + 624; m.claim = b + 628;
  m.marker = reinterpret_cast<int*>(smem + 632);
  //                                                          640..736: routing scratch (cmask, selv)
  m.sx = reinterpret_cast<float*>(smem + 768);             // 256 floats -> 1792
  m.res = reinterpret_cast<float*>(smem + 1792);           // 512 floats -> 3840
  m.ent_off = b + 3840; m.consumed = b + 3904; m.start_abs = b + 3968;   // -> 4032
  m.st_c = reinterpret_cast<Stage*>(smem + 4096);
  m.st_p = reinterpret_cast<Stage*>(smem + 4096 + kStageSlot);
  m.prog = reinterpret_cast<Program*>(smem + 4096 + 2 * kStageSlot);
  m.xregion = smem + kMegaHdr;
  m.ring = b + kMegaHdr + (uint32_t)xregion_bytes;
  return m;
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

// ---- producer side: reserve `need` bytes (multiple of 128) for tile `it`; returns the tile's shared address -------------
struct RingProd { uint32_t head; int oldest; };   // absolute byte counter (mod 2^32) / oldest tile not yet known released
__device__ __forceinline__ uint32_t ring_acquire(const Program& P, const MegaSmem& sm, RingProd& rp, int it, uint32_t need, int max_inflight,
                                                 uint32_t& full) {
  const uint32_t R = (uint32_t)P.ring_bytes;
  uint32_t pos = rp.head % R;
  if (pos + need > R) { rp.head += R - pos; pos = 0; }    // a tile is contiguous: skip the tail of the buffer
  // wait (in issue order) for old tiles to be released until the mbarrier pair is free, the in-flight cap holds and the
  // bytes [start of the oldest live tile, head) leave room for this one
  while (rp.oldest < it) {
    const bool pair_busy = rp.oldest + max_inflight <= it;
    const uint32_t used = rp.head - lds32(sm.start_abs + 4u * (uint32_t)(rp.oldest & (kRingEntries - 1)));
    if (!pair_busy && used + need <= R) break;
    mbar_wait_backoff(sm.empty0 + 8u * (uint32_t)(rp.oldest & (kRingEntries - 1)), (uint32_t)((rp.oldest / kRingEntries) & 1));
    rp.oldest++;
  }
  const uint32_t e = (uint32_t)(it & (kRingEntries - 1));
  sts32(sm.start_abs + 4u * e, rp.head);
  sts32(sm.ent_off + 4u * e, pos);
  rp.head += need;
  full = sm.full0 + 8u * e;
  return sm.ring + pos;
}
// ---- consumer side -----------------------------------------------------------------------------------------------
// next tile of this CTA (warp-uniform; lane 0 takes the ticket)
__device__ __forceinline__ int ring_claim(const MegaSmem& sm) {
  int j = 0;
  if ((threadIdx.x & 31) == 0) asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(j) : "r"(sm.claim) : "memory");
  return __shfl_sync(0xffffffffu, j, 0);
}
// wait until tile j has landed; returns its shared address
__device__ __forceinline__ uint32_t ring_wait(const MegaSmem& sm, int j) {
  const uint32_t e = (uint32_t)(j & (kRingEntries - 1));
  const uint32_t k = (uint32_t)(j / kRingEntries);
  unsigned long long t0 = 0ull;
  for (unsigned spins = 0;; spins++) {   // use k-1 of this entry fully consumed?  (then the parity below cannot alias)
    uint32_t c;
    asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(c) : "r"(sm.consumed + 4u * e) : "memory");
    if (c >= k) break;
    if ((spins & 4095u) == 4095u) spin_check(t0);
  }
  mbar_wait_guard(sm.full0 + 8u * e, k & 1u);
  return sm.ring + lds32(sm.ent_off + 4u * e);
}
// one thread per tile, after every reader of the tile is done with it
__device__ __forceinline__ void ring_release(const MegaSmem& sm, int j) {
  const uint32_t e = (uint32_t)(j & (kRingEntries - 1));
  asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(sm.consumed + 4u * e), "r"((uint32_t)(j / kRingEntries) + 1u) : "memory");
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sm.empty0 + 8u * e) : "memory");
}
// tiles of a stage that land on this CTA (tile t = blockIdx.x + i * gridDim.x < n)
__device__ __forceinline__ int my_tile_count(int n) { return (int)blockIdx.x < n ? (n - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0; }

// routing: softmax|sigmoid(+bias) and greedy / group-limited top-K (moe_gate, src/infer.cpp:493-599), computed redundantly by
// every CTA from the gate logits with ONE THREAD PER EXPERT (E <= 256 = the consumer threads) and no dependent argmax rounds:
//   * the reference's K x "arg-max with strict >, lowest index wins ties" selection picks, in order, the experts of RANK
//     0..K-1 where rank(j) = #{j' selectable : x[j'] > x[j] or (x[j'] == x[j] and j' < j)} — every thread counts its own rank;
//   * GROUP_LIMITED_GREEDY first keeps, inside each group, the topk_group best experts with x > 0 (the reference compares
//     the first candidate against x[-1], which is 0.0f in its heap layout: SURVEY §8 A7) — the same rank inside the group;
//   * wsum is accumulated in selection order like the reference (only when norm_topk_prob).
// `publish` (CTA 0) also writes the state buffers.  A slot with no selectable expert left (the reference then indexes
// x[-1] / mask[-1]: UB) gets expert -1 and weight 0.
__device__ __noinline__ void route_all(const Program* Pp, const Stage* stp, bool publish) {
  const Program& P = *Pp; const Stage& st = *stp;
  extern __shared__ __align__(128) unsigned char dsk_dyn_smem[];
  float* red = reinterpret_cast<float*>(dsk_dyn_smem + 256);
  int* act_smem = reinterpret_cast<int*>(dsk_dyn_smem + 384);
  float* actw_smem = reinterpret_cast<float*>(dsk_dyn_smem + 448);
  unsigned* cmask = reinterpret_cast<unsigned*>(dsk_dyn_smem + 640);   // 8 words: selectable experts (576..624 hold kHdrZero / kHdrOne)
  float* selv = reinterpret_cast<float*>(dsk_dyn_smem + 672);          // 16 floats: score of selection k
  float* sx = reinterpret_cast<float*>(dsk_dyn_smem + 768);            // 256 scores (MegaSmem::sx)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, E = P.E;
  const bool mine = tid < E;
  const float v = mine ? __ldcg(st.gate_logits + tid) : -3.402823466e38f;
  float s;
  if (P.sigmoid) {
    s = 1.0f / (1.0f + expf(-v));
  } else {
    const float mx = cmax(v, red);
    const float e = mine ? expf(v - mx) : 0.f;
    const float sum = csum(e, red);
    s = e / sum;
  }
  if (st.gate_bias && mine) s += st.gate_bias[tid];
  if (mine) sx[tid] = s;
  if (tid < 16) { act_smem[tid] = -1; selv[tid] = 0.f; }
  csync();
  bool cand = mine;
  if (P.topk_method == 1) {   // keep only the topk_group best positive experts of every group
    const int gs = E / P.n_group;
    int rank = 0x7fffffff;
    if (mine && s > 0.0f) {
      const int g0 = (tid / gs) * gs;
      rank = 0;
#pragma unroll 1
      for (int j = g0; j < g0 + gs; j++) { const float o = sx[j]; rank += (o > s || (o == s && j < tid)) ? 1 : 0; }
    }
    cand = rank < P.topk_group;
  }
  const unsigned bal = __ballot_sync(0xffffffffu, cand);
  if (lane == 0) cmask[warp] = bal;
  csync();
  if (cand) {
    int rank = 0;
#pragma unroll 1
    for (int w = 0; w < 8; w++) {
      unsigned m = cmask[w];
#pragma unroll 1
      while (m) {
        const int j = w * 32 + __ffs(m) - 1;
        m &= m - 1;
        const float o = sx[j];
        rank += (o > s || (o == s && j < tid)) ? 1 : 0;
      }
    }
    if (rank < P.K) { act_smem[rank] = tid; selv[rank] = s; }
  }
  csync();
  if (tid < P.K) {
    float wsum = 1.0f;
    if (P.norm_topk_prob) {
      wsum = 0.f;
#pragma unroll 1
      for (int k = 0; k < P.K; k++) if (act_smem[k] >= 0) wsum += selv[k];
    }
    const int e = act_smem[tid];
    const float w = e >= 0 ? selv[tid] / wsum * P.routed_scale : 0.f;
    actw_smem[tid] = w;
    if (publish) { P.act[tid] = e; P.act_w[tid] = w; }
  }
  if (publish && mine) P.moe_scores[tid] = s;   // state buffer for the host (moe_weights after softmax|sigmoid + bias)
  if (tid == 0) *reinterpret_cast<int*>(dsk_dyn_smem + 632) = st.layer + 1;   // MegaSmem::marker: "routing of layer l is here"
  csync();
}

__device__ __forceinline__ void stage_route_hook(const Program& P, const Stage& st, int route, int stage_index) {
  extern __shared__ __align__(128) unsigned char dsk_dyn_smem[];
  if (route >= 0) {   // uniform over the consumer threads: every one of them takes part (thread per expert)
    const unsigned long long tr0 = (P.tstamp && blockIdx.x == 0) ? gtime() : 0ull;
    route_all(&P, &st, blockIdx.x == 0);
    if (threadIdx.x == 0) dep_signal(smem_u32(dsk_dyn_smem) + 624u, route);
    if (P.tstamp && blockIdx.x == 0 && threadIdx.x == 0) { P.tstamp[stage_index * 8 + 6] = (tr0 - P.tstamp[stage_index * 8]) * 1000ull; P.tstamp[stage_index * 8 + 7] = (gtime() - tr0) * 1000ull; }
  }
}

// ---- producer: one tile -> ring slot --------------------------------------------------------------------------
// f8 scale row by TMA: source aligned down to 16 B (allocations are padded), returns the byte shift inside the slot
__device__ __forceinline__ uint32_t scale_copy_bytes(const float* src, int nfloats, const float*& aligned_src, uint32_t& shift) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(src);
  const uintptr_t a0 = a & ~(uintptr_t)15;
  shift = (uint32_t)(a - a0);
  aligned_src = reinterpret_cast<const float*>(a0);
  return (uint32_t)align_up((size_t)shift + (size_t)nfloats * 4, 16);
}

template <int Q>
__device__ __forceinline__ void produce_tile(const Program& P, const Stage& st, int t, uint32_t slot, uint32_t full,
                                             const int* act_smem) {
  if (st.kind == ST_GEMV) {
    int j = 0;
    while (j + 1 < st.njobs && t >= st.job[j + 1].tile_begin) j++;
    const MJob& jb = st.job[j];
    const int r0 = (t - jb.tile_begin) * st.rows_per_tile;
    const int nrows = min(st.rows_per_tile, jb.rows - r0);
    const size_t rb = QTraits<Q>::row_bytes(st.n);
    const uint32_t bytes = (uint32_t)align_up((size_t)nrows * rb, 16);
    const uint32_t part_stride = (uint32_t)align_up((size_t)st.rows_per_tile * rb, 128);
    const int parts = st.epi == EPI_GLU ? 2 : 1;
    size_t woff = (size_t)r0 * rb, soff = 0;
    if (jb.expert_slot >= 0) {
      const int e = act_smem[jb.expert_slot] - P.expert_first;
      if (e < 0 || e >= P.expert_count) { mbar_expect_tx(full, 0); return; }   // not on this rank: empty tile
      woff += (size_t)e * jb.w_stride;
      soff = (size_t)e * jb.s_stride;
    }
    const int ncb = (st.n + P.bs1 - 1) / P.bs1;
    uint32_t total = bytes * parts, sbytes[2] = {0, 0}, shift;
    const float* ssrc[2] = {nullptr, nullptr};
    if (jb.scale) {
      sbytes[0] = scale_copy_bytes(jb.scale + soff + (size_t)(r0 / P.bs0) * ncb, ncb, ssrc[0], shift);
      if (parts == 2) sbytes[1] = scale_copy_bytes(jb.scale_b + soff + (size_t)(r0 / P.bs0) * ncb, ncb, ssrc[1], shift);
      total += sbytes[0] + sbytes[1];
    }
    mbar_expect_tx(full, total);
    bulk_g2s(slot + (uint32_t)P.slot_scale, jb.w + woff, bytes, full);
    if (parts == 2) bulk_g2s(slot + (uint32_t)P.slot_scale + part_stride, jb.w_b + woff, bytes, full);
    if (jb.scale) {
      bulk_g2s(slot, ssrc[0], sbytes[0], full);
      if (parts == 2) bulk_g2s(slot + (uint32_t)P.slot_scale / 2, ssrc[1], sbytes[1], full);
    }
  } else {  // ST_DOWN: K routed segments + the shared/dense segment of rows [i0, i0+nrows)
    const int i0 = t * st.rows_per_tile;
    const int nrows = min(st.rows_per_tile, P.dim - i0);
    const size_t rb_mi = QTraits<Q>::row_bytes(st.mi), rb_sh = QTraits<Q>::row_bytes(st.sh);
    const uint32_t b_mi = (uint32_t)align_up((size_t)nrows * rb_mi, 16), b_sh = (uint32_t)align_up((size_t)nrows * rb_sh, 16);
    const int ncb_mi = (st.mi + P.bs1 - 1) / P.bs1, ncb_sh = (st.sh + P.bs1 - 1) / P.bs1;
    const bool use_shared = st.sw2 != nullptr && st.add_shared;
    uint32_t total = 0, shift;
    const float* ssrc[kMaxJobs];
    uint32_t sbytes[kMaxJobs];
    int eidx[kMaxJobs];
    for (int k = 0; k < st.K; k++) {
      const int e = act_smem[k] - P.expert_first;
      eidx[k] = (e >= 0 && e < P.expert_count) ? e : -1;
      sbytes[k] = 0;
      if (eidx[k] < 0) continue;
      total += b_mi;
      if (st.s2) { sbytes[k] = scale_copy_bytes(st.s2 + (size_t)e * st.s2_stride + (size_t)(i0 / P.bs0) * ncb_mi, ncb_mi, ssrc[k], shift); total += sbytes[k]; }
    }
    sbytes[st.K] = 0;
    if (use_shared) {
      total += b_sh;
      if (st.ss2) { sbytes[st.K] = scale_copy_bytes(st.ss2 + (size_t)(i0 / P.bs0) * ncb_sh, ncb_sh, ssrc[st.K], shift); total += sbytes[st.K]; }
    }
    mbar_expect_tx(full, total);
    const uint32_t sstride = (uint32_t)P.slot_scale / (uint32_t)(st.K + 1) & ~15u;
    for (int k = 0; k < st.K; k++) {
      if (eidx[k] < 0) continue;
      bulk_g2s(slot + (uint32_t)P.slot_scale + (uint32_t)k * st.seg_stride, st.w2 + (size_t)eidx[k] * st.w2_stride + (size_t)i0 * rb_mi, b_mi, full);
      if (sbytes[k]) bulk_g2s(slot + (uint32_t)k * sstride, ssrc[k], sbytes[k], full);
    }
    if (use_shared) {
      bulk_g2s(slot + (uint32_t)P.slot_scale + (uint32_t)st.K * st.seg_stride, st.sw2 + (size_t)i0 * rb_sh, b_sh, full);
      if (sbytes[st.K]) bulk_g2s(slot + (uint32_t)st.K * sstride, ssrc[st.K], sbytes[st.K], full);
    }
  }
}

// ---- consumers: one GEMV tile ---------------------------------------------------------------------------------
template <int Q, int R, bool GLU>
__device__ __forceinline__ void gemv_tile_tasks(const Program& P, const Stage& st, const MJob& jb, int r0, int nrows, uint32_t slot,
                                                uint32_t xs, const Q8Smem& q8, float* res, int soff_floats) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t rb = (uint32_t)QTraits<Q>::row_bytes(st.n);
  const uint32_t part_stride = (uint32_t)align_up((size_t)st.rows_per_tile * rb, 128);
  const int csplit = st.npieces;                        // pieces of ST_GEMV are column splits of the single input
  const int ngroups = (nrows + R - 1) / R;
  constexpr int NACC = R * (GLU ? 2 : 1);
  // scale rows: slot+0 (part 0) and slot+kSlotScale/2 (part 1), shifted by the source misalignment
  uint32_t s0 = 0, s1 = 0;
  if (jb.scale) {
    s0 = slot + (uint32_t)soff_floats;
    s1 = slot + (uint32_t)P.slot_scale / 2 + (uint32_t)soff_floats;
  }
  for (int task = warp; task < ngroups * csplit; task += 8) {
    const int g = task / csplit, pc = task - g * csplit;
    const Piece pcd = st.piece[pc];
    uint32_t wa[NACC], ssm[NACC];
#pragma unroll
    for (int i = 0; i < R; i++) {
      const int lr = min(g * R + i, nrows - 1);
      wa[i] = slot + (uint32_t)P.slot_scale + (uint32_t)lr * rb;
      ssm[i] = s0;
      if constexpr (GLU) { wa[R + i] = wa[i] + part_stride; ssm[R + i] = s1; }
    }
    float v[NACC];
    piece_rows<Q, NACC>(wa, ssm, P.bs1, P.bs1_shift, pcd.g0, pcd.g1, xs, q8, lane, v);
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < R; i++) {
        const int lr = g * R + i;
        if (lr < nrows) {
          res[(lr * 2 + 0) * csplit + pc] = v[i];
          if constexpr (GLU) res[(lr * 2 + 1) * csplit + pc] = v[R + i];
        }
      }
    }
  }
}

// F8 tile through the tensor cores: one 16-row group (8 rows x {gate, up} for GLU) per task, K split over the pieces
template <bool GLU>
__device__ __forceinline__ void gemv_tile_tasks_mma(const Program& P, const Stage& st, const MJob& jb, int nrows, uint32_t slot,
                                                    const X16& x16, float* res, int shift_bytes) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, gid = lane >> 2;
  const uint32_t rb = (uint32_t)f8_pitch((size_t)st.n);   // f8 device row pitch
  const uint32_t part_stride = (uint32_t)align_up((size_t)st.rows_per_tile * rb, 128);
  const int csplit = st.npieces;
  constexpr int RPG = GLU ? 8 : 16;     // tile rows per mma row group
  const int ngroups = (nrows + RPG - 1) / RPG;
  uint32_t s0 = 0, s1 = 0;
  if (jb.scale) { s0 = slot + (uint32_t)shift_bytes; s1 = slot + (uint32_t)P.slot_scale / 2 + (uint32_t)shift_bytes; }
  for (int task = warp; task < ngroups * csplit; task += 8) {
    const int g = task / csplit, pc = task - g * csplit;
    const Piece pcd = st.piece[pc];
    int r_lo, r_hi;
    uint32_t a_lo, a_hi;
    if constexpr (GLU) {
      r_lo = g * 8 + gid; r_hi = r_lo;
      const int lr = min(r_lo, nrows - 1);
      a_lo = slot + (uint32_t)P.slot_scale + (uint32_t)lr * rb;
      a_hi = a_lo + part_stride;
    } else {
      r_lo = g * 16 + gid; r_hi = r_lo + 8;
      a_lo = slot + (uint32_t)P.slot_scale + (uint32_t)min(r_lo, nrows - 1) * rb;
      a_hi = slot + (uint32_t)P.slot_scale + (uint32_t)min(r_hi, nrows - 1) * rb;
    }
    float v_lo, v_hi;
    { const float2 vv = mma_rows_f8(a_lo, a_hi, s0, GLU ? s1 : s0, P.bs1_shift, pcd.g0 * 64, pcd.g1 * 64, x16.hi, x16.lo, x16.gs); v_lo = vv.x; v_hi = vv.y; }
    if ((lane & 3) == 0) {
      if constexpr (GLU) {
        if (r_lo < nrows) { res[(r_lo * 2 + 0) * csplit + pc] = v_lo; res[(r_lo * 2 + 1) * csplit + pc] = v_hi; }
      } else {
        if (r_lo < nrows) res[(r_lo * 2 + 0) * csplit + pc] = v_lo;
        if (r_hi < nrows) res[(r_hi * 2 + 0) * csplit + pc] = v_hi;
      }
    }
  }
}

template <int Q>
__device__ __forceinline__ void consume_gemv_tile(const Program& P, const Stage& st, int t, uint32_t slot, uint32_t xs,
                                                  const Q8Smem& q8, const X16& x16, float* res, const int* act_smem,
                                                  unsigned long long& best, bool& skip) {
  int j = 0;
  while (j + 1 < st.njobs && t >= st.job[j + 1].tile_begin) j++;
  const MJob& jb = st.job[j];
  const int r0 = (t - jb.tile_begin) * st.rows_per_tile;
  const int nrows = min(st.rows_per_tile, jb.rows - r0);
  skip = false;
  size_t soff = 0;
  if (jb.expert_slot >= 0) {
    const int e = act_smem[jb.expert_slot] - P.expert_first;
    if (e < 0 || e >= P.expert_count) { skip = true; return; }
    soff = (size_t)e * jb.s_stride;
  }
  int shift_bytes = 0;
  if (jb.scale) {
    const int ncb = (st.n + P.bs1 - 1) / P.bs1;
    shift_bytes = (int)(reinterpret_cast<uintptr_t>(jb.scale + soff + (size_t)(r0 / P.bs0) * ncb) & 15);
  }
  const bool glu = st.epi == EPI_GLU;
  if (Q == Q_F8 && st.use_mma) {
    if (glu) gemv_tile_tasks_mma<true>(P, st, jb, nrows, slot, x16, res, shift_bytes);
    else gemv_tile_tasks_mma<false>(P, st, jb, nrows, slot, x16, res, shift_bytes);
    (void)best;
    return;
  }
  if (glu) {
    if (st.rpass >= 2) gemv_tile_tasks<Q, 2, true>(P, st, jb, r0, nrows, slot, xs, q8, res, shift_bytes);
    else gemv_tile_tasks<Q, 1, true>(P, st, jb, r0, nrows, slot, xs, q8, res, shift_bytes);
  } else {
    if (st.rpass >= 4) gemv_tile_tasks<Q, 4, false>(P, st, jb, r0, nrows, slot, xs, q8, res, shift_bytes);
    else if (st.rpass >= 2) gemv_tile_tasks<Q, 2, false>(P, st, jb, r0, nrows, slot, xs, q8, res, shift_bytes);
    else gemv_tile_tasks<Q, 1, false>(P, st, jb, r0, nrows, slot, xs, q8, res, shift_bytes);
  }
  (void)best;
}

// ---- multi-GPU partial sums -----------------------------------------------------------------------------------------
// Epilogues write tensor-/expert-parallel partial sums into the LOCAL partial vector; the exchange stage that follows pushes
// it to the peers in coalesced 1 KB chunks (scattered 32-byte NVLink writes straight from the epilogues cost ~8 us per stage).
__device__ __forceinline__ void store_partial(const Program& P, const Stage& st, int i, float acc) {
  (void)st;
  P.partial[i] = acc;
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// LM-head rows are sharded under tensor parallelism: every rank stores its logits into every rank's full-vocabulary buffer
// (NVLink posted writes), made visible by the ST_AMAX exchange that follows the stage
__device__ __forceinline__ void store_logit(const Program& P, const MJob& jb, int r, float val) {
  jb.out[r] = val;
  if (P.tp && P.n_ranks > 1) {
    const size_t row = (size_t)jb.row_base + (size_t)r;
    for (int q = 0; q < P.n_ranks; q++) if (q != P.rank) P.logits_peer[q][row] = val;
  }
}
// combine column pieces in a fixed order + epilogue, one thread per row of the tile
__device__ __forceinline__ void gemv_tile_epilogue(const Program& P, const Stage& st, int t, const float* res, unsigned long long& best,
                                                   float xres) {
  int j = 0;
  while (j + 1 < st.njobs && t >= st.job[j + 1].tile_begin) j++;
  const MJob& jb = st.job[j];
  const int r0 = (t - jb.tile_begin) * st.rows_per_tile;
  const int nrows = min(st.rows_per_tile, jb.rows - r0);
  const int lr = threadIdx.x;
  if (lr >= nrows) return;
  const int csplit = st.npieces;
  float v = 0.f, u = 0.f;
  for (int pc = 0; pc < csplit; pc++) v += res[(lr * 2 + 0) * csplit + pc];
  const int r = r0 + lr;
  float val = v;
  if (st.epi == EPI_GLU) {
    for (int pc = 0; pc < csplit; pc++) u += res[(lr * 2 + 1) * csplit + pc];
    val = (P.act_silu ? silu_f(v) : gelu_f(v)) * u;
  }
  switch (st.epi) {
    case EPI_RESID: jb.out[r] = xres + val; break;
    case EPI_PARTIAL: store_partial(P, st, r, val); break;
    case EPI_KVB: {
      jb.out[r] = val;
      const int per = P.nope + P.vh, hh = r / per, ii = r - hh * per;
      const int kv_pos = P.ctrl->kv_pos;
      if (ii < P.nope) st.kcache[(size_t)kv_pos * P.n_heads * P.hd + hh * P.hd + ii] = __float2half_rn(val);
      else st.vcache[(size_t)kv_pos * P.n_heads * P.vh + hh * P.vh + (ii - P.nope)] = __float2half_rn(val);
      break;
    }
    case EPI_LOGITS: {
      store_logit(P, jb, r, val);
      const unsigned long long key = ((unsigned long long)orderable(val) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)(jb.row_base + r));
      if (key > best) best = key;
      break;
    }
    default: jb.out[r] = val; break;
  }
}

// ---- consumers: one DOWN tile (pieces = column pieces of the K routed segments + the shared segment) -----------------
// ST_DOWN tile through the tensor cores: task = one column piece of one segment, all (<= 8) rows of the tile at once
__device__ __forceinline__ void consume_down_tile_mma(const Program& P, const Stage& st, int t, uint32_t slot, const X16* x16_seg,
                                                      float* res, const int* act_smem) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, gid = lane >> 2;
  const int i0 = t * st.rows_per_tile;
  const int nrows = min(st.rows_per_tile, P.dim - i0);
  const uint32_t rb_mi = (uint32_t)f8_pitch((size_t)st.mi), rb_sh = (uint32_t)f8_pitch((size_t)st.sh);
  const int ncb_mi = (st.mi + P.bs1 - 1) / P.bs1, ncb_sh = (st.sh + P.bs1 - 1) / P.bs1;
  const uint32_t sstride = (uint32_t)P.slot_scale / (uint32_t)(st.K + 1) & ~15u;
  const int np = st.npieces;
  for (int pc = warp; pc < np; pc += 8) {
    const Piece pcd = st.piece[pc];
    const int k = pcd.seg;
    bool live = true;
    uint32_t ssm = 0;
    if (k < st.K) {
      const int e = act_smem[k] - P.expert_first;
      live = e >= 0 && e < P.expert_count;
      if (live && st.s2)
        ssm = slot + (uint32_t)k * sstride + (uint32_t)(reinterpret_cast<uintptr_t>(st.s2 + (size_t)e * st.s2_stride + (size_t)(i0 / P.bs0) * ncb_mi) & 15);
    } else {
      live = st.sw2 != nullptr && st.add_shared;
      if (live && st.ss2)
        ssm = slot + (uint32_t)k * sstride + (uint32_t)(reinterpret_cast<uintptr_t>(st.ss2 + (size_t)(i0 / P.bs0) * ncb_sh) & 15);
    }
    float v_lo = 0.f, v_hi = 0.f;
    if (live) {
      const uint32_t rb = k < st.K ? rb_mi : rb_sh;
      const uint32_t base = slot + (uint32_t)P.slot_scale + (uint32_t)k * st.seg_stride;
      const uint32_t a_lo = base + (uint32_t)min(gid, nrows - 1) * rb, a_hi = base + (uint32_t)min(gid + 8, nrows - 1) * rb;
      { const X16 xk = x16_seg[k]; const float2 vv = mma_rows_f8(a_lo, a_hi, ssm, ssm, P.bs1_shift, pcd.g0 * 64, pcd.g1 * 64, xk.hi, xk.lo, xk.gs); v_lo = vv.x; v_hi = vv.y; }
    }
    if ((lane & 3) == 0) {
      if (gid < nrows) res[gid * np + pc] = v_lo;
      if (gid + 8 < nrows) res[(gid + 8) * np + pc] = v_hi;
    }
  }
}

template <int Q>
__device__ __forceinline__ void consume_down_tile(const Program& P, const Stage& st, int t, uint32_t slot, const uint32_t* xs_seg,
                                                  const Q8Smem* q8_seg, float* res, const int* act_smem) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i0 = t * st.rows_per_tile;
  const int nrows = min(st.rows_per_tile, P.dim - i0);
  const uint32_t rb_mi = (uint32_t)QTraits<Q>::row_bytes(st.mi), rb_sh = (uint32_t)QTraits<Q>::row_bytes(st.sh);
  const int ncb_mi = (st.mi + P.bs1 - 1) / P.bs1, ncb_sh = (st.sh + P.bs1 - 1) / P.bs1;
  const uint32_t sstride = (uint32_t)P.slot_scale / (uint32_t)(st.K + 1) & ~15u;
  const int np = st.npieces;
  for (int task = warp; task < nrows * np; task += 8) {
    const int lr = task / np, pc = task - lr * np;
    const Piece pcd = st.piece[pc];
    const int k = pcd.seg;
    float v[1] = {0.f};
    bool live = true;
    uint32_t ssm[1] = {0};
    if (k < st.K) {
      const int e = act_smem[k] - P.expert_first;
      live = e >= 0 && e < P.expert_count;
      if (live && st.s2)
        ssm[0] = slot + (uint32_t)k * sstride + (uint32_t)(reinterpret_cast<uintptr_t>(st.s2 + (size_t)e * st.s2_stride + (size_t)(i0 / P.bs0) * ncb_mi) & 15);
    } else {
      live = st.sw2 != nullptr && st.add_shared;
      if (live && st.ss2)
        ssm[0] = slot + (uint32_t)k * sstride + (uint32_t)(reinterpret_cast<uintptr_t>(st.ss2 + (size_t)(i0 / P.bs0) * ncb_sh) & 15);
    }
    if (live) {
      const uint32_t wa[1] = {slot + (uint32_t)P.slot_scale + (uint32_t)k * st.seg_stride + (uint32_t)lr * (k < st.K ? rb_mi : rb_sh)};
      piece_rows<Q, 1>(wa, ssm, P.bs1, P.bs1_shift, pcd.g0, pcd.g1, xs_seg[k], q8_seg[k], lane, v);
    }
    if (lane == 0) res[lr * np + pc] = v[0];
  }
}
// x[i] += sum_k w_k * dot_k + dot_shared in the reference's order (src/infer.cpp:873-877, 899-903, 926-930)
__device__ __forceinline__ void down_tile_epilogue(const Program& P, const Stage& st, int t, const float* res, const float* actw_smem,
                                                   const int* act_smem, float xres) {
  const int i0 = t * st.rows_per_tile;
  const int nrows = min(st.rows_per_tile, P.dim - i0);
  const int lr = threadIdx.x;
  if (lr >= nrows) return;
  const int i = i0 + lr, np = st.npieces;
  const bool to_partial = P.partial != nullptr && (st.K > 0 || P.tp != 0);
  float acc = to_partial ? 0.f : xres;
  int pc = 0;
  for (int k = 0; k <= st.K; k++) {
    float v = 0.f;
    for (; pc < np && st.piece[pc].seg == k; pc++) v += res[lr * np + pc];
    if (k < st.K) {
      const int e = act_smem[k] - P.expert_first;
      if (e >= 0 && e < P.expert_count) acc = fmaf(v, actw_smem[k], acc);
    } else if (st.sw2 != nullptr && st.add_shared) {
      acc += v;
    }
  }
  if (to_partial) store_partial(P, st, i, acc); else P.x[i] = acc;
}

// ---- attention stage (one head per CTA; body of attn_kernel with consumer-only barriers) ---------------------------
__device__ __forceinline__ void c_attention(const Program& P, const Stage& st, const MegaSmem& sm, int h) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* stage = reinterpret_cast<float*>(sm.xregion);            // 512 floats
  float* qs = stage + 512;
  const Ctrl* c = P.ctrl;
  const int pos = c->pos, kv_pos = c->kv_pos, kv_len = c->kv_len, kv_sink = c->kv_sink;
  float* att = qs + ((P.hd + 3) & ~3);
  const size_t need = (size_t)(512 + ((P.hd + 3) & ~3) + ((kv_len + 3) & ~3) + kConsumers) * 4;
  if (need > (size_t)P.xregion_bytes) att = P.att_scratch + (size_t)h * (P.max_seq + kConsumers + 8);
  float* qh = P.q + (size_t)h * P.hd;
  const size_t kstride = (size_t)P.n_heads * P.hd, vstride = (size_t)P.n_heads * P.vh;
  const int half_r = P.rope >> 1;
  for (int i = tid; i < P.nope; i += kConsumers) qs[i] = qh[i];
  if (tid < half_r) {
    float cs, sn; rope_cs(P.rope_freq, tid, pos, cs, sn);
    const float v0 = qh[P.nope + 2 * tid], v1 = qh[P.nope + 2 * tid + 1];
    const float r0 = v0 * cs - v1 * sn, r1 = v0 * sn + v1 * cs;
    if (P.is_v3) { qs[P.nope + 2 * tid] = r0; qs[P.nope + 2 * tid + 1] = r1; }
    else { qs[P.nope + tid] = r0; qs[P.nope + tid + half_r] = r1; }
  } else if (tid >= 64 && tid < 64 + half_r) {
    const int t = tid - 64;
    float cs, sn; rope_cs(P.rope_freq, t, pos, cs, sn);
    const float v0 = P.kv_a[P.kv_lora + 2 * t], v1 = P.kv_a[P.kv_lora + 2 * t + 1];
    const float r0 = v0 * cs - v1 * sn, r1 = v0 * sn + v1 * cs;
    __half* kr = st.kcache + (size_t)kv_pos * kstride + (size_t)h * P.hd + P.nope;
    if (P.is_v3) { kr[2 * t] = __float2half_rn(r0); kr[2 * t + 1] = __float2half_rn(r1); }
    else { kr[t] = __float2half_rn(r0); kr[t + half_r] = __float2half_rn(r1); }
  } else if (tid >= 128 && tid < 128 + half_r * kv_sink && kv_sink > 0) {
    const int t = (tid - 128) % half_r, r = (tid - 128) / half_r;
    float cs, sn; rope_cs(P.rope_freq, t, 1, cs, sn);
    __half* kr = st.kcache + (size_t)r * kstride + (size_t)h * P.hd + P.nope;
    const float v0 = __half2float(kr[2 * t]), v1 = __half2float(kr[2 * t + 1]);
    stage[2 * (r * half_r + t)] = v0 * cs - v1 * sn;
    stage[2 * (r * half_r + t) + 1] = v0 * sn + v1 * cs;
  }
  csync();
  if (tid >= 128 && tid < 128 + half_r * kv_sink && kv_sink > 0) {
    const int t = (tid - 128) % half_r, r = (tid - 128) / half_r;
    __half* kr = st.kcache + (size_t)r * kstride + (size_t)h * P.hd + P.nope;
    const float r0 = stage[2 * (r * half_r + t)], r1 = stage[2 * (r * half_r + t) + 1];
    if (P.is_v3) { kr[2 * t] = __float2half_rn(r0); kr[2 * t + 1] = __float2half_rn(r1); }
    else { kr[t] = __float2half_rn(r0); kr[t + half_r] = __float2half_rn(r1); }
  }
  if (tid < P.rope) qh[P.nope + tid] = qs[P.nope + tid];
  csync();
  const float inv = sqrtf((float)P.hd);
  for (int t0 = warp * 8; t0 < kv_len; t0 += 64) {   // 8 cache rows per warp pass, all their loads (up to 32 per lane) in flight
    float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int c0 = 0; c0 < P.hd; c0 += 256) {
#pragma unroll
      for (int c = 0; c < 4; c++) {
        const int i = c0 + lane * 2 + 64 * c;
        if (i < P.hd) {
          const float q0 = qs[i], q1 = qs[i + 1];
          float2 kk[8];
#pragma unroll
          for (int u = 0; u < 8; u++) {
            const int t = min(t0 + u, kv_len - 1);
            kk[u] = __half22float2(*reinterpret_cast<const __half2*>(st.kcache + (size_t)t * kstride + (size_t)h * P.hd + i));
          }
#pragma unroll
          for (int u = 0; u < 8; u++) { s[u] = fmaf(q0, kk[u].x, s[u]); s[u] = fmaf(q1, kk[u].y, s[u]); }
        }
      }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
#pragma unroll
      for (int u = 0; u < 8; u++) s[u] += __shfl_xor_sync(0xffffffffu, s[u], o);
    }
#pragma unroll
    for (int u = 0; u < 8; u++) if (lane == 0 && t0 + u < kv_len) att[t0 + u] = s[u] / inv;
  }
  csync();
  float m = -3.402823466e38f;
  for (int t = tid; t < kv_len; t += kConsumers) m = fmaxf(m, att[t]);
  m = cmax(m, sm.red);
  float sum = 0.f;
  for (int t = tid; t < kv_len; t += kConsumers) { const float e = expf(att[t] - m); att[t] = e; sum += e; }
  sum = csum(sum, sm.red);
  for (int t = tid; t < kv_len; t += kConsumers) att[t] = att[t] / sum;
  csync();
  float* part = att + ((kv_len + 3) & ~3);
  if (P.vh <= kConsumers) {
    const int groups = kConsumers / P.vh;
    const int i = tid % P.vh, g = tid / P.vh;
    float acc = 0.f;
    if (g < groups) {
      const __half* vb = st.vcache + (size_t)h * P.vh + i;
      int t = g;
      for (; t + 15 * groups < kv_len; t += 16 * groups) {   // 16 independent loads in flight, accumulation order unchanged
        __half hv[16];
#pragma unroll
        for (int u = 0; u < 16; u++) hv[u] = vb[(size_t)(t + u * groups) * vstride];
#pragma unroll
        for (int u = 0; u < 16; u++) acc = fmaf(att[t + u * groups], __half2float(hv[u]), acc);
      }
      for (; t + 3 * groups < kv_len; t += 4 * groups) {
        const float v0 = __half2float(vb[(size_t)t * vstride]), v1 = __half2float(vb[(size_t)(t + groups) * vstride]);
        const float v2 = __half2float(vb[(size_t)(t + 2 * groups) * vstride]), v3 = __half2float(vb[(size_t)(t + 3 * groups) * vstride]);
        acc = fmaf(att[t], v0, acc); acc = fmaf(att[t + groups], v1, acc); acc = fmaf(att[t + 2 * groups], v2, acc); acc = fmaf(att[t + 3 * groups], v3, acc);
      }
      for (; t < kv_len; t += groups) acc = fmaf(att[t], __half2float(vb[(size_t)t * vstride]), acc);
      part[g * P.vh + i] = acc;
    }
    csync();
    if (tid < P.vh) {
      float o = 0.f;
      for (int g2 = 0; g2 < groups; g2++) o += part[g2 * P.vh + tid];
      P.xb2[(size_t)h * P.vh + tid] = o;
    }
  } else {
    for (int i = tid; i < P.vh; i += kConsumers) {
      const __half* vb = st.vcache + (size_t)h * P.vh + i;
      float acc = 0.f;
      for (int t = 0; t < kv_len; t++) acc = fmaf(att[t], __half2float(vb[(size_t)t * vstride]), acc);
      P.xb2[(size_t)h * P.vh + i] = acc;
    }
  }
  csync();
}

__device__ __forceinline__ unsigned ld_relaxed_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// ST_XCHG (after the grid barrier that follows the stage whose epilogues filled P.partial): an all-to-all of the partial vector
// in the low-latency "data + flag in one word" style.  Thread i of chunk CTA c packs {partial[i], sequence number} into ONE
// 8-byte word and stores it into every rank's exchange buffer (slot [parity][this rank][i]; 8-byte stores are single NVLink
// transactions), then polls the N words [parity][r][i] of its OWN buffer until each carries this exchange's sequence number
// and adds them to x[i] in rank order — identical on every rank, so the replicated residual stream stays bit-identical across
// GPUs.  No fences, no separate flags, no second barrier: every word validates itself, so the latency is one NVLink store.
// Two parities make the reuse safe: a rank cannot get two exchanges ahead of a peer (it needs that peer's words to finish
// the next one), so slot k & 1 is never overwritten before every reader of exchange k is done with it.
__device__ __forceinline__ void c_xchg(const Program& P, const Stage& st, const MegaSmem& sm) {
  (void)sm;
  const int tid = threadIdx.x, N = P.n_ranks;
  const unsigned seq = (unsigned)P.ctrl->pad[0] + (unsigned)st.xchg_ord + 1u;
  const int nchunk = (P.dim + kConsumers - 1) / kConsumers;
  for (int c = blockIdx.x; c < nchunk; c += gridDim.x) {
    const int i = c * kConsumers + tid;
    if (i >= P.dim) continue;
    const size_t slot = (size_t)(seq & 1u) * (size_t)N * (size_t)P.dim;
    const unsigned long long word = ((unsigned long long)seq << 32) | (unsigned long long)__float_as_uint(__ldcg(P.partial + i));
    for (int q = 0; q < N; q++) {
      unsigned long long* dst = reinterpret_cast<unsigned long long*>(P.xchg_peer[q]) + slot + (size_t)P.rank * (size_t)P.dim + (size_t)i;
      asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(dst), "l"(word) : "memory");
    }
    const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(P.xchg_peer[P.rank]) + slot + (size_t)i;
    float acc = P.x[i];
    const unsigned long long t0 = gtime();
    for (int r = 0; r < N; r++) {
      unsigned long long w;
      for (unsigned spins = 0;; spins++) {
        asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w) : "l"(mine + (size_t)r * (size_t)P.dim) : "memory");
        if ((unsigned)(w >> 32) == seq) break;
        if ((spins & 1023u) == 1023u && gtime() - t0 > 30000000000ull) __trap();   // a peer is more than 30 s late: fail, do not hang
      }
      acc += __uint_as_float((unsigned)(w & 0xffffffffull));
    }
    P.x[i] = acc;
  }
}

// ST_AMAX (tensor parallel, after the row-sharded LM head): the ranks exchange their local arg-max keys (and thereby publish the
// logits they stored into each other's buffers); the global key — largest logit, lowest index on ties, as sample_argmax —
// replaces the local one in Ctrl so the next token's embedding stage and the host read the same token on every rank.
__device__ __forceinline__ void c_amax(const Program& P, const Stage& st) {
  if (blockIdx.x != 0) return;
  const int tid = threadIdx.x;
  const unsigned seq = (unsigned)P.ctrl->pad[0] + (unsigned)st.xchg_ord + 1u;
  if (tid < P.n_ranks) {
    P.amax_peer[tid][(size_t)(seq & 1u) * (size_t)P.n_ranks + (size_t)P.rank] = P.ctrl->argmax_key;
    __threadfence_system();
    const size_t aflag = (size_t)P.n_ranks * (size_t)((P.dim + kConsumers - 1) / kConsumers);   // after the chunk flags
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(P.xflag_peer[tid] + aflag + P.rank), "r"(seq) : "memory");
    const unsigned long long t0 = gtime();
    while ((int)(ld_acquire_sys(P.xflag_peer[P.rank] + aflag + tid) - seq) < 0) {
      if (gtime() - t0 > 30000000000ull) __trap();
    }
  }
  csync();
  if (tid == 0) {
    unsigned long long best = 0ull;
    const unsigned long long* keys = P.amax_peer[P.rank] + (size_t)(seq & 1u) * (size_t)P.n_ranks;
    for (int r = 0; r < P.n_ranks; r++) { const unsigned long long k = *reinterpret_cast<const volatile unsigned long long*>(keys + r); if (k > best) best = k; }
    P.ctrl->argmax_key = best;
  }
}

// ---- embedding stage (CTA 0): token feed + step counters + dequantised row (body of embed_kernel) ------------------
__device__ __forceinline__ void c_embed(const Program& P, int from_argmax, int* s_token) {
  if (threadIdx.x == 0) {
    Ctrl* c = P.ctrl;
    int token = c->token;
    if (from_argmax) {
      // key == 0 means "no LM-head stage ran since the key was cleared" (the host rejects that call sequence; never index
      // the embedding table with it)
      if (c->argmax_key == 0ull) __trap();
      token = (int)(0xFFFFFFFFu - (unsigned)(c->argmax_key & 0xFFFFFFFFull));
      const int pos = c->pos + 1;
      const int sink = pos >= P.original_max ? 2 : 0;
      c->token = token; c->pos = pos; c->kv_sink = sink;
      c->kv_pos = sink + (pos - sink) % (P.original_max - sink);
      c->kv_len = pos >= P.original_max ? P.original_max : pos + 1;
      if (P.token_log && P.step) { P.token_log[*P.step] = token; *P.step = *P.step + 1; }
      c->pad[0] += P.n_xchg;   // peer-memory mode: exchanges completed before this token
    }
    c->argmax_key = 0ull;
    *s_token = token;
  }
  csync();
  const int token = *s_token, dim = P.dim;
  const uint8_t* table = P.embed_w;
  for (int i = threadIdx.x; i < dim; i += kConsumers) {
    float val;
    switch (P.embed_quant) {
      case Q_F32: val = reinterpret_cast<const float*>(table)[(size_t)token * dim + i]; break;
      case Q_F16: val = __half2float(reinterpret_cast<const __half*>(table)[(size_t)token * dim + i]); break;
      case Q_F8: {
        const int ncb = (dim + P.bs1 - 1) / P.bs1;
        val = h2f((uint16_t)((uint16_t)table[(size_t)token * f8_pitch((size_t)dim) + i] << 8)) * P.embed_scale[(size_t)(token / P.bs0) * ncb + i / P.bs1];
        break;
      }
      case Q_Q2K: {
        const int nb = dim >> 8, b = i >> 8, w = i & 255, hh = w >> 7, s = (w >> 5) & 3, l = w & 31;
        const uint8_t* blk = table + ((size_t)token * nb + b) * kQ2Bytes;
        const int sc = blk[8 * hh + 2 * s + (l >> 4)];
        const int qv = (blk[16 + 32 * hh + l] >> (2 * s)) & 3;
        val = (h2f(*reinterpret_cast<const uint16_t*>(blk + 80)) * (float)(sc & 0xF)) * (float)qv -
              h2f(*reinterpret_cast<const uint16_t*>(blk + 82)) * (float)(sc >> 4);
        break;
      }
      default: {
        const int nb = dim >> 8, b = i >> 8, w = i & 255, hh = w >> 7, s = (w >> 5) & 3, l = w & 31;
        const uint8_t* blk = table + ((size_t)token * nb + b) * kQ3Bytes;
        const int j = 8 * hh + 2 * s + (l >> 4);
        const int lob = blk[96 + (j & 7)];
        const int lo4 = j < 8 ? (lob & 0xF) : (lob >> 4);
        const int hi2 = (blk[96 + 8 + (j & 3)] >> (2 * (j >> 2))) & 3;
        const int hb = (blk[l] >> (4 * hh + s)) & 1;
        const int qv = ((blk[32 + 32 * hh + l] >> (2 * s)) & 3) - (hb ? 0 : 4);
        val = (h2f(*reinterpret_cast<const uint16_t*>(blk + 108)) * (float)((lo4 | (hi2 << 4)) - 32)) * (float)qv;
        break;
      }
    }
    P.x[i] = val;
  }
}

// ---- per-stage bodies ------------------------------------------------------------------------------------------
// GEMV activation staging: one pass over the input with the values kept in registers (n <= 8192), RMSNorm and Q8_K fused
template <int Q>
__device__ __forceinline__ void c_stage_gemv_input(const Program& P, const Stage& st, const MegaSmem& sm, float* xs0, const Q8Smem& q80, int dep_count, int stage_index) {
  constexpr bool KQ = QTraits<Q>::kq;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = st.n;
  if (n > 8192) {   // long vectors: two passes through L1/L2
    float sc = 1.0f;
    if (st.norm_w) sc = c_rms_scale(st.in, n, P.eps, sm.red);
    if (st.need_topk) stage_route_hook(P, st, dep_count, stage_index);   // thread-per-expert routing; releases the producer
    c_stage_vec<Q>(st.in, n, st.norm_w, sc, xs0, q80);
    return;
  }
  if constexpr (KQ) {
    float v[4][8];
    const int nb = n >> 8;
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int b = warp + 8 * k;
      if (b < nb) {
        const float4 a0 = *reinterpret_cast<const float4*>(st.in + (b << 8) + lane * 8);
        const float4 a1 = *reinterpret_cast<const float4*>(st.in + (b << 8) + lane * 8 + 4);
        v[k][0] = a0.x; v[k][1] = a0.y; v[k][2] = a0.z; v[k][3] = a0.w; v[k][4] = a1.x; v[k][5] = a1.y; v[k][6] = a1.z; v[k][7] = a1.w;
#pragma unroll
        for (int j = 0; j < 8; j++) ss = fmaf(v[k][j], v[k][j], ss);
      }
    }
    float sc = 1.0f;
    if (st.norm_w) { ss = csum(ss, sm.red); sc = 1.0f / sqrtf(ss / (float)n + P.eps); }
    if (st.need_topk) stage_route_hook(P, st, dep_count, stage_index);   // thread-per-expert routing; releases the producer
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int b = warp + 8 * k;
      if (b < nb) {
        if (st.norm_w) {
          const float4 w0 = *reinterpret_cast<const float4*>(st.norm_w + (b << 8) + lane * 8);
          const float4 w1 = *reinterpret_cast<const float4*>(st.norm_w + (b << 8) + lane * 8 + 4);
          const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
          for (int j = 0; j < 8; j++) v[k][j] = __fmul_rn(__fmul_rn(v[k][j], sc), ww[j]);
        }
        q8_block(v[k], b, lane, q80);
      }
    }
  } else {
    float4 v[8];
    const int nf = n >> 2;
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int f = tid + k * kConsumers;
      if (f < nf) {
        v[k] = reinterpret_cast<const float4*>(st.in)[f];
        ss = fmaf(v[k].x, v[k].x, ss); ss = fmaf(v[k].y, v[k].y, ss); ss = fmaf(v[k].z, v[k].z, ss); ss = fmaf(v[k].w, v[k].w, ss);
      }
    }
    float sc = 1.0f;
    if (st.norm_w) { ss = csum(ss, sm.red); sc = 1.0f / sqrtf(ss / (float)n + P.eps); }
    if (st.need_topk) stage_route_hook(P, st, dep_count, stage_index);   // thread-per-expert routing; releases the producer
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int f = tid + k * kConsumers;
      if (f < nf) {
        float4 o = v[k];
        if (st.norm_w) {
          const float4 w = reinterpret_cast<const float4*>(st.norm_w)[f];
          o.x = __fmul_rn(__fmul_rn(o.x, sc), w.x); o.y = __fmul_rn(__fmul_rn(o.y, sc), w.y);
          o.z = __fmul_rn(__fmul_rn(o.z, sc), w.z); o.w = __fmul_rn(__fmul_rn(o.w, sc), w.w);
        }
        reinterpret_cast<float4*>(xs0)[xswz<Q>(f)] = o;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Warp-per-tile stages (F8E5M2 through the tensor cores).  A tile is small enough (<= slot_data bytes) that ONE warp
// reduces it over the full K range with the accumulators in registers: no per-tile CTA barrier, no partial-sum
// exchange, and with 8 consumer warps up to 8 tiles are being reduced while the producer keeps the other slots filling.
//   ST_GEMV: tile = rows [r0, r0+RT) of one job (RT <= 16; EPI_GLU: RT <= 8 gate rows + the same RT up rows)
//   ST_DOWN: tile = one PIECE (segment k, rows [pr0, pr0+pnr) of an 8-row output group); the warp that completes the
//            group's last piece combines the partial sums in the reference's order and applies the residual update.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void wp_gemv_tile(const Program& P, const Stage& st, int t, uint32_t slot, const X16& x16,
                                             const int* act_smem, unsigned long long& best, long long& c_mma) {
  const int lane = threadIdx.x & 31, gid = lane >> 2, tig = lane & 3;
  int j = 0;
  while (j + 1 < st.njobs && t >= st.job[j + 1].tile_begin) j++;
  const MJob& jb = st.job[j];
  const int r0 = (t - jb.tile_begin) * st.rows_per_tile;
  const int nrows = min(st.rows_per_tile, jb.rows - r0);
  size_t soff = 0;
  if (jb.expert_slot >= 0) {
    const int e = act_smem[jb.expert_slot] - P.expert_first;
    if (e < 0 || e >= P.expert_count) return;
    soff = (size_t)e * jb.s_stride;
  }
  const bool glu = st.epi == EPI_GLU;
  const uint32_t rb = (uint32_t)f8_pitch((size_t)st.n);   // device row pitch of f8 weights
  const uint32_t part_stride = (uint32_t)align_up((size_t)st.rows_per_tile * rb, 128);
  uint32_t s0 = 0, s1 = 0;
  if (jb.scale) {
    const int ncb = (st.n + P.bs1 - 1) / P.bs1;
    const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(jb.scale + soff + (size_t)(r0 / P.bs0) * ncb) & 15);
    s0 = slot + sh;
    s1 = glu ? slot + (uint32_t)P.slot_scale / 2 + sh : s0;
  }
  const uint32_t data = slot + (uint32_t)P.slot_scale;
  const int r_lo = gid, r_hi = glu ? gid : gid + 8;
  const uint32_t a_lo = data + (uint32_t)min(r_lo, nrows - 1) * rb;
  uint32_t a_hi = (glu ? data + part_stride : data) + (uint32_t)min(r_hi, nrows - 1) * rb;
  if (!glu && nrows <= 8) a_hi = 0;   // no upper rows in this tile: skip their loads
  // residual operand first: its L2 latency hides behind the K loop
  float xres_lo = 0.f, xres_hi = 0.f;
  if (st.epi == EPI_RESID && tig == 0) {
    if (r_lo < nrows) xres_lo = jb.out[r0 + r_lo];
    if (r_hi < nrows) xres_hi = jb.out[r0 + r_hi];
  }
  float v_lo, v_hi;
  { const long long km = clock64(); const float2 vv = mma_rows_f8(a_lo, a_hi, s0, s1, P.bs1_shift, 0, st.n, x16.hi, x16.lo, x16.gs); v_lo = vv.x; v_hi = vv.y; c_mma += clock64() - km; }
  if (tig != 0) return;
  if (glu) {
    if (r_lo < nrows) jb.out[r0 + r_lo] = (P.act_silu ? silu_f(v_lo) : gelu_f(v_lo)) * v_hi;
    return;
  }
#pragma unroll
  for (int half = 0; half < 2; half++) {
    const int lr = half ? r_hi : r_lo;
    if (lr >= nrows) continue;
    const int r = r0 + lr;
    const float val = half ? v_hi : v_lo;
    switch (st.epi) {
      case EPI_RESID: jb.out[r] = (half ? xres_hi : xres_lo) + val; break;
      case EPI_PARTIAL: store_partial(P, st, r, val); break;
      case EPI_KVB: {
        jb.out[r] = val;
        const int per = P.nope + P.vh, hh = r / per, ii = r - hh * per;
        const int kv_pos = P.ctrl->kv_pos;
        if (ii < P.nope) st.kcache[(size_t)kv_pos * P.n_heads * P.hd + hh * P.hd + ii] = __float2half_rn(val);
        else st.vcache[(size_t)kv_pos * P.n_heads * P.vh + hh * P.vh + (ii - P.nope)] = __float2half_rn(val);
        break;
      }
      case EPI_LOGITS: {
        store_logit(P, jb, r, val);
        const unsigned long long key = ((unsigned long long)orderable(val) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)(jb.row_base + r));
        if (key > best) best = key;
        break;
      }
      default: jb.out[r] = val; break;
    }
  }
}

// producer side of a warp-per-tile DOWN piece: rows [i0+pr0, +pnr) of segment k, whole rows (contiguous bytes)
__device__ __forceinline__ void wp_produce_down_piece(const Program& P, const Stage& st, int rg, int pc, uint32_t slot, uint32_t full,
                                                      const int* act_smem) {
  const Piece pcd = st.piece[pc];
  const int k = pcd.seg, i0 = rg * st.down_rows + pcd.g0;
  const int nrows = min(pcd.g1, P.dim - i0);
  const bool routed = k < st.K;
  int e = 0;
  if (routed) {
    e = act_smem[k] - P.expert_first;
    if (e < 0 || e >= P.expert_count) { mbar_expect_tx(full, 0); return; }
  } else if (!(st.sw2 != nullptr && st.add_shared)) { mbar_expect_tx(full, 0); return; }
  if (nrows <= 0) { mbar_expect_tx(full, 0); return; }
  const int n = routed ? st.mi : st.sh;
  const size_t pitch = f8_pitch((size_t)n);
  const uint32_t bytes = (uint32_t)(nrows * pitch);
  const uint8_t* src = routed ? st.w2 + (size_t)e * st.w2_stride + (size_t)i0 * pitch : st.sw2 + (size_t)i0 * pitch;
  const float* sc = routed ? st.s2 : st.ss2;
  uint32_t total = bytes, sbytes = 0, shift = 0;
  const float* ssrc = nullptr;
  if (sc) {
    const int ncb = (n + P.bs1 - 1) / P.bs1;
    sbytes = scale_copy_bytes(sc + (routed ? (size_t)e * st.s2_stride : 0) + (size_t)(i0 / P.bs0) * ncb, ncb, ssrc, shift);
    total += sbytes;
  }
  mbar_expect_tx(full, total);
  bulk_g2s(slot + (uint32_t)P.slot_scale, src, bytes, full);
  if (sbytes) bulk_g2s(slot, ssrc, sbytes, full);
}

// consumer side: one warp reduces the piece; the last piece of a row group triggers the ordered combine
__device__ __forceinline__ void wp_down_piece(const Program& P, const Stage& st, const MegaSmem& sm, int rg, int rg_local, int pc,
                                              uint32_t slot, const X16* x16_seg) {
  const int lane = threadIdx.x & 31, gid = lane >> 2, tig = lane & 3;
  const Piece pcd = st.piece[pc];
  const int k = pcd.seg, i0 = rg * st.down_rows + pcd.g0;
  const int nrows = min(pcd.g1, P.dim - i0);
  const bool routed = k < st.K;
  bool live = nrows > 0;
  int e = 0;
  if (routed) { e = sm.act[k] - P.expert_first; live = live && e >= 0 && e < P.expert_count; }
  else live = live && st.sw2 != nullptr && st.add_shared;
  const int np = st.npieces;
  // [row groups in flight][np <= 16][16 rows]: a window of n_slots <= 8 consecutive pieces touches up to (8 - 2) / np + 2 row
  // groups, each of which needs its own partial sums and completion counter until its last piece has been combined
  float* part = sm.res + (size_t)(rg_local & (st.down_nbuf - 1)) * (size_t)(np * 16);
  int* cnt = sm.sel + (rg_local & (st.down_nbuf - 1));
  if (live) {
    const int n = routed ? st.mi : st.sh;
    const float* sc = routed ? st.s2 : st.ss2;
    uint32_t ssm = 0;
    if (sc) {
      const int ncb = (n + P.bs1 - 1) / P.bs1;
      ssm = slot + (uint32_t)(reinterpret_cast<uintptr_t>(sc + (routed ? (size_t)e * st.s2_stride : 0) + (size_t)(i0 / P.bs0) * ncb) & 15);
    }
    const uint32_t data = slot + (uint32_t)P.slot_scale;
    const uint32_t a_lo = data + (uint32_t)min(gid, nrows - 1) * (uint32_t)f8_pitch((size_t)n);
    const uint32_t a_hi = nrows > 8 ? data + (uint32_t)min(gid + 8, nrows - 1) * (uint32_t)f8_pitch((size_t)n) : 0u;
    float v_lo, v_hi;
    { const X16 xk = x16_seg[k]; const float2 vv = mma_rows_f8(a_lo, a_hi, ssm, ssm, P.bs1_shift, 0, n, xk.hi, xk.lo, xk.gs); v_lo = vv.x; v_hi = vv.y; }
    if (tig == 0 && gid < nrows) part[pc * 16 + pcd.g0 + gid] = v_lo;
    if (tig == 0 && gid + 8 < nrows) part[pc * 16 + pcd.g0 + gid + 8] = v_hi;
  }
  __syncwarp();
  int last = 0;
  if (lane == 0) { __threadfence_block(); last = atomicAdd(cnt, 1) == np - 1; }
  last = __shfl_sync(0xffffffffu, last, 0);
  if (!last) return;
  __threadfence_block();
  // combine: lane r < rows of the group; segments in the reference's order (src/infer.cpp:873-877, 899-903, 926-930)
  const int gi0 = rg * st.down_rows;
  const int grows = min(st.down_rows, P.dim - gi0);
  if (lane < grows) {
    const int i = gi0 + lane;
    const bool to_partial = P.partial != nullptr && (st.K > 0 || P.tp != 0);
    float acc = to_partial ? 0.f : P.x[i];
    int p2 = 0;
    for (int kk = 0; kk <= st.K; kk++) {
      float v = 0.f;
      bool any = false;
      for (; p2 < np && st.piece[p2].seg == kk; p2++) {
        const Piece q = st.piece[p2];
        if (lane >= q.g0 && lane < q.g0 + q.g1) { v += part[p2 * 16 + lane]; any = true; }
      }
      if (!any) continue;
      if (kk < st.K) {
        const int ee = sm.act[kk] - P.expert_first;
        if (ee >= 0 && ee < P.expert_count) acc = fmaf(v, sm.actw[kk], acc);
      } else if (st.sw2 != nullptr && st.add_shared) acc += v;
    }
    if (to_partial) store_partial(P, st, i, acc); else P.x[i] = acc;
  }
  __syncwarp();
  if (lane == 0) *cnt = 0;
}

// ------------------------------------------------------------------------------------------------------------------
// Warp-per-tile K-quant stages (Q2_K / Q3_K x Q8_K, w2a8 / w3a8 integer dots — src/quant.cpp:434-783).
// Lane l owns quarter-blocks l, l+32, ... of every row (block b = qb/4, 128-half h, 16-byte half c), so its slice of the
// Q8_K activation vector — 64 int8, 4 block sums, the block scale — is loaded into REGISTERS once per stage (per piece
// for ST_DOWN) and reused for every weight row: the shared-memory pipe only carries the weight tile itself.
// Rows are reduced one after another (integer dp4a per quarter block, exact int32 block sums via 2 shuffles, fp32 across
// blocks, one warp_sum per row); lane r keeps the result of tile row r, so the epilogue runs 32 rows wide.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kKqMaxPass = 16;  // warp-per-tile K-quant rows up to 512 quarter-blocks (n <= 32768); passes are a rolled loop
struct YRegs { int4 y[4]; int bs[4]; int bs01, bs23; float d; };

__device__ __forceinline__ int dp4a_us(uint32_t a, int b, int c) {   // unsigned bytes x signed bytes
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

// this lane's slice of the Q8_K activations for quarter block qb: 64 int8, 4 sub-block sums, scale
__device__ __forceinline__ void kq_load_y(uint32_t a_qs, uint32_t a_d, uint32_t a_bsums, int nb, int qb, YRegs& yr) {
  const int b = min(qb >> 2, nb - 1), h = (qb >> 1) & 1, c = qb & 1;
  const uint32_t y = a_qs + (uint32_t)(b * 256 + 128 * h + 16 * c);
#pragma unroll
  for (int s = 0; s < 4; s++) {
    const uint4 t = lds128(y + 32u * s);
    yr.y[s] = make_int4((int)t.x, (int)t.y, (int)t.z, (int)t.w);
    short bsv;
    asm volatile("ld.shared.s16 %0, [%1];" : "=h"(bsv) : "r"(a_bsums + (uint32_t)(b * 16 + 8 * h + c + 2 * s) * 2u));
    yr.bs[s] = (int)bsv;
  }
  yr.bs01 = (yr.bs[0] & 0xffff) | (yr.bs[1] << 16);
  yr.bs23 = (yr.bs[2] & 0xffff) | (yr.bs[3] << 16);
  yr.d = __uint_as_float(lds32(a_d + (uint32_t)b * 4u));
}

// one row x one quarter block (64 weights); returns the quarter's fp32 contribution (0 for lanes past the end of the row)
template <int Q>
__device__ __forceinline__ float kq_quarter(uint32_t row, int qb, int nqb, const YRegs& yr) {
  // branch-free: lanes past the end of the row (qb >= nqb) read the last block and contribute zero, so the four rows a
  // caller interleaves form one basic block
  const bool act = qb < nqb;
  const int qc = act ? qb : nqb - 1;
  const int b = qc >> 2, h = (qc >> 1) & 1, c = qc & 1;
  float out = 0.f;
  if constexpr (Q == Q_Q2K) {
    // ggml_vec_dot_q2_K_q8_K (src/quant.cpp:666-783) for sub-blocks j = 8h + 2s + c, s = 0..3.  The 2-bit fields are NOT
    // shifted down: dp4a of the masked word (u8 = q << 2s) against the int8 activations gives 4^s x the sub-block dot, an
    // exact multiple, shifted back once per sub-block — one LOP3 per dp4a instead of SHF + LOP3.  The four scale bytes are
    // gathered with one PRMT; the four (min x block-sum) products are two dp2a.
    const uint32_t blk = row + (uint32_t)b * kQ2Bytes;
    const uint32_t qp = blk + 16 + 32 * h + 16 * c;
    const uint32_t w0 = lds32(qp), w1 = lds32(qp + 4), w2 = lds32(qp + 8), w3 = lds32(qp + 12);
    const uint32_t sA = lds32(blk + 8 * h), sB = lds32(blk + 8 * h + 4);
    const uint32_t dm = lds32(blk + 80);
    const uint32_t pack = __byte_perm(sA, sB, c ? 0x7531u : 0x6420u);   // scale bytes of s = 0..3
    const uint32_t sc4 = pack & 0x0F0F0F0Fu, m4 = (pack >> 4) & 0x0F0F0F0Fu;
    int isum = 0;
#pragma unroll
    for (int s = 0; s < 4; s++) {
      const uint32_t mk = 0x03030303u << (2 * s);
      int dp = dp4a_us(w0 & mk, yr.y[s].x, 0);
      dp = dp4a_us(w1 & mk, yr.y[s].y, dp);
      dp = dp4a_us(w2 & mk, yr.y[s].z, dp);
      dp = dp4a_us(w3 & mk, yr.y[s].w, dp);
      isum += (int)__byte_perm(sc4, 0u, 0x4440u + s) * (dp >> (2 * s));   // scale byte s (one PRMT); dp is an exact multiple of 4^s
    }
    int summs = __dp2a_lo(yr.bs01, (int)m4, 0);
    summs = __dp2a_hi(yr.bs23, (int)m4, summs);
    // per-quarter fp32 contribution (the integer sums are exact in fp32; the four quarters of a block are added in fp32 by the
    // row reduction instead of in int32 first: same value up to fp32 re-association)
    const float o = (yr.d * h2f((uint16_t)(dm & 0xffff))) * (float)isum - (yr.d * h2f((uint16_t)(dm >> 16))) * (float)summs;
    out = act ? o : 0.f;
  } else {
    int isum = 0;
    const uint32_t blk = row + (uint32_t)b * kQ3Bytes;
    const uint4 hm = lds128(blk + 16 * c);
    const uint4 qq = lds128(blk + 32 + 32 * h + 16 * c);
    const uint32_t s0 = lds32(blk + 96), s1 = lds32(blk + 100), s2 = lds32(blk + 104);
    const uint32_t dw = lds32(blk + 108);
#pragma unroll
    for (int s = 0; s < 4; s++) {
      const int bit = 4 * h + s;
      int dp = __dp4a((int)(((qq.x >> (2 * s)) & 0x03030303u) | (((hm.x >> bit) & 0x01010101u) << 2)), yr.y[s].x, 0);
      dp = __dp4a((int)(((qq.y >> (2 * s)) & 0x03030303u) | (((hm.y >> bit) & 0x01010101u) << 2)), yr.y[s].y, dp);
      dp = __dp4a((int)(((qq.z >> (2 * s)) & 0x03030303u) | (((hm.z >> bit) & 0x01010101u) << 2)), yr.y[s].z, dp);
      dp = __dp4a((int)(((qq.w >> (2 * s)) & 0x03030303u) | (((hm.w >> bit) & 0x01010101u) << 2)), yr.y[s].w, dp);
      dp -= 4 * yr.bs[s];
      const int t = 2 * s + c;
      const uint32_t lw = (t < 4) ? s0 : s1;
      const int lob = (lw >> (8 * (t & 3))) & 0xff;
      const int lo4 = h ? (lob >> 4) : (lob & 0xF);
      const int hib = (s2 >> (8 * (t & 3))) & 0xff;
      const int hi2 = (hib >> (2 * (2 * h + (t >> 2)))) & 3;
      isum += ((lo4 | (hi2 << 4)) - 32) * dp;
    }
    const float o = (h2f((uint16_t)(dw & 0xffff)) * yr.d) * (float)isum;
    out = act ? o : 0.f;
  }
  return out;
}

// all rows of a tile: returns in lane r the dot product of tile row r (r < nrows <= 32).  ONE out-of-line copy per
// quant for every GEMV / DOWN stage (instruction-cache footprint).
// Lane mapping: a row of nqb quarter blocks is spread over L = 32 / G lanes, where G = 4, 2 or 1 rows share a warp pass
// (short rows — kv_b's 512 columns are 8 quarter blocks — would otherwise leave 24 of 32 lanes idle); lane l works on quarter
// (l mod L) + L*p of row group member l / L.  Four such row groups are interleaved per iteration (independent dp4a chains), so an
// iteration covers 4 G rows; passes over long rows are a rolled loop that re-reads this lane's activation slice from shared
// memory (kept in registers across the whole tile when one pass covers the row).
template <int Q>
__device__ __noinline__ float kq_tile_rows(uint32_t base, uint32_t rb, int nrows, int nb, uint32_t q_qs, uint32_t q_d, uint32_t q_bsums) {
  const int lane = threadIdx.x & 31;
  const int nqb = nb * 4;
  const int G = nqb <= 8 ? 4 : (nqb <= 16 ? 2 : 1), L = 32 / G;
  const int sub = lane / L, ql = lane & (L - 1);
  const int npass = (nqb + L - 1) / L;
  float mine = 0.f;
  YRegs yr;
  if (npass == 1) kq_load_y(q_qs, q_d, q_bsums, nb, ql, yr);
#pragma unroll 1
  for (int r0 = 0; r0 < nrows; r0 += 4 * G) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
    for (int p = 0; p < npass; p++) {
      if (npass > 1) kq_load_y(q_qs, q_d, q_bsums, nb, ql + L * p, yr);
#pragma unroll
      for (int i = 0; i < 4; i++) acc[i] += kq_quarter<Q>(base + (uint32_t)min(r0 + i * G + sub, nrows - 1) * rb, ql + L * p, nqb, yr);
    }
#pragma unroll 1
    for (int o = L >> 1; o; o >>= 1) {
#pragma unroll
      for (int i = 0; i < 4; i++) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
    }
    // lane r keeps row r: row r0 + i G + m was reduced by the lanes [m L, (m + 1) L)
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int rel = lane - r0 - i * G;
      const float t = __shfl_sync(0xffffffffu, acc[i], (rel & (G - 1)) * L);
      if (rel >= 0 && rel < G && lane < nrows) mine = t;
    }
  }
  return mine;
}

template <int Q>
__device__ __forceinline__ void wp_kq_gemv_tile(const Program& P, const Stage& st, int t, uint32_t slot, const Q8Smem& q8,
                                                const int* act_smem, unsigned long long& best) {
  const int lane = threadIdx.x & 31;
  int j = 0;
  while (j + 1 < st.njobs && t >= st.job[j + 1].tile_begin) j++;
  const MJob& jb = st.job[j];
  const int r0 = (t - jb.tile_begin) * st.rows_per_tile;
  const int nrows = min(st.rows_per_tile, jb.rows - r0);
  if (jb.expert_slot >= 0) {
    const int e = act_smem[jb.expert_slot] - P.expert_first;
    if (e < 0 || e >= P.expert_count) return;
  }
  const bool glu = st.epi == EPI_GLU;
  const uint32_t rb = (uint32_t)QTraits<Q>::row_bytes(st.n);
  const uint32_t part_stride = (uint32_t)align_up((size_t)st.rows_per_tile * rb, 128);
  const uint32_t data = slot + (uint32_t)P.slot_scale;
  const int nb = st.n >> 8;
  float xres = 0.f;
  if (st.epi == EPI_RESID && lane < nrows) xres = jb.out[r0 + lane];
  const float v = kq_tile_rows<Q>(data, rb, nrows, nb, smem_u32(q8.qs), smem_u32(q8.d), smem_u32(q8.bsums));
  float u = 0.f;
  if (glu) u = kq_tile_rows<Q>(data + part_stride, rb, nrows, nb, smem_u32(q8.qs), smem_u32(q8.d), smem_u32(q8.bsums));
  if (lane >= nrows) return;
  const int r = r0 + lane;
  float val = v;
  if (glu) val = (P.act_silu ? silu_f(v) : gelu_f(v)) * u;
  switch (st.epi) {
    case EPI_RESID: jb.out[r] = xres + val; break;
    case EPI_PARTIAL: store_partial(P, st, r, val); break;
    case EPI_KVB: {
      jb.out[r] = val;
      const int per = P.nope + P.vh, hh = r / per, ii = r - hh * per;
      const int kv_pos = P.ctrl->kv_pos;
      if (ii < P.nope) st.kcache[(size_t)kv_pos * P.n_heads * P.hd + hh * P.hd + ii] = __float2half_rn(val);
      else st.vcache[(size_t)kv_pos * P.n_heads * P.vh + hh * P.vh + (ii - P.nope)] = __float2half_rn(val);
      break;
    }
    case EPI_LOGITS: {
      store_logit(P, jb, r, val);
      const unsigned long long key = ((unsigned long long)orderable(val) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)(jb.row_base + r));
      if (key > best) best = key;
      break;
    }
    default: jb.out[r] = val; break;
  }
}

// ST_DOWN for K-quants: ONE tile = a row group of DR = st.down_rows output rows x ALL segments (the K routed experts' w2 rows, then
// the shared / dense rows), K + 1 bulk copies on one mbarrier, reduced by ONE warp: lane r accumulates output row r over the
// segments in the reference's order (src/infer.cpp:873-877, 899-903, 926-930) — no cross-warp partial sums, no completion
// counters, and a handful of tiles per CTA instead of (row groups x segments) small pieces, whose per-tile cost on the single
// producer thread (~0.5 us each) used to bound the stage.  Segment k lives at k * st.seg_stride inside the tile.
template <int Q>
__device__ __forceinline__ void kq_produce_down_group(const Program& P, const Stage& st, int rg, uint32_t slot, uint32_t full,
                                                      const int* act_smem) {
  const int i0 = rg * st.down_rows;
  const int nrows = min(st.down_rows, P.dim - i0);
  const uint32_t rb_mi = (uint32_t)QTraits<Q>::row_bytes(st.mi), rb_sh = (uint32_t)QTraits<Q>::row_bytes(st.sh);
  const uint32_t b_mi = (uint32_t)align_up((size_t)nrows * rb_mi, 16), b_sh = (uint32_t)align_up((size_t)nrows * rb_sh, 16);
  const bool use_shared = st.sw2 != nullptr && st.add_shared && st.sh > 0;
  uint32_t live = 0, total = 0;
  for (int k = 0; k < st.K; k++) {
    const int e = act_smem[k] - P.expert_first;
    if (e >= 0 && e < P.expert_count) { live |= 1u << k; total += b_mi; }
  }
  if (use_shared) total += b_sh;
  mbar_expect_tx(full, total);
  for (int k = 0; k < st.K; k++) {
    if (!((live >> k) & 1u)) continue;
    const int e = act_smem[k] - P.expert_first;
    bulk_g2s(slot + (uint32_t)k * (uint32_t)st.seg_stride, st.w2 + (size_t)e * st.w2_stride + (size_t)i0 * rb_mi, b_mi, full);
  }
  if (use_shared) bulk_g2s(slot + (uint32_t)st.K * (uint32_t)st.seg_stride, st.sw2 + (size_t)i0 * rb_sh, b_sh, full);
}

template <int Q>
__device__ __forceinline__ void kq_down_group(const Program& P, const Stage& st, const MegaSmem& sm, int rg, uint32_t slot,
                                              const Q8Smem* q8_seg) {
  const int lane = threadIdx.x & 31;
  const int i0 = rg * st.down_rows;
  const int nrows = min(st.down_rows, P.dim - i0);
  const bool to_partial = P.partial != nullptr && (st.K > 0 || P.tp != 0);
  float acc = 0.f;
  if (!to_partial && lane < nrows) acc = P.x[i0 + lane];   // residual operand first: its L2 latency hides behind the dots
  const uint32_t rb_mi = (uint32_t)QTraits<Q>::row_bytes(st.mi);
#pragma unroll 1
  for (int k = 0; k < st.K; k++) {
    const int e = sm.act[k] - P.expert_first;
    if (e < 0 || e >= P.expert_count) continue;
    const Q8Smem qk = q8_seg[k];
    const float v = kq_tile_rows<Q>(slot + (uint32_t)k * (uint32_t)st.seg_stride, rb_mi, nrows, st.mi >> 8, smem_u32(qk.qs), smem_u32(qk.d), smem_u32(qk.bsums));
    acc = fmaf(v, sm.actw[k], acc);
  }
  if (st.sw2 != nullptr && st.add_shared && st.sh > 0) {
    const Q8Smem qk = q8_seg[st.K];
    acc += kq_tile_rows<Q>(slot + (uint32_t)st.K * (uint32_t)st.seg_stride, (uint32_t)QTraits<Q>::row_bytes(st.sh), nrows, st.sh >> 8, smem_u32(qk.qs), smem_u32(qk.d), smem_u32(qk.bsums));
  }
  if (lane < nrows) { if (to_partial) store_partial(P, st, i0 + lane, acc); else P.x[i0 + lane] = acc; }
}

// the tile loop of a warp-per-tile K-quant GEMV stage; this lane's activation slice lives in NP register sets
template <int Q>
__device__ __forceinline__ void kq_gemv_loop(const Program& P, const Stage& st, const MegaSmem& sm, const Q8Smem& q80, int& it,
                                             unsigned long long& best_key, int stage_index) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long c_wait = 0, c_task = 0;
  int n_mine = 0;
  const int it0 = it, cnt = my_tile_count(st.ntiles);
  for (;;) {
    const int j = ring_claim(sm);
    if (j >= it0 + cnt) break;
    const int t = (int)blockIdx.x + (j - it0) * (int)gridDim.x;
    const long long k0 = clock64();
    const uint32_t slot = ring_wait(sm, j);
    const long long k1 = clock64();
    wp_kq_gemv_tile<Q>(P, st, t, slot, q80, sm.act, best_key);
    __syncwarp();
    if (lane == 0) ring_release(sm, j);
    c_wait += k1 - k0; c_task += clock64() - k1; n_mine++;
  }
  it = it0 + cnt;
  if (lane == 0 && blockIdx.x == 0 && P.tstamp && n_mine && !st.need_topk) {   // profiling: cycles per tile waiting / reducing
    if (warp == 0) { P.tstamp[stage_index * 8 + 4] = (unsigned long long)(c_wait / n_mine); P.tstamp[stage_index * 8 + 5] = (unsigned long long)(c_task / n_mine); }
    if (warp == 1) { P.tstamp[stage_index * 8 + 6] = (unsigned long long)(c_wait / n_mine); P.tstamp[stage_index * 8 + 7] = (unsigned long long)(c_task / n_mine); }
  }
}

// ---- shared activation staging (ONE out-of-line copy each, used by every GEMV stage and by the DOWN stage) -------------
// The code of a stage phase runs once per stage with a cold instruction cache, so its cost is its size (tools/
// icache_bench.cu: ~0.2 us per KB fetched from L2): both routines are deliberately small, and the DOWN stage reuses them
// on a flat layout (routed hidden vectors concatenated: K*mi values, then the shared/dense vector).
//   n floats at `in` -> RMSNorm (optional) -> fp16 hi/lo split (stage_x16) or Q8_K blocks (stage_q8); `route` >= 0: warp 0
//   runs the MoE routing between the norm reduction and its share of the conversion and releases the producer (dep count).
__device__ __noinline__ void stage_x16(const Program* Pp, const Stage* stp, const float* __restrict__ in, int n, const float* __restrict__ norm_w,
                                       uint32_t xhi, uint32_t xlo, uint32_t xgs, int route, int stage_index) {
  extern __shared__ __align__(128) unsigned char dsk_dyn_smem[];
  const Program& P = *Pp; const Stage& st = *stp;
  float* red = reinterpret_cast<float*>(dsk_dyn_smem + 256);
  const int tid = threadIdx.x, nf = n >> 2;
  float sc = 1.0f;
  bool sc_known = norm_w == nullptr;
  if (norm_w && nf > 8 * kConsumers) { sc = c_rms_scale(in, n, P.eps, red); sc_known = true; }   // long vectors: extra pass
  bool routed = route < 0;
#pragma unroll 1
  for (int c0 = 0; c0 < nf; c0 += 8 * kConsumers) {   // eight float4 per thread in flight
    float4 v[8];
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int f = c0 + tid + k * kConsumers;
      v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (f < nf) {
        v[k] = reinterpret_cast<const float4*>(in)[f];
        ss = fmaf(v[k].x, v[k].x, ss); ss = fmaf(v[k].y, v[k].y, ss); ss = fmaf(v[k].z, v[k].z, ss); ss = fmaf(v[k].w, v[k].w, ss);
      }
    }
    if (!sc_known) { ss = csum(ss, red); sc = 1.0f / sqrtf(ss / (float)n + P.eps); sc_known = true; }
    if (!routed) { stage_route_hook(P, st, route, stage_index); routed = true; }
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int f = c0 + tid + k * kConsumers;
      if (f < nf) {   // nf is a multiple of 16, so the 16 lanes of a 64-column group are in or out together
        float4 o = v[k];
        if (norm_w) {
          const float4 w = reinterpret_cast<const float4*>(norm_w)[f];
          o.x = __fmul_rn(__fmul_rn(o.x, sc), w.x); o.y = __fmul_rn(__fmul_rn(o.y, sc), w.y);
          o.z = __fmul_rn(__fmul_rn(o.z, sc), w.z); o.w = __fmul_rn(__fmul_rn(o.w, sc), w.w);
        }
        x16_store_nf(xhi, xlo, xgs, f, o);
      }
    }
  }
  if (!routed) stage_route_hook(P, st, route, stage_index);
}
__device__ __noinline__ void stage_q8(const Program* Pp, const Stage* stp, const float* __restrict__ in, int n, const float* __restrict__ norm_w,
                                      int8_t* q_qs, float* q_d, short* q_bsums, int route, int stage_index) {
  extern __shared__ __align__(128) unsigned char dsk_dyn_smem[];
  const Program& P = *Pp; const Stage& st = *stp;
  float* red = reinterpret_cast<float*>(dsk_dyn_smem + 256);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nb = n >> 8;
  float sc = 1.0f;
  bool sc_known = norm_w == nullptr;
  if (norm_w && nb > 32) { sc = c_rms_scale(in, n, P.eps, red); sc_known = true; }
  bool routed = route < 0;
#pragma unroll 1
  for (int b0 = 0; b0 < nb; b0 += 32) {   // a warp per 256-block, four blocks per warp in flight
    float4 va[4], vb[4];
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int b = b0 + warp + 8 * k;
      va[k] = make_float4(0.f, 0.f, 0.f, 0.f); vb[k] = va[k];
      if (b < nb) {
        va[k] = *reinterpret_cast<const float4*>(in + (b << 8) + lane * 8);
        vb[k] = *reinterpret_cast<const float4*>(in + (b << 8) + lane * 8 + 4);
        ss = fmaf(va[k].x, va[k].x, ss); ss = fmaf(va[k].y, va[k].y, ss); ss = fmaf(va[k].z, va[k].z, ss); ss = fmaf(va[k].w, va[k].w, ss);
        ss = fmaf(vb[k].x, vb[k].x, ss); ss = fmaf(vb[k].y, vb[k].y, ss); ss = fmaf(vb[k].z, vb[k].z, ss); ss = fmaf(vb[k].w, vb[k].w, ss);
      }
    }
    if (!sc_known) { ss = csum(ss, red); sc = 1.0f / sqrtf(ss / (float)n + P.eps); sc_known = true; }
    if (!routed) { stage_route_hook(P, st, route, stage_index); routed = true; }
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int b = b0 + warp + 8 * k;
      if (b < nb) {
        float4 a = va[k], c = vb[k];
        if (norm_w) {
          const float4 w0 = *reinterpret_cast<const float4*>(norm_w + (b << 8) + lane * 8);
          const float4 w1 = *reinterpret_cast<const float4*>(norm_w + (b << 8) + lane * 8 + 4);
          a.x = __fmul_rn(__fmul_rn(a.x, sc), w0.x); a.y = __fmul_rn(__fmul_rn(a.y, sc), w0.y);
          a.z = __fmul_rn(__fmul_rn(a.z, sc), w0.z); a.w = __fmul_rn(__fmul_rn(a.w, sc), w0.w);
          c.x = __fmul_rn(__fmul_rn(c.x, sc), w1.x); c.y = __fmul_rn(__fmul_rn(c.y, sc), w1.y);
          c.z = __fmul_rn(__fmul_rn(c.z, sc), w1.z); c.w = __fmul_rn(__fmul_rn(c.w, sc), w1.w);
        }
        q8_block_nf(a, c, b, q_qs, q_d, q_bsums);
      }
    }
  }
  if (!routed) stage_route_hook(P, st, route, stage_index);
}
// flat activation layout of a DOWN stage: [routed: K*mi values][shared / dense: sh values]; per-segment views by offset
__host__ __device__ inline size_t down_x16_bytes(int K, int mi, int sh) { return (K * mi != 0 ? x16_bytes(K * mi) : 0) + (sh != 0 ? x16_bytes(sh) : 0); }
__device__ __forceinline__ void carve_down_x16(unsigned char* p, int K, int mi, int sh, X16* seg) {
  const X16 r = carve_x16(p, K * mi);
  for (int k = 0; k < K; k++) { seg[k].hi = r.hi + (uint32_t)(k * mi) * 2u; seg[k].lo = r.lo + (uint32_t)(k * mi) * 2u; seg[k].gs = r.gs + (uint32_t)(k * (mi >> 6)) * 4u; }
  seg[K] = carve_x16(p + (K * mi != 0 ? x16_bytes(K * mi) : 0), sh);
}
template <int Q>
__host__ __device__ inline size_t down_q8_bytes(int K, int mi, int sh) { return (K * mi != 0 ? xvec_bytes<Q>(K * mi) : 0) + (sh != 0 ? xvec_bytes<Q>(sh) : 0); }
template <int Q>
__device__ __forceinline__ void carve_down_q8(unsigned char* p, int K, int mi, int sh, Q8Smem* seg) {
  float* dummy = nullptr;
  Q8Smem r{};
  carve_x<Q>(p, K * mi, dummy, r);
  const int nbm = mi >> 8;
  for (int k = 0; k < K; k++) { seg[k].qs = r.qs + k * mi; seg[k].d = r.d + k * nbm; seg[k].bsums = r.bsums + k * nbm * 16; }
  carve_x<Q>(p + (K * mi != 0 ? xvec_bytes<Q>(K * mi) : 0), sh, dummy, seg[K]);
}

// ---- MoE gate logits of a quantised model (F32 weights, E rows): a dedicated compact stage --------------------------
// replaces: matmul_unscaled(s.moe_weights(), s.xb(), moegate->data, dim, n_routed_experts) on the rmsnorm'ed residual
// (src/infer.cpp:846-851: the gate is F32 in every quant, src/model.cpp:806-812)
// One row per tile (one 4n-byte TMA copy), one tile per CTA for E <= 148; all eight warps split the columns of the row and
// a block reduction in a fixed order finishes it.  Deliberately tiny: it replaces a whole second template instantiation of
// the generic consumer (cold code every layer) for 0.5 MB of weights.
__device__ __noinline__ int gate_f32_stage(int it, int dep_count, int stage_index) {
  extern __shared__ __align__(128) unsigned char dsk_dyn_smem[];
  const Program& P = *reinterpret_cast<const Program*>(dsk_dyn_smem + 4096 + 2 * kStageSlot);
  const Stage& st = *reinterpret_cast<const Stage*>(dsk_dyn_smem + 4096);
  const MegaSmem sm = carve_mega(dsk_dyn_smem, P.xregion_bytes);
  const int tid = threadIdx.x, n = st.n, nf = n >> 2;
  if ((int)blockIdx.x >= st.ntiles) {
    if (tid == 0) dep_signal(sm.dep, dep_count);
    return it;
  }
  float4* xs = reinterpret_cast<float4*>(sm.xregion);
  {  // x -> RMSNorm -> shared memory (fp32)
    float sc = 1.0f;
    if (st.norm_w && nf > 8 * kConsumers) sc = c_rms_scale(st.in, n, P.eps, sm.red);
#pragma unroll 1
    for (int c0 = 0; c0 < nf; c0 += 8 * kConsumers) {
      float4 v[8];
      float ss = 0.f;
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const int f = c0 + tid + k * kConsumers;
        v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (f < nf) {
          v[k] = reinterpret_cast<const float4*>(st.in)[f];
          ss = fmaf(v[k].x, v[k].x, ss); ss = fmaf(v[k].y, v[k].y, ss); ss = fmaf(v[k].z, v[k].z, ss); ss = fmaf(v[k].w, v[k].w, ss);
        }
      }
      if (st.norm_w && nf <= 8 * kConsumers) { ss = csum(ss, sm.red); sc = 1.0f / sqrtf(ss / (float)n + P.eps); }
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const int f = c0 + tid + k * kConsumers;
        if (f < nf) {
          float4 o = v[k];
          if (st.norm_w) {
            const float4 w = reinterpret_cast<const float4*>(st.norm_w)[f];
            o.x = __fmul_rn(__fmul_rn(o.x, sc), w.x); o.y = __fmul_rn(__fmul_rn(o.y, sc), w.y);
            o.z = __fmul_rn(__fmul_rn(o.z, sc), w.z); o.w = __fmul_rn(__fmul_rn(o.w, sc), w.w);
          }
          xs[f] = o;
        }
      }
    }
  }
  csync();
  if (tid == 0) dep_signal(sm.dep, dep_count);
  if (tid == 0 && blockIdx.x == 0 && P.tstamp) P.tstamp[stage_index * 8 + 1] = gtime();
  if (P.dbg_x != nullptr && blockIdx.x == 0) {   // test tap: the RMS-normalised fp32 vector the gate rows are multiplied with
    for (int i = tid; i < n; i += kConsumers) P.dbg_x[i] = reinterpret_cast<const float*>(xs)[i];
  }
  const MJob& jb = st.job[0];
#pragma unroll 1
  for (int t = blockIdx.x; t < st.ntiles; t += gridDim.x, it++) {
    const uint32_t wrow = ring_wait(sm, it) + (uint32_t)P.slot_scale;
    float acc = 0.f;
#pragma unroll 1
    for (int f = tid; f < nf; f += kConsumers) {
      const uint4 wv = lds128(wrow + (uint32_t)f * 16u);
      const float4 x = xs[f];
      acc = fmaf(__uint_as_float(wv.x), x.x, acc); acc = fmaf(__uint_as_float(wv.y), x.y, acc);
      acc = fmaf(__uint_as_float(wv.z), x.z, acc); acc = fmaf(__uint_as_float(wv.w), x.w, acc);
    }
    acc = csum(acc, sm.red);          // every warp is done with the slot after this reduction
    if (tid == 0) {
      ring_release(sm, it);
      if (t < jb.rows) jb.out[t] = acc;
    }
  }
  return it;
}

// ---- test taps (Program::dbg_q8 / dbg_x, null in production) ---------------------------------------------------------
// CTA 0 copies the staged activation vector of a GEMV stage to global memory exactly as the tile loop reads it:
// mode 0 = fp32 vector (generic path), 1 = Q8_K blocks as block_q8_K records {float d; int8 qs[256]; int16 bsums[16]}
// (src/quant.h:104-109), 2 = the fp16 hi/lo split of the tensor-core path, reconstructed as (hi + lo) * 2^-e.
template <int Q>
__device__ __noinline__ void dbg_tap(const Program* Pp, int n, int mode, const float* xs0, const int8_t* qs, const float* qd,
                                     const short* qb, uint32_t xhi, uint32_t xlo, uint32_t xgs) {
  const Program& P = *Pp;
  const int tid = threadIdx.x;
  if (mode == 1 && P.dbg_q8) {
    for (int b = 0; b < (n >> 8); b++) {
      unsigned char* o = P.dbg_q8 + (size_t)b * 292;
      if (tid == 0) { const float d = qd[b]; memcpy(o, &d, 4); }
      o[4 + tid] = (unsigned char)qs[b * 256 + tid];
      if (tid < 16) { const short v = qb[b * 16 + tid]; memcpy(o + 260 + 2 * tid, &v, 2); }
    }
  } else if (mode == 0 && P.dbg_x) {
    for (int i = tid; i < n; i += kConsumers) P.dbg_x[i] = xs0[(xswz<Q>(i >> 2) << 2) | (i & 3)];
  } else if (mode == 2 && P.dbg_x) {
    for (int i = tid; i < n; i += kConsumers) {
      unsigned short h, l;
      asm volatile("ld.shared.u16 %0, [%1];" : "=h"(h) : "r"(xhi + (uint32_t)i * 2u));
      asm volatile("ld.shared.u16 %0, [%1];" : "=h"(l) : "r"(xlo + (uint32_t)i * 2u));
      const float g = __uint_as_float(lds32(xgs + (uint32_t)(i >> 6) * 4u));
      P.dbg_x[i] = (h2f(h) + h2f(l)) * g;
    }
  }
}

template <int Q>
__device__ __forceinline__ void consumer_stage(const Program& P, const Stage& st, const MegaSmem& sm, int& it,
                                               unsigned long long& best_key, int dep_count, int stage_index) {
  constexpr bool KQ = QTraits<Q>::kq;
  const int tid = threadIdx.x;
  const int my_units = (st.kind == ST_DOWN && st.wp) ? (P.dim + st.down_rows - 1) / st.down_rows : st.ntiles;
  if ((int)blockIdx.x >= my_units) {    // no tile of this stage lands on this CTA: nothing to stage
    if (tid == 0) dep_signal(sm.dep, dep_count);
    return;
  }
  // routing first (one warp, registers): it unblocks the producer's routed-expert tiles
  if (st.kind == ST_DOWN && st.K > 0) {
    // the routing of this layer is normally still in this CTA's shared memory (left by route_all in the S56 stage of the same
    // launch, ordered by that stage's barriers); otherwise (first stage of a launch, CTA without S56 tiles) fetch the published copy
    if (*sm.marker != st.layer + 1) {
      if (tid < st.K) { sm.act[tid] = P.act[tid]; sm.actw[tid] = P.act_w[tid]; }
      csync();
    }
    if (tid == 0) dep_signal(sm.dep, dep_count);
    if (tid == 0 && blockIdx.x == 0 && P.tstamp) P.tstamp[stage_index * 8 + 4] = (gtime() - P.tstamp[stage_index * 8]) * 1000ull;
  }
  // activation vector(s) -> shared memory
  float* xs0 = nullptr;
  Q8Smem q80{};
  uint32_t xs_seg[kMaxJobs];
  Q8Smem q8_seg[kMaxJobs];
  X16 x16_0{};
  X16 x16_seg[kMaxJobs];
  const bool mma = Q == Q_F8 && st.use_mma;
  const int route = st.need_topk ? dep_count : -1;
  if (st.kind == ST_GEMV && mma) {
    x16_0 = carve_x16(sm.xregion, st.n);
    stage_x16(&P, &st, st.in, st.n, st.norm_w, x16_0.hi, x16_0.lo, x16_0.gs, route, stage_index);
  } else if (st.kind == ST_DOWN && mma) {
    const bool use_shared = st.sw2 != nullptr && st.add_shared;
    carve_down_x16(sm.xregion, st.K, st.mi, st.sh, x16_seg);
    if (st.K > 0) stage_x16(&P, &st, P.hbk, st.K * st.mi, nullptr, x16_seg[0].hi, x16_seg[0].lo, x16_seg[0].gs, -1, stage_index);
    if (use_shared && st.sh > 0) stage_x16(&P, &st, P.hbs, st.sh, nullptr, x16_seg[st.K].hi, x16_seg[st.K].lo, x16_seg[st.K].gs, -1, stage_index);
  } else if (st.kind == ST_GEMV && KQ) {
    carve_x<Q>(sm.xregion, st.n, xs0, q80);
    stage_q8(&P, &st, st.in, st.n, st.norm_w, q80.qs, q80.d, q80.bsums, route, stage_index);
  } else if (st.kind == ST_DOWN && KQ) {
    const bool use_shared = st.sw2 != nullptr && st.add_shared;
    carve_down_q8<Q>(sm.xregion, st.K, st.mi, st.sh, q8_seg);
    for (int k = 0; k <= st.K; k++) xs_seg[k] = 0u;
    if (st.K > 0) stage_q8(&P, &st, P.hbk, st.K * st.mi, nullptr, q8_seg[0].qs, q8_seg[0].d, q8_seg[0].bsums, -1, stage_index);
    if (use_shared && st.sh > 0) stage_q8(&P, &st, P.hbs, st.sh, nullptr, q8_seg[st.K].qs, q8_seg[st.K].d, q8_seg[st.K].bsums, -1, stage_index);
  } else if (KQ) {
    __trap();   // K-quant stages are always planned warp-per-tile (the planner rejects rows that do not fit a ring slot)
  } else if (st.kind == ST_GEMV) {
    carve_x<Q>(sm.xregion, st.n, xs0, q80);
    c_stage_gemv_input<Q>(P, st, sm, xs0, q80, dep_count, stage_index);
  } else {
    unsigned char* p = sm.xregion;
    const bool use_shared = st.sw2 != nullptr && st.add_shared;
    for (int k = 0; k <= st.K; k++) {
      const int n = k < st.K ? st.mi : st.sh;
      float* xk = nullptr;
      carve_x<Q>(p, n, xk, q8_seg[k]);
      xs_seg[k] = smem_u32(xk);
      p += n ? xvec_bytes<Q>(n) : 0;
      bool live = n > 0;
      if (k < st.K) { const int e = sm.act[k] - P.expert_first; live = live && e >= 0 && e < P.expert_count; }
      else live = live && use_shared;
      if (live) c_stage_vec<Q>(k < st.K ? P.hbk + (size_t)k * st.mi : P.hbs, n, nullptr, 1.0f, xk, q8_seg[k]);
    }
  }
  if (tid == 0 && blockIdx.x == 0 && P.tstamp && st.kind == ST_DOWN) P.tstamp[stage_index * 8 + 5] = (gtime() - P.tstamp[stage_index * 8]) * 1000ull;
  csync();
  if (!(st.need_topk || (st.kind == ST_DOWN && st.K > 0)) && tid == 0) dep_signal(sm.dep, dep_count);
  if (tid == 0 && blockIdx.x == 0 && P.tstamp) P.tstamp[stage_index * 8 + 1] = gtime();
  if ((P.dbg_q8 != nullptr || P.dbg_x != nullptr) && blockIdx.x == 0 && st.kind == ST_GEMV)
    dbg_tap<Q>(&P, st.n, mma ? 2 : (KQ ? 1 : 0), xs0, q80.qs, q80.d, q80.bsums, x16_0.hi, x16_0.lo, x16_0.gs);
  if (KQ && st.wp) {   // warp-per-tile K-quant stage
    const int lane = tid & 31;
    if (st.kind == ST_GEMV) {
      kq_gemv_loop<Q>(P, st, sm, q80, it, best_key, stage_index);
    } else {
      const int nrg = (P.dim + st.down_rows - 1) / st.down_rows;
      const int it0 = it, cnt = my_tile_count(nrg);
      long long c_wait = 0, c_task = 0;
      for (;;) {
        const int j = ring_claim(sm);
        if (j >= it0 + cnt) break;
        const long long k0 = clock64();
        const uint32_t slot = ring_wait(sm, j);
        const long long k1 = clock64();
        kq_down_group<Q>(P, st, sm, (int)blockIdx.x + (j - it0) * (int)gridDim.x, slot, q8_seg);
        __syncwarp();
        if (lane == 0) ring_release(sm, j);
        c_wait += k1 - k0; c_task += clock64() - k1;
      }
      it = it0 + cnt;
      if (lane == 0 && blockIdx.x == 0 && P.tstamp && c_task) {   // profiling: the busiest warp's cycles waiting for TMA / reducing
        atomicMax(&P.tstamp[stage_index * 8 + 6], (unsigned long long)c_wait);
        atomicMax(&P.tstamp[stage_index * 8 + 7], (unsigned long long)c_task);
      }
    }
    return;
  }
  if (mma && st.wp) {   // warp-per-tile tensor-core stage: warps claim tiles dynamically
    const int warp = tid >> 5, lane = tid & 31;
    if (st.kind == ST_GEMV) {
      long long c_wait = 0, c_task = 0, c_mma = 0;
      int n_mine = 0;
      const int it0 = it, cnt = my_tile_count(st.ntiles);
      for (;;) {
        const int j = ring_claim(sm);
        if (j >= it0 + cnt) break;
        const int t = (int)blockIdx.x + (j - it0) * (int)gridDim.x;
        const long long k0 = clock64();
        const uint32_t slot = ring_wait(sm, j);
        const long long k1 = clock64();
        wp_gemv_tile(P, st, t, slot, x16_0, sm.act, best_key, c_mma);
        __syncwarp();
        if (lane == 0) ring_release(sm, j);
        c_wait += k1 - k0; c_task += clock64() - k1; n_mine++;
      }
      it = it0 + cnt;
      if (lane == 0 && blockIdx.x == 0 && P.tstamp) {   // CTA 0: cycles per tile waiting for TMA / reducing, warp 0 (shares its
        if (warp == 0 && n_mine) {                        // sub-partition with the producer warp) and warp 1 (does not)
          P.tstamp[stage_index * 8 + 4] = (unsigned long long)(c_wait / n_mine); P.tstamp[stage_index * 8 + 5] = (unsigned long long)(c_task / n_mine);
        }
        if (warp == 1 && n_mine) {   // warp 1: cycles inside mma_rows_f8 / whole tile
          P.tstamp[stage_index * 8 + 6] = (unsigned long long)(c_mma / n_mine); P.tstamp[stage_index * 8 + 7] = (unsigned long long)(c_task / n_mine);
        }
      }
    } else {
      if (tid < 16) sm.sel[tid] = 0;
      csync();
      const int nrg = (P.dim + st.down_rows - 1) / st.down_rows;
      const int it0 = it, cnt = my_tile_count(nrg) * st.npieces;
      for (;;) {
        const int j = ring_claim(sm);
        if (j >= it0 + cnt) break;
        const int rgl = (j - it0) / st.npieces, pc = (j - it0) - rgl * st.npieces;
        const uint32_t slot = ring_wait(sm, j);
        wp_down_piece(P, st, sm, (int)blockIdx.x + rgl * (int)gridDim.x, rgl, pc, slot, x16_seg);
        __syncwarp();
        if (lane == 0) ring_release(sm, j);
      }
      it = it0 + cnt;
    }
    return;
  }
  if constexpr (KQ) return;   // (generic cooperative tiles below: F32 / F16 / F8 without the tensor-core plan)
  const uint32_t xs = KQ ? 0u : smem_u32(xs0);
  int parity_res = 0;
  long long c_wait = 0, c_task = 0, c_sync = 0, c_epi = 0;
  const bool timing = tid == 0 && blockIdx.x == 0 && P.tstamp;
  for (int t = blockIdx.x; t < st.ntiles; t += gridDim.x, it++) {
    const long long k0 = clock64();
    const uint32_t slot = ring_wait(sm, it);
    const long long k1 = clock64();
    float* res = sm.res + parity_res * 256;
    bool skip = false;
    // residual operand of the epilogue fetched now, so its L2 latency hides behind the tile's dot products
    float xres = 0.f;
    if (st.kind == ST_DOWN) {
      const int i = t * st.rows_per_tile + tid;
      if (tid < st.rows_per_tile && i < P.dim && !(P.partial != nullptr && (st.K > 0 || P.tp != 0))) xres = P.x[i];
    } else if (st.epi == EPI_RESID) {
      const int r = (t - st.job[0].tile_begin) * st.rows_per_tile + tid;
      if (tid < st.rows_per_tile && r < st.job[0].rows) xres = st.job[0].out[r];
    }
    if (st.kind == ST_GEMV) consume_gemv_tile<Q>(P, st, t, slot, xs, q80, x16_0, res, sm.act, best_key, skip);
    else if (mma) consume_down_tile_mma(P, st, t, slot, x16_seg, res, sm.act);
    else consume_down_tile<Q>(P, st, t, slot, xs_seg, q8_seg, res, sm.act);
    const long long k2 = clock64();
    csync();                       // every warp is done with the slot -> one arrival frees it
    if (tid == 0) ring_release(sm, it);
    const long long k3 = clock64();
    if (!skip) {
      if (st.kind == ST_GEMV) gemv_tile_epilogue(P, st, t, res, best_key, xres);
      else down_tile_epilogue(P, st, t, res, sm.actw, sm.act, xres);
    }
    parity_res ^= 1;
    c_wait += k1 - k0; c_task += k2 - k1; c_sync += k3 - k2; c_epi += clock64() - k3;
  }
  if (timing) {
    P.tstamp[stage_index * 8 + 4] = (unsigned long long)c_wait; P.tstamp[stage_index * 8 + 5] = (unsigned long long)c_task;
    P.tstamp[stage_index * 8 + 6] = (unsigned long long)c_sync; P.tstamp[stage_index * 8 + 7] = (unsigned long long)c_epi;
  }
}

// Producer loop of a GEMV stage with everything tile-invariant hoisted into registers: per tile it only advances the
// source pointers, waits for the slot, posts the byte count and issues the bulk copies.
template <int Q>
__device__ __forceinline__ void producer_gemv_fast(const Program& P, const Stage& st, const MegaSmem& sm, int& it, RingProd& rp,
                                                   int dep_count, bool& dep_waited) {
  const size_t rb = (st.quant == Q_F32 && Q != Q_F32) ? (size_t)st.n * 4 : QTraits<Q>::row_bytes(st.n);   // F32 gate rows in a quantised model
  const int RT = st.rows_per_tile, parts = st.epi == EPI_GLU ? 2 : 1;
  const uint32_t part_stride = (uint32_t)align_up((size_t)RT * rb, 128);
  const uint32_t full_bytes = (uint32_t)align_up((size_t)RT * rb, 16);
  const uint32_t slot_scale = (uint32_t)P.slot_scale;
  const uint32_t tile_need = (uint32_t)align_up((size_t)slot_scale + (size_t)parts * part_stride, 128);
  const int ncb = (st.n + P.bs1 - 1) / P.bs1, bs0 = P.bs0;
  // current job, cached
  int j = -1, j_begin = 0, j_end = 0, j_rows = 0;
  const uint8_t *w = nullptr, *wb = nullptr;
  const float *sc = nullptr, *scb = nullptr;
  bool j_dyn = false, j_live = true;
  for (int t = blockIdx.x; t < st.ntiles; t += gridDim.x, it++) {
    if (t >= j_end) {   // (re)load the job this tile belongs to
      do {
        j++;
        j_begin = st.job[j].tile_begin;
        j_end = (j + 1 < st.njobs) ? st.job[j + 1].tile_begin : st.ntiles;
      } while (t >= j_end);
      const MJob& jb = st.job[j];
      j_rows = jb.rows; w = jb.w; wb = jb.w_b; sc = jb.scale; scb = jb.scale_b;
      j_dyn = jb.expert_slot >= 0;
      j_live = true;
      if (j_dyn) {
        if (!dep_waited) { dep_wait(sm.dep, dep_count); dep_waited = true; }
        const int e = sm.act[jb.expert_slot] - P.expert_first;
        j_live = e >= 0 && e < P.expert_count;
        if (j_live) {
          w += (size_t)e * jb.w_stride;
          if (wb) wb += (size_t)e * jb.w_stride;
          if (sc) sc += (size_t)e * jb.s_stride;
          if (scb) scb += (size_t)e * jb.s_stride;
        }
      }
    }
    uint32_t full;
    const uint32_t slot = ring_acquire(P, sm, rp, it, j_live ? tile_need : 0u, st.max_inflight, full);
    if (!j_live) { mbar_expect_tx(full, 0); continue; }
    const int r0 = (t - j_begin) * RT;
    const int nrows = min(RT, j_rows - r0);
    const uint32_t bytes = nrows == RT ? full_bytes : (uint32_t)align_up((size_t)nrows * rb, 16);
    const size_t woff = (size_t)r0 * rb;
    uint32_t total = bytes * (uint32_t)parts, sb0 = 0, sb1 = 0, shift;
    const float *ss0 = nullptr, *ss1 = nullptr;
    if (sc) {
      const size_t so = (size_t)(r0 / bs0) * ncb;
      sb0 = scale_copy_bytes(sc + so, ncb, ss0, shift);
      if (parts == 2) sb1 = scale_copy_bytes(scb + so, ncb, ss1, shift);
      total += sb0 + sb1;
    }
    mbar_expect_tx(full, total);
    bulk_g2s(slot + slot_scale, w + woff, bytes, full);
    if (parts == 2) bulk_g2s(slot + slot_scale + part_stride, wb + woff, bytes, full);
    if (sc) {
      bulk_g2s(slot, ss0, sb0, full);
      if (parts == 2) bulk_g2s(slot + slot_scale / 2, ss1, sb1, full);
    }
  }
}

template <int Q>
__device__ __forceinline__ void producer_stage(const Program& P, const Stage& st, const MegaSmem& sm, int& it, RingProd& rp,
                                               int dep_count) {
  bool dep_waited = false;
  const bool dyn_all = st.kind == ST_DOWN && st.K > 0;
  if (st.kind == ST_DOWN && st.wp && QTraits<Q>::kq) {   // K-quants: one tile per row group (all segments)
    const int nrg = (P.dim + st.down_rows - 1) / st.down_rows;
    const uint32_t need = (uint32_t)align_up((size_t)st.K * (size_t)st.seg_stride + (size_t)st.down_rows * QTraits<Q>::row_bytes(st.sh), 128);
    for (int rg = blockIdx.x; rg < nrg; rg += gridDim.x, it++) {
      if (st.K > 0 && !dep_waited) { dep_wait(sm.dep, dep_count); dep_waited = true; }
      uint32_t full;
      const uint32_t slot = ring_acquire(P, sm, rp, it, need, st.max_inflight, full);
      kq_produce_down_group<Q>(P, st, rg, slot, full, sm.act);
    }
    if (!dep_waited) dep_wait(sm.dep, dep_count);
    return;
  }
  if (st.kind == ST_DOWN && st.wp) {
    const int nrg = (P.dim + st.down_rows - 1) / st.down_rows;
    for (int rg = blockIdx.x; rg < nrg; rg += gridDim.x) {
      for (int pc = 0; pc < st.npieces; pc++, it++) {
        const Piece pcd = st.piece[pc];
        if (pcd.seg < st.K && !dep_waited) { dep_wait(sm.dep, dep_count); dep_waited = true; }
        size_t rbp;
        if constexpr (QTraits<Q>::kq) rbp = QTraits<Q>::row_bytes(pcd.seg < st.K ? st.mi : st.sh);
        else rbp = f8_pitch((size_t)(pcd.seg < st.K ? st.mi : st.sh));
        const uint32_t need = (uint32_t)align_up((size_t)P.slot_scale + (size_t)pcd.g1 * rbp, 128);
        uint32_t full;
        const uint32_t slot = ring_acquire(P, sm, rp, it, need, st.max_inflight, full);
        if constexpr (!QTraits<Q>::kq) wp_produce_down_piece(P, st, rg, pc, slot, full, sm.act);
      }
    }
    if (!dep_waited) dep_wait(sm.dep, dep_count);
    return;
  }
  if (st.kind == ST_GEMV) {
    producer_gemv_fast<Q>(P, st, sm, it, rp, dep_count, dep_waited);
    if (!dep_waited) dep_wait(sm.dep, dep_count);
    return;
  }
  if constexpr (!QTraits<Q>::kq) {
  for (int t = blockIdx.x; t < st.ntiles; t += gridDim.x, it++) {
    bool dyn = dyn_all;
    if (dyn && !dep_waited) { dep_wait(sm.dep, dep_count); dep_waited = true; }
    uint32_t full;
    const uint32_t slot = ring_acquire(P, sm, rp, it, (uint32_t)align_up((size_t)P.slot_bytes, 128), st.max_inflight, full);
    produce_tile<Q>(P, st, t, slot, full, sm.act);
  }
  }
  if (!dep_waited) dep_wait(sm.dep, dep_count);   // bounds the run-ahead to one stage
}

// ---- the interpreter ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void copy_desc(void* dst, const void* src, int bytes, int t, int nthreads) {
  for (int i = t; i < bytes / 16; i += nthreads) reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(src)[i];
}

// grid barrier wait: acquire-spin on the monotonic arrival counter, bounded in TIME (a CTA that is not resident — the launch is
// cooperative, so that cannot happen silently — or a protocol bug traps instead of hanging the GPU)
__device__ __forceinline__ void grid_wait(const unsigned int* counter, unsigned int target) {
  unsigned long long t0 = 0ull;
  for (unsigned spins = 0; (int)(ld_acquire(counter) - target) < 0; spins++) {
    if ((spins & 4095u) == 4095u) spin_check(t0);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// True multi-head latent attention (BlockMLA::_attention_impl, src/infer.cpp:1051-1141; attn_mla 766-804): checkpoints
// converted with --mla carry wc (the k_nope_b^T . q_nope_b absorb), wq_rope_b and wv_b; the KV cache holds ONE latent row
// (kv_lora_rank fp16) + ONE rotated rope key (qk_rope_head_dim fp16) per token instead of n_heads full K / V rows.
//   ST_MLA_CACHE (CTA 0): kv_a -> rmsnorm(latent part) -> fp16 latent cache row kv_pos; rope(k_rope) -> fp16 rope cache row;
//                         re-rotation of the sink rows (src/infer.cpp:1084-1111).  Its own stage because the caches are
//                         shared by all heads: every head's CTA of the next stage only READS them.
//   ST_ATTN_MLA (CTA per head): rope(q_rope[h]); scores over the latent + rope caches; softmax; latent mix; then the head's
//                         value up-projection v_b[h] = wv_b[h] (v_head_dim x kv_lora_rank) . latent — matmul_expert with
//                         expert = head (src/infer.cpp:1133-1137) — with the weight quant's own arithmetic (Q8_K latent +
//                         integer dots for K-quants, exact fp16 split + mma for F8E5M2, fp32 otherwise).
// ------------------------------------------------------------------------------------------------------------------
__host__ __device__ inline size_t mla_floats(int kv_lora, int rope, int max_seq) {   // fp32 scratch: stage | qc | qr | out | att | part
  return (size_t)512 + (size_t)kv_lora + (size_t)((rope + 3) & ~3) + (size_t)kv_lora + (size_t)((max_seq + 3) & ~3) + 16;
}
template <int Q>
__host__ __device__ inline size_t mla_weight_bytes(int kv_lora, int vh) {   // the head's wv_b slab + staged latent, in shared memory
  if (QTraits<Q>::kq) return align_up((size_t)vh * QTraits<Q>::row_bytes(kv_lora), 128) + xvec_bytes<Q>(kv_lora);
  if (Q == Q_F8) return align_up((size_t)vh * f8_pitch((size_t)kv_lora), 128) + x16_bytes(kv_lora) + 1024;   // + the head's scale rows (<= 256 floats)
  return 0;   // F32 / F16: straight from global memory
}

__device__ __forceinline__ void c_mla_cache(const Program& P, const Stage& st, const MegaSmem& sm) {
  if (blockIdx.x != 0) return;
  const int tid = threadIdx.x;
  const Ctrl* c = P.ctrl;
  const int pos = c->pos, kv_pos = c->kv_pos, kv_sink = c->kv_sink;
  const int half_r = P.rope >> 1;
  float* stage = reinterpret_cast<float*>(sm.xregion);
  // latent part: rmsnorm(kv_a[0:kv_lora]) (src/infer.cpp:1079) -> fp16 row of the latent cache
  float ss = 0.f;
  for (int i = tid; i < P.kv_lora; i += kConsumers) { const float v = P.kv_a[i]; ss = fmaf(v, v, ss); }
  ss = csum(ss, sm.red);
  const float scale = 1.0f / sqrtf(ss / (float)P.kv_lora + P.eps);
  for (int i = tid; i < P.kv_lora; i += kConsumers) {
    const float v = __fmul_rn(__fmul_rn(P.kv_a[i], scale), st.norm_w[i]);
    P.kv_a[i] = v;                                                       // the reference normalises s.kv_a() in place
    st.kcache[(size_t)kv_pos * P.kv_lora + i] = __float2half_rn(v);
  }
  // rope part of the new key (one shared key per token) and the sink rows, exactly as the MHA stage does per head
  if (tid < half_r) {
    float cs, sn; rope_cs(P.rope_freq, tid, pos, cs, sn);
    const float v0 = P.kv_a[P.kv_lora + 2 * tid], v1 = P.kv_a[P.kv_lora + 2 * tid + 1];
    const float r0 = v0 * cs - v1 * sn, r1 = v0 * sn + v1 * cs;
    __half* kr = st.vcache + (size_t)kv_pos * P.rope;
    if (P.is_v3) { kr[2 * tid] = __float2half_rn(r0); kr[2 * tid + 1] = __float2half_rn(r1); }
    else { kr[tid] = __float2half_rn(r0); kr[tid + half_r] = __float2half_rn(r1); }
  } else if (tid >= 128 && tid < 128 + half_r * kv_sink && kv_sink > 0) {
    const int t = (tid - 128) % half_r, r = (tid - 128) / half_r;
    float cs, sn; rope_cs(P.rope_freq, t, 1, cs, sn);
    const __half* kr = st.vcache + (size_t)r * P.rope;
    const float v0 = __half2float(kr[2 * t]), v1 = __half2float(kr[2 * t + 1]);
    stage[2 * (r * half_r + t)] = v0 * cs - v1 * sn;
    stage[2 * (r * half_r + t) + 1] = v0 * sn + v1 * cs;
  }
  csync();
  if (tid >= 128 && tid < 128 + half_r * kv_sink && kv_sink > 0) {
    const int t = (tid - 128) % half_r, r = (tid - 128) / half_r;
    __half* kr = st.vcache + (size_t)r * P.rope;
    const float r0 = stage[2 * (r * half_r + t)], r1 = stage[2 * (r * half_r + t) + 1];
    if (P.is_v3) { kr[2 * t] = __float2half_rn(r0); kr[2 * t + 1] = __float2half_rn(r1); }
    else { kr[t] = __float2half_rn(r0); kr[t + half_r] = __float2half_rn(r1); }
  }
  csync();
}

template <int Q>
__device__ __noinline__ void c_attention_mla(const Program* Pp, const Stage* stp, int h) {
  extern __shared__ __align__(128) unsigned char dsk_dyn_smem[];
  const Program& P = *Pp; const Stage& st = *stp;
  const MegaSmem sm = carve_mega(dsk_dyn_smem, P.xregion_bytes);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const Ctrl* c = P.ctrl;
  const int pos = c->pos, kv_len = c->kv_len;
  const int L = P.kv_lora, R = P.rope, half_r = R >> 1;
  float* stage = reinterpret_cast<float*>(sm.xregion);                  // 512 floats (unused here; keeps the MHA layout)
  float* qc = stage + 512;
  float* qr = qc + L;
  float* out = qr + ((R + 3) & ~3);
  float* att = out + L;
  unsigned char* wbase = sm.xregion + align_up(mla_floats(L, R, P.max_seq) * 4, 128);
  const __half* ckv = st.kcache;    // (max_seq, kv_lora)
  const __half* krc = st.vcache;    // (max_seq, rope)
  float* qrh = P.q + (size_t)h * R;
  for (int i = tid; i < L; i += kConsumers) qc[i] = P.q_c[(size_t)h * L + i];
  if (tid < half_r) {
    float cs, sn; rope_cs(P.rope_freq, tid, pos, cs, sn);
    const float v0 = qrh[2 * tid], v1 = qrh[2 * tid + 1];
    const float r0 = v0 * cs - v1 * sn, r1 = v0 * sn + v1 * cs;
    if (P.is_v3) { qr[2 * tid] = r0; qr[2 * tid + 1] = r1; }
    else { qr[tid] = r0; qr[tid + half_r] = r1; }
  }
  csync();
  if (tid < R) qrh[tid] = qr[tid];                                       // the reference rotates s.q_rope(h) in place
  // scores: one warp per cached position
  const float inv = sqrtf((float)P.hd);
  for (int t = warp; t < kv_len; t += 8) {
    float s = 0.f;
    const __half* row = ckv + (size_t)t * L;
    for (int i = lane * 2; i < L; i += 64) {
      const float2 kk = __half22float2(*reinterpret_cast<const __half2*>(row + i));
      s = fmaf(qc[i], kk.x, s); s = fmaf(qc[i + 1], kk.y, s);
    }
    const __half* rr = krc + (size_t)t * R;
    for (int i = lane * 2; i < R; i += 64) {
      const float2 kk = __half22float2(*reinterpret_cast<const __half2*>(rr + i));
      s = fmaf(qr[i], kk.x, s); s = fmaf(qr[i + 1], kk.y, s);
    }
    s = warp_sum(s);
    if (lane == 0) att[t] = s / inv;
  }
  csync();
  float m = -3.402823466e38f;
  for (int t = tid; t < kv_len; t += kConsumers) m = fmaxf(m, att[t]);
  m = cmax(m, sm.red);
  float sum = 0.f;
  for (int t = tid; t < kv_len; t += kConsumers) { const float e = expf(att[t] - m); att[t] = e; sum += e; }
  sum = csum(sum, sm.red);
  for (int t = tid; t < kv_len; t += kConsumers) att[t] = att[t] / sum;
  csync();
  // latent mix: out[i] = sum_t att[t] * ckv[t][i]
  for (int i = tid; i < L; i += kConsumers) {
    float acc = 0.f;
    for (int t = 0; t < kv_len; t++) acc = fmaf(att[t], __half2float(ckv[(size_t)t * L + i]), acc);
    out[i] = acc;
    P.xb2[(size_t)h * L + i] = acc;
  }
  csync();
  // value up-projection of this head: kv_b[h*vh + r] = wv_b[h][r] . out
  const MJob& jb = st.job[0];
  const int vh = P.vh;
  float* vout = P.kv_b + (size_t)h * vh;
  const uint8_t* wsrc = jb.w + (size_t)h * (size_t)jb.w_stride;
  if constexpr (QTraits<Q>::kq) {
    const uint32_t rb = (uint32_t)QTraits<Q>::row_bytes(L);
    float* dummy = nullptr;
    Q8Smem q8{};
    carve_x<Q>(wbase + align_up((size_t)vh * rb, 128), L, dummy, q8);
    for (int b = warp; b < (L >> 8); b += 8) {
      const float4 va = *reinterpret_cast<const float4*>(out + (b << 8) + lane * 8);
      const float4 vb = *reinterpret_cast<const float4*>(out + (b << 8) + lane * 8 + 4);
      q8_block_nf(va, vb, b, q8.qs, q8.d, q8.bsums);
    }
    for (int i = tid; i < (int)((size_t)vh * rb / 16); i += kConsumers)
      reinterpret_cast<uint4*>(wbase)[i] = reinterpret_cast<const uint4*>(wsrc)[i];
    csync();
    for (int g = warp; g * 16 < vh; g += 8) {
      const int nrows = min(16, vh - g * 16);
      const float v = kq_tile_rows<Q>(smem_u32(wbase) + (uint32_t)(g * 16) * rb, rb, nrows, L >> 8, smem_u32(q8.qs), smem_u32(q8.d), smem_u32(q8.bsums));
      if (lane < nrows) vout[g * 16 + lane] = v;
    }
  } else if (Q == Q_F8 && st.use_mma) {
    const uint32_t pitch = (uint32_t)f8_pitch((size_t)L);
    unsigned char* xb = wbase + align_up((size_t)vh * pitch, 128);
    const X16 x16 = carve_x16(xb, L);
    float* srow = reinterpret_cast<float*>(xb + x16_bytes(L));           // the head's scale row (kv_lora / bs1 floats)
    if (tid < (L >> 2)) x16_store_nf(x16.hi, x16.lo, x16.gs, tid, *reinterpret_cast<const float4*>(out + 4 * tid));
    for (int f = kConsumers + tid; f < (L >> 2); f += kConsumers) x16_store_nf(x16.hi, x16.lo, x16.gs, f, *reinterpret_cast<const float4*>(out + 4 * f));
    for (int i = tid; i < (int)((size_t)vh * pitch / 16); i += kConsumers)
      reinterpret_cast<uint4*>(wbase)[i] = reinterpret_cast<const uint4*>(wsrc)[i];
    const int ncb = (L + P.bs1 - 1) / P.bs1;
    if (jb.scale && tid < (int)jb.s_stride) srow[tid] = jb.scale[(size_t)h * (size_t)jb.s_stride + tid];   // cdiv(vh, bs0) rows of ncb
    csync();
    const int gid = lane >> 2;
    for (int g = warp; g * 16 < vh; g += 8) {
      const int nrows = min(16, vh - g * 16);
      const uint32_t base = smem_u32(wbase) + (uint32_t)(g * 16) * pitch;
      const uint32_t a_lo = base + (uint32_t)min(gid, nrows - 1) * pitch;
      const uint32_t a_hi = nrows > 8 ? base + (uint32_t)min(gid + 8, nrows - 1) * pitch : 0u;
      const uint32_t ssm = jb.scale ? smem_u32(srow + (size_t)((g * 16) / P.bs0) * ncb) : 0u;   // bs0 % 16 == 0: one scale row per group
      const float2 vv = mma_rows_f8(a_lo, a_hi, ssm, ssm, P.bs1_shift, 0, L, x16.hi, x16.lo, x16.gs);
      if ((lane & 3) == 0) {
        if (gid < nrows) vout[g * 16 + gid] = vv.x;
        if (gid + 8 < nrows) vout[g * 16 + gid + 8] = vv.y;
      }
    }
  } else {   // F32 / F16 (and F8 without the tensor-core plan): a warp per row straight from global memory
    const size_t rb = QTraits<Q>::row_bytes(L);
    const int ncb = (L + P.bs1 - 1) / P.bs1;
    for (int r = warp; r < vh; r += 8) {
      const uint8_t* wr = wsrc + (size_t)r * rb;
      float acc = 0.f;
      if constexpr (Q == Q_F8) {
        const float* sc = jb.scale ? jb.scale + (size_t)h * (size_t)jb.s_stride + (size_t)(r / P.bs0) * ncb : nullptr;
        for (int b = 0; b < ncb; b++) {
          float p = 0.f;
          for (int i = b * P.bs1 + lane; i < min(L, (b + 1) * P.bs1); i += 32) p = fmaf(h2f((uint16_t)((uint16_t)wr[i] << 8)), out[i], p);
          acc = fmaf(p, sc ? sc[b] : 1.0f, acc);
        }
      } else if constexpr (Q == Q_F16) {
        for (int i = lane; i < L; i += 32) acc = fmaf(__half2float(reinterpret_cast<const __half*>(wr)[i]), out[i], acc);
      } else if constexpr (Q == Q_F32) {
        for (int i = lane; i < L; i += 32) acc = fmaf(reinterpret_cast<const float*>(wr)[i], out[i], acc);
      }
      acc = warp_sum(acc);
      if (lane == 0) vout[r] = acc;
    }
  }
  csync();
}

// Stages [s_begin, s_end) for n_tokens consecutive tokens in ONE launch (dsk_decode_greedy: the token loop lives inside the
// persistent kernel; tokens after the first always take their id from the on-device arg-max of the previous LM-head stage).
// Launched cooperatively with one CTA per SM: co-residency of the grid barrier's participants is guaranteed by the launch.
template <int Q>
__global__ void __launch_bounds__(kMegaThreads, 1) decode_kernel(const Program* __restrict__ prog, int s_begin, int s_end,
                                                                 int from_argmax, int n_tokens) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x;
  const bool producer = tid >= kConsumers;
  __shared__ int s_token;
  // header copy first: everything below reads the Program / Stage descriptors out of shared memory
  copy_desc(smem + 4096 + 2 * kStageSlot, prog, kProgHdrBytes, tid, kMegaThreads);
  __syncthreads();
  const Program& P = *reinterpret_cast<const Program*>(smem + 4096 + 2 * kStageSlot);
  const MegaSmem sm = carve_mega(smem, P.xregion_bytes);
  if (tid == 0) {
    for (int i = 0; i < kRingEntries; i++) {
      mbar_init(sm.full0 + 8u * i, 1); mbar_init(sm.empty0 + 8u * i, 1);
      sts32(sm.consumed + 4u * i, 0u); sts32(sm.ent_off + 4u * i, 0u); sts32(sm.start_abs + 4u * i, 0u);
    }
    sts32(sm.claim, 0u);
    for (int i = 0; i < 8; i++) reinterpret_cast<uint32_t*>(smem + kHdrZero)[i] = 0u;
    for (int i = 0; i < 4; i++) reinterpret_cast<float*>(smem + kHdrOne)[i] = 1.0f;
    dep_signal(sm.dep, 0);
    *sm.marker = 0;
    fence_proxy_async();
  }
  __syncthreads();
  const unsigned int base = *P.sync_base;
  const unsigned int G = gridDim.x;
  int it = 0;                      // tiles this CTA has pushed through the ring (same sequence on both sides)
  RingProd rp{0u, 0};              // producer thread only
  unsigned long long best_key = 0ull;
  int nstage_seen = 0;             // stages executed by this launch so far (all tokens): barrier targets and dep counts
#pragma unroll 1
  for (int tok = 0; tok < n_tokens; tok++) {
  const int feed = (tok > 0) ? 1 : from_argmax;
#pragma unroll 1
  for (int s = s_begin; s < s_end; s++, nstage_seen++) {
    if (producer) {
      // the producer keeps its own descriptor copy: it may already be one stage ahead of the consumers
      copy_desc(sm.st_p, &prog->stage[s], (int)sizeof(Stage), tid - kConsumers, 32);
      __syncwarp();
      const Stage& st = *sm.st_p;
      const bool streams = st.kind == ST_GEMV || st.kind == ST_DOWN;
      if (tid == kConsumers && streams) {
        producer_stage<Q>(P, st, sm, it, rp, nstage_seen + 1);   // (the F32 gate stage of a quantised model is a GEMV stage without scales)
      } else if (tid == kConsumers) {
        dep_wait(sm.dep, nstage_seen + 1);
      }
      __syncwarp();
      continue;
    }
    // ---- consumers ----
    // descriptor copy + norm-weight prefetch happen BEFORE the grid barrier: they do not depend on the previous stage
    copy_desc(sm.st_c, &prog->stage[s], (int)sizeof(Stage), tid, kConsumers);
    csync();
    const Stage& st = *sm.st_c;
    if (st.kind == ST_GEMV && st.norm_w && (int)blockIdx.x < st.ntiles) {
      for (int i = tid * 32; i < st.n; i += kConsumers * 32) asm volatile("prefetch.global.L1 [%0];" ::"l"(st.norm_w + i));
    }
    if (nstage_seen > 0) {         // grid barrier: every CTA has finished (and released) the previous stage
      if (tid == 0) grid_wait(P.sync_counter, base + (unsigned int)nstage_seen * G);
      csync();
    }
    // pull this stage's input vector(s) into L1 with every line in flight at once: the staging loops then take ONE L2 round
    // trip instead of one per batch of blocks (the acquire above has dropped any stale copy)
    if (st.kind == ST_GEMV) {
      if ((int)blockIdx.x < st.ntiles)
        for (int i = tid * 32; i < st.n; i += kConsumers * 32) asm volatile("prefetch.global.L1 [%0];" ::"l"(st.in + i));
    } else if (st.kind == ST_DOWN) {
      for (int i = tid * 32; i < st.K * st.mi; i += kConsumers * 32) asm volatile("prefetch.global.L1 [%0];" ::"l"(P.hbk + i));
      for (int i = tid * 32; i < st.sh; i += kConsumers * 32) asm volatile("prefetch.global.L1 [%0];" ::"l"(P.hbs + i));
    }
    if (tid == 0 && blockIdx.x == 0 && P.tstamp) { P.tstamp[s * 8 + 0] = gtime(); P.tstamp[s * 8 + 1] = 0; P.tstamp[s * 8 + 4] = 0; P.tstamp[s * 8 + 5] = 0; P.tstamp[s * 8 + 6] = 0; P.tstamp[s * 8 + 7] = 0; }
    if (st.kind == ST_EMBED) {
      if (blockIdx.x == 0) c_embed(P, feed, &s_token);
      if (tid == 0) dep_signal(sm.dep, nstage_seen + 1);
    } else if (st.kind == ST_XCHG) {
      c_xchg(P, st, sm);
      if (tid == 0) dep_signal(sm.dep, nstage_seen + 1);
    } else if (st.kind == ST_AMAX) {
      c_amax(P, st);
      if (tid == 0) dep_signal(sm.dep, nstage_seen + 1);
    } else if (st.kind == ST_ATTN) {
      for (int h = blockIdx.x; h < P.n_heads; h += gridDim.x) c_attention(P, st, sm, h);
      if (tid == 0) dep_signal(sm.dep, nstage_seen + 1);
    } else if (st.kind == ST_MLA_CACHE) {
      c_mla_cache(P, st, sm);
      if (tid == 0) dep_signal(sm.dep, nstage_seen + 1);
    } else if (st.kind == ST_ATTN_MLA) {
      for (int h = blockIdx.x; h < P.n_heads; h += gridDim.x) c_attention_mla<Q>(&P, &st, h);
      if (tid == 0) dep_signal(sm.dep, nstage_seen + 1);
    } else {
      if (st.quant == Q_F32 && Q != Q_F32) it = gate_f32_stage(it, nstage_seen + 1, s);
      else consumer_stage<Q>(P, st, sm, it, best_key, nstage_seen + 1, s);
      if (st.epi == EPI_LOGITS) {
        unsigned long long b = best_key;
#pragma unroll
        for (int o = 16; o; o >>= 1) { const unsigned long long ob = __shfl_xor_sync(0xffffffffu, b, o); if (ob > b) b = ob; }
        if ((tid & 31) == 0 && b) atomicMax(&P.ctrl->argmax_key, b);
        best_key = 0ull;
      }
    }
    // stage done: the CTA barrier orders every consumer's writes before thread 0's release-add (cumulative)
    if (tid == 0 && blockIdx.x == 0 && P.tstamp) P.tstamp[s * 8 + 2] = gtime();
    // peer-memory stores of this stage (partial sums / logits into the other ranks' exchange buffers): drain them with one
    // system-scope fence per thread, before this CTA's arrival on the grid barrier that precedes the exchange stage
    if (st.peer_stores) __threadfence_system();
    csync();
    if (tid == 0) sts32(sm.claim, (uint32_t)it);   // the claim counter overshoots by up to one ticket per warp at the end of a stage
    const bool last = (s + 1 == s_end) && (tok + 1 == n_tokens);
    if (!last && tid == 0) red_release_add(P.sync_counter, 1u);
    if (tid == 0 && blockIdx.x == 0 && P.tstamp) P.tstamp[s * 8 + 3] = gtime();
  }
  }
  // last stage of the launch: publish the new barrier base for the next launch (single writer, after all arrivals)
  if (!producer && blockIdx.x == 0 && tid == 0 && nstage_seen > 1) {
    const unsigned int target = base + (unsigned int)(nstage_seen - 1) * G;
    grid_wait(P.sync_counter, target);
    *P.sync_base = target;
  }
}

}  // namespace dsk
