/* oracle/dsk_oracle.c — plain-C restatement of the reference's single-token decode path.
 *
 * TEST INFRASTRUCTURE ONLY: loaded by tests/, __graft_entry__.smoke() and bench.py's CPU legs as the
 * checker.  The product never links it.  Pinned against oracle/_ref/libdsref.so (unmodified reference)
 * and the committed vectors under tests/golden (kat.json, ops.npz, e2e.npz, e2e_mla.npz) — see dsk_oracle.h.
 * MHA blocks and true-MLA blocks (use_mla).  Citations: /root/reference @ 8db9e56.
 */
#include "dsk_oracle.h"

#include <float.h>
#include <omp.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---- scalar conversions ------------------------------------------------------------------------ */

/* half_to_float: src/codec.h:22-24 (_cvtsh_ss). Exact IEEE binary16 -> binary32. */
float ork_half_to_float(uint16_t h) {
  uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
  uint32_t exp = (h >> 10) & 0x1Fu;
  uint32_t man = h & 0x3FFu;
  uint32_t bits;
  if (exp == 0) {
    if (man == 0) {
      bits = sign;
    } else { /* subnormal: normalise */
      int e = -1;
      do { man <<= 1; e++; } while (!(man & 0x400u));
      man &= 0x3FFu;
      bits = sign | ((uint32_t)(127 - 15 - e) << 23) | (man << 13);
    }
  } else if (exp == 31) {
    bits = sign | 0x7F800000u | (man << 13);
  } else {
    bits = sign | ((exp + 112u) << 23) | (man << 13);
  }
  float f;
  memcpy(&f, &bits, 4);
  return f;
}

/* float_to_half: src/codec.h:25-27 (_cvtss_sh(x, 0) = round to nearest even). */
uint16_t ork_float_to_half(float f) {
  uint32_t x;
  memcpy(&x, &f, 4);
  uint32_t sign = (x >> 16) & 0x8000u;
  uint32_t ax = x & 0x7FFFFFFFu;
  if (ax >= 0x7F800000u) { /* inf / nan */
    return (uint16_t)(sign | 0x7C00u | ((ax > 0x7F800000u) ? (0x200u | ((ax >> 13) & 0x3FFu)) : 0));
  }
  if (ax >= 0x477FF000u) { /* rounds to >= 65520 -> inf */
    return (uint16_t)(sign | 0x7C00u);
  }
  if (ax < 0x33000001u) { /* < 2^-25 (or == 2^-25, ties to even 0) */
    return (uint16_t)sign;
  }
  int32_t e = (int32_t)(ax >> 23) - 127;
  uint32_t man = (ax & 0x7FFFFFu) | 0x800000u;
  int shift;
  uint32_t hexp;
  if (e < -14) { shift = 13 + (-14 - e); hexp = 0; } else { shift = 13; hexp = (uint32_t)(e + 15); }
  uint32_t r = man >> shift;
  uint32_t rem = man & ((1u << shift) - 1u);
  uint32_t half = 1u << (shift - 1);
  if (rem > half || (rem == half && (r & 1u))) r++;
  uint32_t out;
  if (hexp == 0) out = r;               /* subnormal; carry into exponent is naturally correct */
  else out = ((hexp - 1) << 10) + r;    /* r has the implicit bit at 0x400 */
  return (uint16_t)(sign | out);
}

/* float8e5m2_to_float: src/codec.h:40-48 — the byte is the top 8 bits of an fp16. */
float ork_f8e5m2_to_float(uint8_t b) { return ork_half_to_float((uint16_t)((uint16_t)b << 8)); }

/* nearest_int: src/quant.cpp:34-39 */
static inline int nearest_int(float fval) {
  float val = fval + 12582912.f;
  int i;
  memcpy(&i, &val, sizeof(int));
  return (i & 0x007fffff) - 0x00400000;
}

/* ---- quant.cpp --------------------------------------------------------------------------------- */

/* quantize_row_q8_K_ref: src/quant.cpp:616-653.
 * Bit-exactness note: the reference is built with -ffast-math (Makefile:31); its object code computes
 * iscale = -127.f/max as a true division but folds `d = 1/iscale` into `max * (-1/127.f)` (one vmulss by a
 * constant) — measured: 3000/3000 random blocks match that form, 2171/3000 match the literal 1/iscale.
 * The restatement (and the CUDA kernel) follow the compiled form; tests/test_oracle.py::test_q8k_bit_exact
 * pins it byte-for-byte against libdsref.so. */
void ork_quantize_row_q8_K(const float* x, ork_block_q8_K* y, long k) {
  const long nb = k / ORK_QK_K;
  for (long i = 0; i < nb; i++) {
    float max = 0, amax = 0;
    for (int j = 0; j < ORK_QK_K; ++j) {
      float ax = fabsf(x[j]);
      if (ax > amax) { amax = ax; max = x[j]; }
    }
    if (!amax) {
      y[i].d = 0;
      memset(y[i].qs, 0, ORK_QK_K);
      /* NB the reference leaves bsums untouched here (quant.cpp:630-635); zero them for determinism */
      memset(y[i].bsums, 0, sizeof(y[i].bsums));
      x += ORK_QK_K;
      continue;
    }
    const float iscale = -127.f / max;
    for (int j = 0; j < ORK_QK_K; ++j) {
      int v = nearest_int(iscale * x[j]);
      y[i].qs[j] = (int8_t)(v < 127 ? v : 127);
    }
    for (int j = 0; j < ORK_QK_K / 16; ++j) {
      int sum = 0;
      for (int ii = 0; ii < 16; ++ii) sum += y[i].qs[j * 16 + ii];
      y[i].bsums[j] = (int16_t)sum;
    }
    y[i].d = max * (-1.0f / 127.0f);
    x += ORK_QK_K;
  }
}

/* dequantize_row_q2_K: src/quant.cpp:217-247 */
void ork_dequantize_row_q2_K(const ork_block_q2_K* x, float* y, long k) {
  const long nb = k / ORK_QK_K;
  for (long i = 0; i < nb; i++) {
    const float d = ork_half_to_float(x[i].d);
    const float min = ork_half_to_float(x[i].dmin);
    const uint8_t* q = x[i].qs;
    int is = 0;
    for (int n = 0; n < ORK_QK_K; n += 128) {
      int shift = 0;
      for (int j = 0; j < 4; ++j) {
        uint8_t sc = x[i].scales[is++];
        float dl = d * (sc & 0xF), ml = min * (sc >> 4);
        for (int l = 0; l < 16; ++l) *y++ = dl * ((int8_t)((q[l] >> shift) & 3)) - ml;
        sc = x[i].scales[is++];
        dl = d * (sc & 0xF); ml = min * (sc >> 4);
        for (int l = 0; l < 16; ++l) *y++ = dl * ((int8_t)((q[l + 16] >> shift) & 3)) - ml;
        shift += 2;
      }
      q += 32;
    }
  }
}

/* 6-bit scale unpack shared by Q3_K paths: src/quant.cpp:402-407 */
static void q3_unpack_scales(const uint8_t* packed, int8_t scales[16]) {
  const uint32_t kmask1 = 0x03030303, kmask2 = 0x0f0f0f0f;
  uint32_t aux[4];
  memcpy(aux, packed, 12);
  uint32_t tmp = aux[2];
  aux[2] = ((aux[0] >> 4) & kmask2) | (((tmp >> 4) & kmask1) << 4);
  aux[3] = ((aux[1] >> 4) & kmask2) | (((tmp >> 6) & kmask1) << 4);
  aux[0] = (aux[0] & kmask2) | (((tmp >> 0) & kmask1) << 4);
  aux[1] = (aux[1] & kmask2) | (((tmp >> 2) & kmask1) << 4);
  memcpy(scales, aux, 16);
}

/* dequantize_row_q3_K: src/quant.cpp:384-432 */
void ork_dequantize_row_q3_K(const ork_block_q3_K* x, float* y, long k) {
  const long nb = k / ORK_QK_K;
  int8_t scales[16];
  for (long i = 0; i < nb; i++) {
    const float d_all = ork_half_to_float(x[i].d);
    const uint8_t* q = x[i].qs;
    const uint8_t* hm = x[i].hmask;
    uint8_t m = 1;
    q3_unpack_scales(x[i].scales, scales);
    int is = 0;
    for (int n = 0; n < ORK_QK_K; n += 128) {
      int shift = 0;
      for (int j = 0; j < 4; ++j) {
        float dl = d_all * (scales[is++] - 32);
        for (int l = 0; l < 16; ++l) *y++ = dl * ((int8_t)((q[l + 0] >> shift) & 3) - ((hm[l + 0] & m) ? 0 : 4));
        dl = d_all * (scales[is++] - 32);
        for (int l = 0; l < 16; ++l) *y++ = dl * ((int8_t)((q[l + 16] >> shift) & 3) - ((hm[l + 16] & m) ? 0 : 4));
        shift += 2;
        m <<= 1;
      }
      q += 32;
    }
  }
}

/* ggml_vec_dot_q2_K_q8_K: src/quant.cpp:666-783.  Integer part restated exactly (scalar branch 744-781);
 * the fp32 accumulation across blocks uses the scalar branch's order (the AVX2 branch keeps 8 lanes and
 * hsums at the end) — agreement with the reference is therefore to fp32 re-association (~1e-6 rel). */
float ork_vec_dot_q2_K_q8_K(int n, const ork_block_q2_K* x, const ork_block_q8_K* y) {
  const int nb = n / ORK_QK_K;
  float sumf = 0;
  for (int i = 0; i < nb; ++i) {
    const uint8_t* q2 = x[i].qs;
    const int8_t* q8 = y[i].qs;
    const uint8_t* sc = x[i].scales;
    int summs = 0;
    for (int j = 0; j < 16; ++j) summs += y[i].bsums[j] * (sc[j] >> 4);
    const float dall = y[i].d * ork_half_to_float(x[i].d);
    const float dmin = y[i].d * ork_half_to_float(x[i].dmin);
    int isum = 0, is = 0;
    for (int k = 0; k < ORK_QK_K / 128; ++k) {
      int shift = 0;
      for (int j = 0; j < 4; ++j) {
        int d = sc[is++] & 0xF;
        int isuml = 0;
        for (int l = 0; l < 16; ++l) isuml += q8[l] * ((q2[l] >> shift) & 3);
        isum += d * isuml;
        d = sc[is++] & 0xF;
        isuml = 0;
        for (int l = 16; l < 32; ++l) isuml += q8[l] * ((q2[l] >> shift) & 3);
        isum += d * isuml;
        shift += 2;
        q8 += 32;
      }
      q2 += 32;
    }
    sumf += dall * isum - dmin * summs;
  }
  return sumf;
}

/* ggml_vec_dot_q3_K_q8_K: src/quant.cpp:434-614 (scalar branch 558-610 for the integer part). */
float ork_vec_dot_q3_K_q8_K(int n, const ork_block_q3_K* x, const ork_block_q8_K* y) {
  const int nb = n / ORK_QK_K;
  int8_t scales[16];
  float sumf = 0;
  for (int i = 0; i < nb; ++i) {
    const uint8_t* q3 = x[i].qs;
    const uint8_t* hm = x[i].hmask;
    const int8_t* q8 = y[i].qs;
    q3_unpack_scales(x[i].scales, scales);
    int isum = 0, is = 0;
    uint8_t m = 1;
    for (int k = 0; k < ORK_QK_K / 128; ++k) {
      int shift = 0;
      for (int j = 0; j < 4; ++j) {
        int isuml = 0;
        for (int l = 0; l < 16; ++l) isuml += q8[l] * ((int)((q3[l] >> shift) & 3) - ((hm[l] & m) ? 0 : 4));
        isum += (scales[is++] - 32) * isuml;
        isuml = 0;
        for (int l = 16; l < 32; ++l) isuml += q8[l] * ((int)((q3[l] >> shift) & 3) - ((hm[l] & m) ? 0 : 4));
        isum += (scales[is++] - 32) * isuml;
        shift += 2;
        m <<= 1;
        q8 += 32;
      }
      q3 += 32;
    }
    const float d = ork_half_to_float(x[i].d) * y[i].d;
    sumf += d * isum;
  }
  return sumf;
}

/* ---- infer.cpp: GEMVs -------------------------------------------------------------------------- */

static int cdiv(int a, int b) { return (a + b - 1) / b; }

/* hsum of the two 8-lane accumulators exactly as src/infer.cpp:221-227 / 301-307:
 * sum8 = lo + hi; sum4 = sum8[0:4] + sum8[4:8]; _mm_dp_ps(sum4, 1, 0xf1) = (s0+s1)+(s2+s3). */
static float hsum16(const float* lo, const float* hi) {
  float s8[8], s4[4];
  for (int l = 0; l < 8; l++) s8[l] = lo[l] + hi[l];
  for (int l = 0; l < 4; l++) s4[l] = s8[l] + s8[l + 4];
  return (s4[0] + s4[1]) + (s4[2] + s4[3]);
}

/* _matmul(float*): src/infer.cpp:121-157 */
static void matmul_f32(float* xout, const float* x, const float* w, int n, int d, int bs0, int bs1,
                       const float* scale) {
  static const float one = 1.0f;
  if (!scale) { scale = &one; bs0 = d; bs1 = n; }
  int scale_num_cols = cdiv(n, bs1);
#pragma omp parallel for
  for (int i = 0; i < d; i++) {
    int scale_i = i / bs0;
    float val = 0.0f;
    for (int scale_j = 0; scale_j < scale_num_cols; scale_j++) {
      float scale_val = scale[scale_i * scale_num_cols + scale_j];
      for (int jj = 0; jj < bs1; jj++) {
        int j = scale_j * bs1 + jj;
        if (j >= n) break;
        val += (w[(size_t)i * n + j] * x[j]) * scale_val;
      }
    }
    xout[i] = val;
  }
}

/* _matmul(f16_t*): src/infer.cpp:161-233 and _matmul(f8e5m2_t*): 238-313 share one 16-lane structure:
 * w_ps = cvt(w) * scale ; sum = fma(w_ps, x, sum) in lanes j%16 ; hsum16. */
static void matmul_f16_f8(float* xout, const float* x, const void* w, int is_f8, int n, int d, int bs0, int bs1,
                          const float* scale) {
  static const float one = 1.0f;
  if (!scale) { scale = &one; bs0 = d; bs1 = n; }
  int scale_num_cols = cdiv(n, bs1);
#pragma omp parallel for
  for (int i = 0; i < d; i++) {
    int scale_i = i / bs0;
    float lo[8] = {0}, hi[8] = {0};
    for (int scale_j = 0; scale_j < scale_num_cols; scale_j++) {
      float scale_val = scale[scale_i * scale_num_cols + scale_j];
      for (int jj = 0; jj < bs1; jj += 16) {
        int j = scale_j * bs1 + jj;
        if (j >= n) break;
        for (int l = 0; l < 16; l++) {
          size_t idx = (size_t)i * n + j + l;
          float wv = is_f8 ? ork_f8e5m2_to_float(((const uint8_t*)w)[idx]) : ork_half_to_float(((const uint16_t*)w)[idx]);
          wv = wv * scale_val;
          float* acc = l < 8 ? &lo[l] : &hi[l - 8];
          *acc = fmaf(wv, x[j + l], *acc);
        }
      }
    }
    xout[i] = hsum16(lo, hi);
  }
}

/* _matmul(block_q2_K*) / (block_q3_K*): src/infer.cpp:315-379 — quantise acts to Q8_K, integer dot per row. */
static void matmul_kquant(float* xout, const float* x, const void* w, int q3, int n, int d) {
  int nb = n / ORK_QK_K;
  ork_block_q8_K* aq = (ork_block_q8_K*)malloc(sizeof(ork_block_q8_K) * (size_t)nb);
  ork_quantize_row_q8_K(x, aq, (long)nb * ORK_QK_K);
#pragma omp parallel for
  for (int i = 0; i < d; i++) {
    if (q3) xout[i] = ork_vec_dot_q3_K_q8_K(n, (const ork_block_q3_K*)w + (size_t)i * nb, aq);
    else    xout[i] = ork_vec_dot_q2_K_q8_K(n, (const ork_block_q2_K*)w + (size_t)i * nb, aq);
  }
  free(aq);
}

/* matmul src/infer.cpp:381-417 and matmul_expert 423-469 (expert < 0 or n_experts == 0 -> plain). */
void ork_matmul(float* xout, const float* x, const ork_tensor* w, int expert, int bs0, int bs1) {
  int n = w->cols, d = w->rows;
  size_t eidx = (w->n_experts > 0 && expert >= 0) ? (size_t)expert : 0;
  size_t elem_off = eidx * (size_t)d * (size_t)n;
  const float* scale = w->scale;
  if (scale && bs0 > 0) scale += eidx * (size_t)cdiv(d, bs0) * (size_t)cdiv(n, bs1);
  switch (w->quant) {
    case ORK_F32: matmul_f32(xout, x, (const float*)w->data + elem_off, n, d, bs0, bs1, scale); break;
    case ORK_F16: matmul_f16_f8(xout, x, (const uint16_t*)w->data + elem_off, 0, n, d, bs0, bs1, scale); break;
    case ORK_F8E5M2: matmul_f16_f8(xout, x, (const uint8_t*)w->data + elem_off, 1, n, d, bs0, bs1, scale); break;
    case ORK_Q2_K: matmul_kquant(xout, x, (const ork_block_q2_K*)w->data + elem_off / ORK_QK_K, 0, n, d); break;
    case ORK_Q3_K: matmul_kquant(xout, x, (const ork_block_q3_K*)w->data + elem_off / ORK_QK_K, 1, n, d); break;
    default: abort();
  }
}

/* ---- infer.cpp: small ops ---------------------------------------------------------------------- */

/* softmax: src/infer.cpp:472-487 */
void ork_softmax(float* o, const float* x, int size) {
  float score_max = -FLT_MAX;
  for (int i = 0; i < size; ++i) if (x[i] > score_max) score_max = x[i];
  float score_sum = 0.0f;
  for (int i = 0; i < size; ++i) { o[i] = expf(x[i] - score_max); score_sum += o[i]; }
  for (int i = 0; i < size; ++i) o[i] /= score_sum;
}

static float sigmoidf(float x) { return 1.0f / (1.0f + expf(-x)); } /* src/infer.cpp:489-491 */

/* moe_gate: src/infer.cpp:493-599.
 * GROUP_LIMITED_GREEDY first pass (551-566): `best` starts at -1 and the test is `x[j] > x[best]`, i.e. it reads
 * x[-1] (UB).  For the reference's `new float[]` buffer that word is 0.0f, so the rule is
 * "the first unmasked j is accepted iff x[j] > 0".  Restated with an explicit 0.0f; if no candidate is found
 * (all remaining scores <= 0) nothing is marked for that (group, k) — the reference would index mask[-1]. */
void ork_moe_gate(float* moe_weights, const float* bias, int* active_experts, float* x, int n_routed, int n_active,
                  int norm_topk_prob, float routed_scaling_factor, int scoring_sigmoid, int topk_method, int n_group,
                  int topk_group) {
  if (scoring_sigmoid) { for (int i = 0; i < n_routed; i++) x[i] = sigmoidf(x[i]); }
  else ork_softmax(x, x, n_routed);
  if (bias) for (int i = 0; i < n_routed; ++i) x[i] += bias[i];
  uint8_t mask[32];
  memset(mask, 0, sizeof(mask));
  float wsum = 0.0f;
  if (topk_method == 1) {
    int group_size = n_routed / n_group;
    for (int g = 0; g < n_group; g++) {
      for (int k = 0; k < topk_group; k++) {
        int best = -1;
        for (int j = g * group_size; j < (g + 1) * group_size; j++) {
          float xb = best < 0 ? 0.0f : x[best];
          if ((mask[j / 8] & (1u << (j % 8))) == 0 && x[j] > xb) best = j;
        }
        if (best >= 0) mask[best / 8] |= (uint8_t)(1u << (best % 8));
      }
    }
    for (int i = 0; i < 32; i++) mask[i] = (uint8_t)~mask[i];
  }
  for (int k = 0; k < n_active; ++k) {
    int best = -1;
    for (int j = 0; j < n_routed; ++j) {
      if ((mask[j / 8] & (1u << (j % 8))) == 0 && (best == -1 || x[j] > x[best])) best = j;
    }
    active_experts[k] = best;
    wsum += x[best];
    mask[best / 8] |= (uint8_t)(1u << (best % 8));
  }
  if (!norm_topk_prob) wsum = 1.0f;
  for (int k = 0; k < n_active; ++k) moe_weights[k] = x[active_experts[k]] / wsum * routed_scaling_factor;
}

/* rmsnorm: src/infer.cpp:601-611 */
void ork_rmsnorm(float* o, const float* x, const float* w, int size, float eps) {
  float rms = 0.0f;
  for (int i = 0; i < size; ++i) rms += x[i] * x[i];
  rms = sqrtf(rms / size + eps);
  float scale = 1.0f / rms;
  for (int i = 0; i < size; ++i) o[i] = x[i] * scale * w[i];
}

/* rope (V2, de-interleaving): src/infer.cpp:648-668 */
void ork_rope(float* vec, int d, int head_dim, int pos, float theta) {
  float buf[1024];
  for (int i = 0; i < d; i += 2) {
    int j_head = i % head_dim;
    float freq = 1.0f / powf(theta, (float)j_head / (float)head_dim);
    float val = pos * freq;
    float fcr = cosf(val), fci = sinf(val);
    float v0 = vec[i], v1 = vec[i + 1];
    buf[i / 2] = v0 * fcr - v1 * fci;
    buf[i / 2 + d / 2] = v0 * fci + v1 * fcr;
  }
  for (int i = 0; i < d; i++) vec[i] = buf[i];
}

/* rope_v3 (interleaved, in place): src/infer.cpp:670-685 */
void ork_rope_v3(float* vec, int d, int head_dim, int pos, float theta) {
  for (int i = 0; i < d; i += 2) {
    int j_head = i % head_dim;
    float freq = 1.0f / powf(theta, (float)j_head / (float)head_dim);
    float val = pos * freq;
    float fcr = cosf(val), fci = sinf(val);
    float v0 = vec[i], v1 = vec[i + 1];
    vec[i] = v0 * fcr - v1 * fci;
    vec[i + 1] = v0 * fci + v1 * fcr;
  }
}

/* fp16 in-place variants for sink re-rotation: src/infer.cpp:687-707, 709-724 */
void ork_rope_f16(uint16_t* vec, int d, int head_dim, int pos, float theta) {
  float buf[1024];
  for (int i = 0; i < d; i++) buf[i] = ork_half_to_float(vec[i]);
  ork_rope(buf, d, head_dim, pos, theta);
  for (int i = 0; i < d; i++) vec[i] = ork_float_to_half(buf[i]);
}
void ork_rope_v3_f16(uint16_t* vec, int d, int head_dim, int pos, float theta) {
  float buf[1024];
  for (int i = 0; i < d; i++) buf[i] = ork_half_to_float(vec[i]);
  ork_rope_v3(buf, d, head_dim, pos, theta);
  for (int i = 0; i < d; i++) vec[i] = ork_float_to_half(buf[i]);
}

float ork_silu(float x) { return x / (1.0f + expf(-x)); } /* src/infer.cpp:640-642 */
float ork_gelu(float x) { return 0.5f * x * (1.0f + tanhf(0.797885f * (x + 0.044715f * x * x * x))); } /* 636-638 */

/* attn: src/infer.cpp:728-762 */
void ork_attn(float* xout, float* atth, const float* qh, const uint16_t* kh, const uint16_t* vh, int head_dim,
              int v_head_dim, int n_heads, int kv_len) {
  int k_stride = n_heads * head_dim;
  for (int t = 0; t < kv_len; ++t) {
    float score = 0.0f;
    for (int i = 0; i < head_dim; ++i) score += qh[i] * ork_half_to_float(kh[(size_t)t * k_stride + i]);
    score /= sqrtf((float)head_dim);
    atth[t] = score;
  }
  ork_softmax(atth, atth, kv_len);
  int v_stride = n_heads * v_head_dim;
  for (int i = 0; i < v_head_dim; ++i) {
    float vi = 0.0f;
    for (int t = 0; t < kv_len; ++t) vi += atth[t] * ork_half_to_float(vh[(size_t)t * v_stride + i]);
    xout[i] = vi;
  }
}

/* attn_mla: src/infer.cpp:766-804 — scores over the latent + rope caches, softmax, latent mix */
void ork_attn_mla(float* xout, float* atth, const float* qh_c, const float* qh_rope, const uint16_t* ckv, const uint16_t* k_rope,
                  int head_dim, int kv_lora_rank, int qk_rope_head_dim, int kv_len) {
  for (int t = 0; t < kv_len; ++t) {
    float score = 0.0f;
    for (int i = 0; i < kv_lora_rank; ++i) score += qh_c[i] * ork_half_to_float(ckv[(size_t)t * kv_lora_rank + i]);
    for (int i = 0; i < qk_rope_head_dim; ++i) score += qh_rope[i] * ork_half_to_float(k_rope[(size_t)t * qk_rope_head_dim + i]);
    score /= sqrtf((float)head_dim);
    atth[t] = score;
  }
  ork_softmax(atth, atth, kv_len);
  for (int i = 0; i < kv_lora_rank; ++i) {
    float vi = 0.0f;
    for (int t = 0; t < kv_len; ++t) vi += atth[t] * ork_half_to_float(ckv[(size_t)t * kv_lora_rank + i]);
    xout[i] = vi;
  }
}

/* ---- model level ------------------------------------------------------------------------------- */

/* Model::_copy_embedding: src/infer.cpp:1217-1263 */
void ork_copy_embedding(const ork_model* m, ork_state* s, int token) {
  const ork_config* c = &m->cfg;
  const ork_tensor* e = &m->embed;
  switch (c->quant) {
    case ORK_F32: for (int i = 0; i < c->dim; ++i) s->x[i] = ((const float*)e->data)[(size_t)token * c->dim + i]; break;
    case ORK_F16: for (int i = 0; i < c->dim; ++i) s->x[i] = ork_half_to_float(((const uint16_t*)e->data)[(size_t)token * c->dim + i]); break;
    case ORK_F8E5M2: {
      int scale_num_cols = cdiv(c->dim, c->bs1);
      for (int i = 0; i < c->dim; ++i) {
        float scale = e->scale[(token / c->bs0) * scale_num_cols + i / c->bs1];
        s->x[i] = ork_f8e5m2_to_float(((const uint8_t*)e->data)[(size_t)token * c->dim + i]) * scale;
      }
      break;
    }
    case ORK_Q2_K: ork_dequantize_row_q2_K((const ork_block_q2_K*)e->data + (size_t)token * (c->dim / ORK_QK_K), s->x, c->dim); break;
    case ORK_Q3_K: ork_dequantize_row_q3_K((const ork_block_q3_K*)e->data + (size_t)token * (c->dim / ORK_QK_K), s->x, c->dim); break;
    default: abort();
  }
}

static float act_fn(const ork_config* c, float v) { return c->act_silu ? ork_silu(v) : ork_gelu(v); }

/* BlockMHA::_attention_impl: src/infer.cpp:934-1049 */
static void attention_mha(const ork_model* m, ork_state* s, const ork_layer* L, int pos, int kv_sink, int kv_pos,
                          int kv_len) {
  const ork_config* c = &m->cfg;
  if (c->q_lora_rank > 0) {
    ork_matmul(s->q_a, s->xb, &L->wq_a, -1, c->bs0, c->bs1);
    ork_rmsnorm(s->q_a, s->q_a, L->rms_q_a, c->q_lora_rank, c->norm_eps);
    ork_matmul(s->q, s->q_a, &L->wq_b, -1, c->bs0, c->bs1);
  } else {
    ork_matmul(s->q, s->xb, &L->wq, -1, c->bs0, c->bs1);
  }
  ork_matmul(s->kv_a, s->xb, &L->wkv_a, -1, c->bs0, c->bs1);
  int q_pe_offset = c->head_dim - c->qk_rope_head_dim;
  for (int h = 0; h < c->n_heads; h++) {
    float* qpe = s->q + (size_t)h * c->head_dim + q_pe_offset;
    if (c->is_v3) ork_rope_v3(qpe, c->qk_rope_head_dim, c->qk_rope_head_dim, pos, c->rope_theta);
    else ork_rope(qpe, c->qk_rope_head_dim, c->qk_rope_head_dim, pos, c->rope_theta);
  }
  float* k_rope = s->kv_a + c->kv_lora_rank;
  if (c->is_v3) ork_rope_v3(k_rope, c->qk_rope_head_dim, c->qk_rope_head_dim, pos, c->rope_theta);
  else ork_rope(k_rope, c->qk_rope_head_dim, c->qk_rope_head_dim, pos, c->rope_theta);
  ork_rmsnorm(s->kv_a, s->kv_a, L->rms_kv_a, c->kv_lora_rank, c->norm_eps);
  int nope = c->head_dim - c->qk_rope_head_dim;
  ork_matmul(s->kv_b, s->kv_a, &L->wkv_b, -1, c->bs0, c->bs1);
  /* assemble K/V and write the fp16 cache row (979-1002) */
  uint16_t* krow = L->key_cache + (size_t)kv_pos * c->n_heads * c->head_dim;
  uint16_t* vrow = L->value_cache + (size_t)kv_pos * c->n_heads * c->v_head_dim;
  int kvb_stride = nope + c->v_head_dim;
  for (int h = 0; h < c->n_heads; h++) {
    for (int i = 0; i < nope; i++) krow[h * c->head_dim + i] = ork_float_to_half(s->kv_b[h * kvb_stride + i]);
    for (int i = 0; i < c->qk_rope_head_dim; i++) krow[h * c->head_dim + nope + i] = ork_float_to_half(k_rope[i]);
    for (int i = 0; i < c->v_head_dim; i++) vrow[h * c->v_head_dim + i] = ork_float_to_half(s->kv_b[h * kvb_stride + nope + i]);
  }
  /* sink re-rotation by one position (1008-1020) */
  for (int r = 0; r < kv_sink; r++) {
    uint16_t* key = L->key_cache + (size_t)r * c->n_heads * c->head_dim;
    for (int h = 0; h < c->n_heads; h++) {
      uint16_t* kh = key + h * c->head_dim + q_pe_offset;
      if (c->is_v3) ork_rope_v3_f16(kh, c->qk_rope_head_dim, c->qk_rope_head_dim, 1, c->rope_theta);
      else ork_rope_f16(kh, c->qk_rope_head_dim, c->qk_rope_head_dim, 1, c->rope_theta);
    }
  }
#pragma omp parallel for
  for (int h = 0; h < c->n_heads; h++) {
    ork_attn(s->xb2 + (size_t)h * c->v_head_dim, s->att + (size_t)h * c->max_seq_len, s->q + (size_t)h * c->head_dim,
             L->key_cache + h * c->head_dim, L->value_cache + h * c->v_head_dim, c->head_dim, c->v_head_dim,
             c->n_heads, kv_len);
  }
  ork_matmul(s->hb, s->xb2, &L->wo, -1, c->bs0, c->bs1);
}

/* BlockMLA::_attention_impl: src/infer.cpp:1051-1141 */
static void attention_mla(const ork_model* m, ork_state* s, const ork_layer* L, int pos, int kv_sink, int kv_pos,
                          int kv_len) {
  const ork_config* c = &m->cfg;
  const int R = c->qk_rope_head_dim, KL = c->kv_lora_rank;
  if (c->q_lora_rank <= 0) abort(); /* assert, 1057 */
  ork_matmul(s->q_a, s->xb, &L->wq_a, -1, c->bs0, c->bs1);
  ork_rmsnorm(s->q_a, s->q_a, L->rms_q_a, c->q_lora_rank, c->norm_eps);
  ork_matmul(s->kv_a, s->xb, &L->wkv_a, -1, c->bs0, c->bs1);
  ork_matmul(s->q_rope, s->q_a, &L->wq_rope_b, -1, c->bs0, c->bs1);
  ork_matmul(s->q_c, s->q_a, &L->wc, -1, c->bs0, c->bs1);
  for (int h = 0; h < c->n_heads; h++) {
    if (c->is_v3) ork_rope_v3(s->q_rope + (size_t)h * R, R, R, pos, c->rope_theta);
    else ork_rope(s->q_rope + (size_t)h * R, R, R, pos, c->rope_theta);
  }
  float* k_rope = s->kv_a + KL;
  if (c->is_v3) ork_rope_v3(k_rope, R, R, pos, c->rope_theta);
  else ork_rope(k_rope, R, R, pos, c->rope_theta);
  ork_rmsnorm(s->kv_a, s->kv_a, L->rms_kv_a, KL, c->norm_eps); /* latent part only (1079) */
  uint16_t* nrow = L->key_cache + (size_t)kv_pos * KL;           /* kv_nope_cache(kv_pos) */
  uint16_t* rrow = L->value_cache + (size_t)kv_pos * R;          /* kv_rope_cache(kv_pos) */
  for (int i = 0; i < KL; ++i) nrow[i] = ork_float_to_half(s->kv_a[i]);
  for (int i = 0; i < R; ++i) rrow[i] = ork_float_to_half(k_rope[i]);
  for (int r = 0; r < kv_sink; r++) { /* sink rope keys move one position per step (1099-1111) */
    uint16_t* kv = L->value_cache + (size_t)r * R;
    if (c->is_v3) ork_rope_v3_f16(kv, R, R, 1, c->rope_theta);
    else ork_rope_f16(kv, R, R, 1, c->rope_theta);
  }
#pragma omp parallel for
  for (int h = 0; h < c->n_heads; h++) {
    ork_attn_mla(s->xb2 + (size_t)h * KL, s->att + (size_t)h * c->max_seq_len, s->q_c + (size_t)h * KL, s->q_rope + (size_t)h * R,
                 L->key_cache, L->value_cache, c->head_dim, KL, R, kv_len);
  }
  /* per-head value up-projection: matmul_expert with expert = head (1133-1137); kv_b is reused for the outputs */
  for (int h = 0; h < c->n_heads; h++)
    ork_matmul(s->kv_b + (size_t)h * c->v_head_dim, s->xb2 + (size_t)h * KL, &L->wv_b, h, c->bs0, c->bs1);
  ork_matmul(s->hb, s->kv_b, &L->wo, -1, c->bs0, c->bs1);
}

/* Block::_block_cpu: src/infer.cpp:810-932 */
void ork_block(const ork_model* m, ork_state* s, int layer, int pos, int kv_sink, int kv_pos, int kv_len) {
  const ork_config* c = &m->cfg;
  const ork_layer* L = &m->layers[layer];
  ork_rmsnorm(s->xb, s->x, L->rms_att, c->dim, c->norm_eps);
  if (c->use_mla) attention_mla(m, s, L, pos, kv_sink, kv_pos, kv_len);
  else attention_mha(m, s, L, pos, kv_sink, kv_pos, kv_len);
  for (int i = 0; i < c->dim; ++i) s->x[i] += s->hb[i];
  ork_rmsnorm(s->xb, s->x, L->rms_ffn, c->dim, c->norm_eps);
  if (L->is_moe) {
    ork_tensor gate = {ORK_F32, 0, c->n_routed_experts, c->dim, L->moegate, 0};
    ork_matmul(s->moe_weights, s->xb, &gate, -1, 0, 0);
    ork_moe_gate(s->active_experts_weights, L->moegate_bias, s->active_experts, s->moe_weights, c->n_routed_experts,
                 c->n_active_routed, c->norm_topk_prob, c->routed_scaling_factor, c->scoring_sigmoid, c->topk_method,
                 c->n_group, c->topk_group);
    for (int k = 0; k < c->n_active_routed; ++k) {
      int e = s->active_experts[k];
      ork_matmul(s->hb, s->xb, &L->w1, e, c->bs0, c->bs1);
      ork_matmul(s->hb2, s->xb, &L->w3, e, c->bs0, c->bs1);
      for (int i = 0; i < c->moe_intermediate_size; ++i) s->hb[i] = act_fn(c, s->hb[i]) * s->hb2[i];
      ork_matmul(s->xb2, s->hb, &L->w2, e, c->bs0, c->bs1);
      float ew = s->active_experts_weights[k];
      for (int i = 0; i < c->dim; ++i) s->x[i] += s->xb2[i] * ew;
    }
    if (c->n_shared_experts > 0) {
      ork_matmul(s->hb, s->xb, &L->sw1, -1, c->bs0, c->bs1);
      ork_matmul(s->hb2, s->xb, &L->sw3, -1, c->bs0, c->bs1);
      int hs = c->n_shared_experts * c->moe_intermediate_size;
      for (int i = 0; i < hs; ++i) s->hb[i] = act_fn(c, s->hb[i]) * s->hb2[i];
      ork_matmul(s->xb2, s->hb, &L->sw2, -1, c->bs0, c->bs1);
      for (int i = 0; i < c->dim; ++i) s->x[i] += s->xb2[i];
    }
  } else {
    ork_matmul(s->hb, s->xb, &L->w1, -1, c->bs0, c->bs1);
    ork_matmul(s->hb2, s->xb, &L->w3, -1, c->bs0, c->bs1);
    for (int i = 0; i < c->hidden_dim; ++i) s->hb[i] = act_fn(c, s->hb[i]) * s->hb2[i];
    ork_matmul(s->xb2, s->hb, &L->w2, -1, c->bs0, c->bs1);
    for (int i = 0; i < c->dim; ++i) s->x[i] += s->xb2[i];
  }
}

/* Model::_forward_cpu: src/infer.cpp:1265-1317 */
void ork_forward(const ork_model* m, ork_state* s, int token, int pos, int output_logits) {
  const ork_config* c = &m->cfg;
  ork_copy_embedding(m, s, token);
  int omp_ = c->original_max_position;
  int kv_sink = pos >= omp_ ? 2 : 0; /* KV_SINKS = 2, src/model.h:14 */
  int kv_pos = kv_sink + (pos - kv_sink) % (omp_ - kv_sink);
  int kv_len = pos >= omp_ ? omp_ : pos + 1;
  for (int l = 0; l < c->n_layers; l++) ork_block(m, s, l, pos, kv_sink, kv_pos, kv_len);
  if (!output_logits) return;
  ork_rmsnorm(s->x, s->x, m->rms_final, c->dim, c->norm_eps);
  ork_matmul(s->logits, s->x, &m->wcls, -1, c->bs0, c->bs1);
}

/* Sampler::sample_argmax: src/sampler.cpp:28-39 */
int ork_argmax(const float* logits, int n) {
  int argmax = 0;
  float max_val = -FLT_MAX;
  for (int i = 0; i < n; ++i) if (logits[i] > max_val) { max_val = logits[i]; argmax = i; }
  return argmax;
}

/* host thread control for the checkers (tests cap it: the reference forks one parallel region per 128-row band) */
void ork_set_num_threads(int n) { omp_set_num_threads(n); }
