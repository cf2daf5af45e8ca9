// oracle/ref_shim.cpp — TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// A thin extern "C" harness around the UNMODIFIED reference sources under
// /root/reference/src.  It textually includes the reference's infer.cpp so that its
// file-static kernels (matmul, moe_gate, rmsnorm, rope, ...) are reachable, and is
// compiled with -fno-access-control so Model::_copy_embedding (private) can be driven.
// Built by oracle/Makefile into oracle/_ref/libdsref.so (git-ignored; travels to the
// GPU box with gpurun).  No reference source is copied into this repository.
//
// Every wrapper names the reference symbol (file:line @ 8db9e56) it forwards to.

#include "infer.cpp"  // /root/reference/src/infer.cpp (via -I)

#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "sampler.h"
#include "tokenizer.h"

extern "C" {

// ---- quant.cpp ---------------------------------------------------------------------
// quantize_row_q8_K_ref  src/quant.cpp:616-653
void ref_quantize_q8_K(const float* x, void* y, long k) { quantize_row_q8_K_ref(x, (block_q8_K*)y, k); }
// quantize_row_q2_K_ref  src/quant.cpp:147-215 ; quantize_row_q3_K_ref 308-382
void ref_quantize_q2_K(const float* x, void* y, long k) { quantize_row_q2_K_ref(x, (block_q2_K*)y, k); }
void ref_quantize_q3_K(const float* x, void* y, long k) { quantize_row_q3_K_ref(x, (block_q3_K*)y, k); }
// row-parallel helpers for minting checkpoints (same per-row call as quantizer.cpp:4-34)
void ref_quantize_rows(const float* x, void* y, long rows, long cols, int q3) {
  size_t bs = q3 ? sizeof(block_q3_K) : sizeof(block_q2_K);
  size_t row_bytes = (size_t)(cols / QK_K) * bs;
#pragma omp parallel for
  for (long r = 0; r < rows; r++) {
    if (q3) quantize_row_q3_K_ref(x + r * cols, (block_q3_K*)((char*)y + r * row_bytes), cols);
    else    quantize_row_q2_K_ref(x + r * cols, (block_q2_K*)((char*)y + r * row_bytes), cols);
  }
}
// dequantize_row_q2_K src/quant.cpp:217-247 ; dequantize_row_q3_K 384-432
void ref_dequantize_q2_K(const void* x, float* y, long k) { dequantize_row_q2_K((const block_q2_K*)x, y, k); }
void ref_dequantize_q3_K(const void* x, float* y, long k) { dequantize_row_q3_K((const block_q3_K*)x, y, k); }
// ggml_vec_dot_q2_K_q8_K src/quant.cpp:666-783 ; ggml_vec_dot_q3_K_q8_K 434-614
void ref_vec_dot_q2_K(int n, float* s, const void* vx, const void* vy) { ggml_vec_dot_q2_K_q8_K(n, s, vx, vy); }
void ref_vec_dot_q3_K(int n, float* s, const void* vx, const void* vy) { ggml_vec_dot_q3_K_q8_K(n, s, vx, vy); }

// ---- infer.cpp statics -------------------------------------------------------------
// matmul dispatcher src/infer.cpp:381-417.  quant: 0 F32, 1 F16, 2 F8E5M2, 3 Q2_K, 4 Q3_K.
// scale may be NULL (unscaled); bs = {block rows, block cols}.
void ref_matmul(float* xout, float* x, void* w, int quant, int d, int n, float* scale, int bs0, int bs1) {
  QTensor wt((Quant)quant, {d, n, 0, 0}, w, 0);
  int bs[2] = {bs0, bs1};
  std::vector<uint8_t> aqb(((size_t)n / QK_K + 2) * sizeof(block_q8_K));
  if (scale) {
    QTensor st(Quant::F32, {cdiv(d, bs0), cdiv(n, bs1), 0, 0}, scale, 0);
    matmul(xout, x, wt, bs, st, aqb.data());
  } else {
    matmul(xout, x, wt, bs, std::nullopt, aqb.data());
  }
}
// matmul_expert src/infer.cpp:423-469
void ref_matmul_expert(float* xout, float* x, void* w, int quant, int n_experts, int expert, int d, int n,
                       float* scale, int bs0, int bs1) {
  QTensor wt((Quant)quant, {n_experts, d, n, 0}, w, 0);
  int bs[2] = {bs0, bs1};
  std::vector<uint8_t> aqb(((size_t)n / QK_K + 2) * sizeof(block_q8_K));
  if (scale) {
    QTensor st(Quant::F32, {n_experts, cdiv(d, bs0), cdiv(n, bs1), 0}, scale, 0);
    matmul_expert(xout, x, wt, expert, bs, st, aqb.data());
  } else {
    matmul_expert(xout, x, wt, expert, bs, std::nullopt, aqb.data());
  }
}
// softmax src/infer.cpp:472-487
void ref_softmax(float* o, float* x, int size) { softmax(o, x, size); }
// moe_gate src/infer.cpp:493-599.  `x` (E floats) is modified in place like the reference.
// The GROUP_LIMITED_GREEDY branch reads x[-1] (infer.cpp:558, UB); callers must pass a
// pointer whose preceding float is 0.0f to reproduce the value the reference sees for
// its `new float[]` buffer — ref_moe_gate_padded below does that.
void ref_moe_gate(float* moe_weights, float* bias, int* active_experts, float* x, int n_routed, int n_active,
                  int norm_topk_prob, float routed_scaling_factor, int scoring_sigmoid, int topk_method,
                  int n_group, int topk_group) {
  std::optional<QTensor> b = std::nullopt;
  if (bias) b = QTensor(Quant::F32, {n_routed, 0, 0, 0}, bias, 0);
  moe_gate(moe_weights, b, active_experts, x, n_routed, n_active, norm_topk_prob != 0, routed_scaling_factor,
           scoring_sigmoid ? ScoringFunc::SIGMOID : ScoringFunc::SOFTMAX, (TopKMethod)topk_method, n_group,
           topk_group);
}
void ref_moe_gate_padded(float* moe_weights, float* bias, int* active_experts, float* x, int n_routed, int n_active,
                         int norm_topk_prob, float routed_scaling_factor, int scoring_sigmoid, int topk_method,
                         int n_group, int topk_group) {
  std::vector<float> buf(n_routed + 1, 0.0f);
  memcpy(buf.data() + 1, x, sizeof(float) * n_routed);
  ref_moe_gate(moe_weights, bias, active_experts, buf.data() + 1, n_routed, n_active, norm_topk_prob,
               routed_scaling_factor, scoring_sigmoid, topk_method, n_group, topk_group);
  memcpy(x, buf.data() + 1, sizeof(float) * n_routed);
}
// rmsnorm src/infer.cpp:601-611
void ref_rmsnorm(float* o, float* x, float* w, int size, float eps) { rmsnorm(o, x, w, size, eps); }
// rope (fp32, V2 de-interleaving) src/infer.cpp:648-668 ; rope_v3 670-685
void ref_rope(float* vec, int d, int head_dim, int pos, float theta) {
  std::vector<float> buf(d);
  rope(buf.data(), vec, d, head_dim, pos, theta);
}
void ref_rope_v3(float* vec, int d, int head_dim, int pos, float theta) { rope_v3(vec, d, head_dim, pos, theta); }
// fp16 in-place variants (sink re-rotation) src/infer.cpp:687-707, 709-724
void ref_rope_f16(uint16_t* vec, int d, int head_dim, int pos, float theta) {
  std::vector<float> buf(d);
  rope(buf.data(), (f16_t*)vec, d, head_dim, pos, theta);
}
void ref_rope_v3_f16(uint16_t* vec, int d, int head_dim, int pos, float theta) {
  rope_v3((f16_t*)vec, d, head_dim, pos, theta);
}
// silu src/infer.cpp:640-642 ; gelu 636-638
float ref_silu(float x) { return silu(x); }
float ref_gelu(float x) { return gelu(x); }
// attn src/infer.cpp:728-762
void ref_attn(float* xout, float* atth, const float* qh, const uint16_t* kh, const uint16_t* vh, int head_dim,
              int v_head_dim, int n_heads, int kv_len) {
  attn(xout, atth, qh, (const f16_t*)kh, (const f16_t*)vh, head_dim, v_head_dim, n_heads, kv_len);
}
// half conversions src/codec.h:22-37
float ref_half_to_float(uint16_t h) { return half_to_float(h); }
uint16_t ref_float_to_half(float f) { return float_to_half(f); }

// ---- model level (Model::forward src/model.cpp:874-883; Block::block 290-322) --------
struct RefSession {
  std::unique_ptr<YALMData> data;
  std::unique_ptr<Model> model;
  std::unique_ptr<InferenceState> state;
};

void* ref_session_create(const char* dir, int context) {
  auto* s = new RefSession();
  s->data = std::make_unique<YALMData>(std::string(dir), false);
  s->model = std::make_unique<Model>(*s->data, context);
  s->state = std::make_unique<InferenceState>(s->model->config);
  return s;
}
void ref_session_destroy(void* h) { delete (RefSession*)h; }

int ref_config_int(void* h, const char* key) {
  const Config& c = *((RefSession*)h)->model->config;
  std::string k(key);
  if (k == "dim") return c.dim;
  if (k == "hidden_dim") return c.hidden_dim;
  if (k == "n_layers") return c.n_layers;
  if (k == "n_heads") return c.n_heads;
  if (k == "vocab_size") return c.vocab_size;
  if (k == "max_seq_len") return c.max_seq_len;
  if (k == "n_routed_experts") return c.n_routed_experts;
  if (k == "n_active_routed") return c.n_active_routed;
  if (k == "moe_intermediate_size") return c.moe_intermediate_size;
  if (k == "head_dim") return c.head_dim;
  if (k == "v_head_dim") return c.v_head_dim;
  if (k == "kv_lora_rank") return c.kv_lora_rank;
  if (k == "q_lora_rank") return c.q_lora_rank;
  if (k == "qk_rope_head_dim") return c.qk_rope_head_dim;
  if (k == "first_k_dense_replace") return c.first_k_dense_replace;
  return -1;
}

// Model::forward(state, token, pos, mode)   mode: 0 HYDRATE_KV_CACHE, 1 OUTPUT_LOGITS
void ref_forward(void* h, int token, int pos, int mode) {
  auto* s = (RefSession*)h;
  s->model->forward(*s->state, token, pos, mode ? InferenceMode::OUTPUT_LOGITS : InferenceMode::HYDRATE_KV_CACHE);
}
// Model::_copy_embedding src/infer.cpp:1217-1263 (private; -fno-access-control)
void ref_copy_embedding(void* h, int token) {
  auto* s = (RefSession*)h;
  s->model->_copy_embedding(*s->state, token);
}
// Block::block(s, pos, kv_sink, kv_pos, kv_len)  src/model.cpp:290-322
void ref_block(void* h, int layer, int pos, int kv_sink, int kv_pos, int kv_len) {
  auto* s = (RefSession*)h;
  s->model->blocks[layer]->block(*s->state, pos, kv_sink, kv_pos, kv_len);
}
// Sampler::sample_argmax src/sampler.cpp:28-39
int ref_argmax(void* h) {
  auto* s = (RefSession*)h;
  Sampler smp(s->model->config, 0);
  return smp.sample_argmax(*s->state);
}
// Named views of InferenceState buffers (src/model.h:101-179). Returns element count.
long ref_state_buffer(void* h, const char* name, float** out) {
  auto* s = (RefSession*)h;
  const Config& c = *s->model->config;
  InferenceState& st = *s->state;
  std::string k(name);
  if (k == "x") { *out = st.x(); return c.dim; }
  if (k == "xb") { *out = st.xb(); return c.dim; }
  if (k == "xb2") { *out = st.xb2(); return std::max({c.dim, c.n_heads * c.v_head_dim, c.n_heads * c.kv_lora_rank}); }
  if (k == "hb") { *out = st.hb(); return std::max(c.dim, c.hidden_dim); }
  if (k == "hb2") { *out = st.hb2(); return c.hidden_dim; }
  if (k == "q") { *out = st.q(); return c.n_heads * c.head_dim; }
  if (k == "q_c" && c.use_mla) { *out = st.q_c(); return c.n_heads * c.kv_lora_rank; }
  if (k == "q_rope" && c.use_mla) { *out = st.q_rope(); return c.n_heads * c.qk_rope_head_dim; }
  if (k == "kv_a") { *out = st.kv_a(); return c.kv_lora_rank + c.qk_rope_head_dim; }
  if (k == "kv_b") { *out = st.kv_b(); return c.n_heads * (c.head_dim - c.qk_rope_head_dim + c.v_head_dim); }
  if (k == "logits") { *out = st.logits(); return c.vocab_size; }
  if (k == "moe_weights") { *out = st.moe_weights(); return c.n_routed_experts; }
  if (k == "active_experts_weights") { *out = st.active_experts_weights(); return c.n_active_routed; }
  *out = nullptr;
  return 0;
}
int ref_active_experts(void* h, int* out) {
  auto* s = (RefSession*)h;
  const Config& c = *s->model->config;
  for (int k = 0; k < c.n_active_routed; k++) out[k] = s->state->active_experts()[k];
  return c.n_active_routed;
}
// fp16 KV cache of one block.  MHA (src/model.h:361-362): which 0 key, 1 value.  MLA (src/model.h:411-414): which 0 the
// latent rows (kv_nope_cache), 1 the rope keys (kv_rope_cache).
long ref_kv_cache(void* h, int layer, int which, uint16_t** out) {
  auto* s = (RefSession*)h;
  const Config& c = *s->model->config;
  if (auto* b = dynamic_cast<BlockMHA*>(s->model->blocks[layer].get())) {
    if (which == 0) { *out = b->key_cache(); return (long)c.max_seq_len * c.n_heads * c.head_dim; }
    *out = b->value_cache();
    return (long)c.max_seq_len * c.n_heads * c.v_head_dim;
  }
  if (auto* b = dynamic_cast<BlockMLA*>(s->model->blocks[layer].get())) {
    if (which == 0) { *out = b->kv_nope_cache(); return (long)c.max_seq_len * c.kv_lora_rank; }
    *out = b->kv_rope_cache();
    return (long)c.max_seq_len * c.qk_rope_head_dim;
  }
  *out = nullptr;
  return 0;
}

int ref_num_threads() { return omp_get_max_threads(); }
void ref_set_num_threads(int n) { omp_set_num_threads(n); }

// Timed decode loop for bench.py's reference arm: greedy `steps` tokens after a prompt.
// Mirrors run_completion (src/main.cpp:277-361) minus tokenizer/printing; returns decode seconds.
double ref_timed_decode(void* h, const int* prompt, int n_prompt, int steps, int* out_tokens) {
  auto* s = (RefSession*)h;
  Sampler smp(s->model->config, 0);
  int pos = 0;
  for (; pos < n_prompt; pos++) {
    s->model->forward(*s->state, prompt[pos], pos,
                      pos + 1 == n_prompt ? InferenceMode::OUTPUT_LOGITS : InferenceMode::HYDRATE_KV_CACHE);
  }
  double t0 = omp_get_wtime();
  for (int i = 0; i < steps; i++) {
    int tok = smp.sample_argmax(*s->state);
    if (out_tokens) out_tokens[i] = tok;
    s->model->forward(*s->state, tok, pos++, InferenceMode::OUTPUT_LOGITS);
  }
  return omp_get_wtime() - t0;
}

}  // extern "C"
