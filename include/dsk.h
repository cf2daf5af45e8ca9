/* include/dsk.h — C-ABI of the B200-native single-batch DeepSeek decoder (libdsk.so).
 *
 * This is the drop-in boundary for the reference's decode hot path.  The reference
 * (andrewkchan/deepseek.cpp @ 8db9e56) has no FFI: its boundary is the link-time seam between
 * src/model.cpp (dispatch) and src/infer.cpp (backend TU) plus `enum class Device` (src/model.h:36-38).
 * Each entry point below names the reference interface it replaces; INTEGRATION.md shows the
 * `Device::CUDA` binding a maintainer would add on the reference side.
 *
 * Conventions: plain C types only; every call returns 0 on success and a negative code on failure
 * (dsk_last_error() has the text); one caller thread per model (like the reference, not re-entrant);
 * there is NO CPU fallback — without a CUDA device every call fails loudly.
 */
#ifndef DSK_H
#define DSK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DSK_ABI_VERSION 3

/* Quant: src/codec.h:79-85 (same numbering) */
enum { DSK_F32 = 0, DSK_F16 = 1, DSK_F8E5M2 = 2, DSK_Q2_K = 3, DSK_Q3_K = 4 };
/* CodecDType: src/codec.h:62-72 (same numbering; .dseek payloads use F32, F16, F8E5M2 and U8) */
enum { DSK_DT_F32 = 0, DSK_DT_F16 = 1, DSK_DT_BF16 = 2, DSK_DT_F8E5M2 = 3, DSK_DT_F8E4M3 = 4, DSK_DT_I32 = 5, DSK_DT_I16 = 6,
       DSK_DT_I8 = 7, DSK_DT_U8 = 8 };
/* InferenceMode: src/model.h:40-43 */
enum { DSK_HYDRATE_KV_CACHE = 0, DSK_OUTPUT_LOGITS = 1 };
/* TopKMethod / ScoringFunc: src/model.h:25-34 */
enum { DSK_TOPK_GREEDY = 0, DSK_TOPK_GROUP_LIMITED_GREEDY = 1 };

/* Config: src/model.h:47-96 (the fields the decode path reads; filled from .dseek metadata, src/model.cpp:22-127).
 * use_mla models (convert.py --mla) need q_lora_rank > 0 (src/infer.cpp:1057), kv_lora_rank % 64 == 0 (K-quants: % 256) and,
 * with f8e5m2 block scales, v_head_dim % bs0 == 0 (matmul_expert's per-head scale offset, src/infer.cpp:437); they are not
 * tensor parallel (routed experts are still sharded).  dsk_model_create() rejects anything else with a message. */
typedef struct dsk_config {
  int dim, hidden_dim, n_layers, n_heads, vocab_size, max_seq_len;
  float rope_theta, norm_eps;
  int act_silu;              /* 1 = SiLU, 0 = GELU (ActivationType) */
  int first_k_dense_replace;
  int n_shared_experts, n_routed_experts, n_active_routed, moe_intermediate_size;
  float routed_scaling_factor;
  int n_group, norm_topk_prob, scoring_sigmoid, topk_group, topk_method;
  int is_v3;                 /* has_moegate_bias: gate bias + interleaved RoPE (src/infer.cpp:958) */
  int kv_lora_rank, q_lora_rank, qk_nope_head_dim, qk_rope_head_dim, v_head_dim;
  int quant;                 /* DSK_* weight quant */
  int bs0, bs1;              /* f8e5m2 scale block (128,128) */
  int original_max_position; /* rope_scaling_original_max_position_embeddings (sinks, src/infer.cpp:1274) */
  int use_mla;               /* 1: true-MLA blocks (BlockMLA, src/model.h:366-453): wc / wq_rope_b / wv_b tensors, latent KV cache */
} dsk_config;

typedef struct dsk_model dsk_model;
typedef struct dsk_state dsk_state;

/* ---- process / device -------------------------------------------------------------------------- */
int dsk_abi_version(void);
const char* dsk_last_error(void);
/* Binds the calling process to CUDA device `device` (one process per GPU). */
int dsk_init(int device);
int dsk_sync(void);
/* Device facts for logs/bench: name (<=128 chars), SM count, HBM bytes. */
int dsk_device_info(char* name128, int* sm_count, size_t* hbm_bytes);

/* ---- model: replaces Model::Model weight binding (src/model.cpp:756-871) + device upload -------- */
/* rank/n_ranks: shard placement, decided here so that dsk_upload_tensor() can keep only this rank's slices.
 * Routed expert e lives on rank e / ceil(E/n_ranks).  With n_ranks > 1 and peer memory (default; DSK_TP=0 or
 * DSK_P2P=0 switch it off) the model is also TENSOR PARALLEL: attention heads (wq/wq_b/wkv_b rows, wo columns, KV cache),
 * the hidden units of the shared experts and of the dense FFN layers (256-blocks) and the LM-head rows are split across
 * the ranks; wq_a / wkv_a / the gate / norms / embedding stay replicated.  Requires n_heads % n_ranks == 0 and
 * (n_heads / n_ranks) * v_head_dim % 256 == 0, otherwise only the experts are sharded. */
dsk_model* dsk_model_create(const dsk_config* cfg, int rank, int n_ranks);
/* What dsk_model_create decided for this rank: tensor_parallel (0/1), local attention heads, local routed experts. */
int dsk_model_sharding(const dsk_model* m, int* tensor_parallel, int* local_heads, int* local_experts);
void dsk_model_destroy(dsk_model* m);
/* Upload one .dseek tensor by its on-disk name (src/model.cpp:766-871), raw payload bytes exactly as
 * stored (K-quants: U8 block rows).  Expert stacks (E, rows, cols): only this rank's slice is kept.
 * dtype / shape are validated against the tensor's role like check_tensor / QTensor::from_codec_tensor
 * (src/model.cpp:129-136, src/codec.cpp:166-234): dtype must be the quant's codec dtype; K-quant payloads are
 * checked by byte count, everything else by the exact 4-slot shape (unused slots 0).  Host payloads travel
 * through a pinned double buffer with asynchronous copies (the call returns once the last chunk is staged;
 * dsk_model_finalize() joins the pipeline).  `src_on_device` != 0 means `data` is already a device pointer
 * (GPU-side minting). */
int dsk_upload_tensor(dsk_model* m, const char* name, int dtype, const int64_t shape[4], const void* data,
                      size_t nbytes, int src_on_device);
/* Checks every tensor the config requires is present (check_tensor, src/model.cpp:129-136). */
int dsk_model_finalize(dsk_model* m);
/* Bytes of weights resident on this GPU / algorithmic weight bytes streamed per decoded token
 * (replaces Model::active_bytes, src/model.cpp:885-901, which under-counts). */
size_t dsk_model_resident_bytes(const dsk_model* m);
double dsk_model_active_bytes_per_token(const dsk_model* m);

/* ---- state: replaces InferenceState (src/model.h:101-179) -------------------------------------- */
dsk_state* dsk_state_create(dsk_model* m);
void dsk_state_destroy(dsk_state* s);
/* Named buffer access for parity taps: "x","xb2","hb","q","q_c" (use_mla),"kv_a","kv_b","moe_weights",
 * "active_experts_weights","logits" (float) and "active_experts" (int32, via the _i32 variant). */
int dsk_state_read(dsk_state* s, const char* buffer, float* dst, size_t n);
int dsk_state_write(dsk_state* s, const char* buffer, const float* src, size_t n);
int dsk_state_read_i32(dsk_state* s, const char* buffer, int32_t* dst, size_t n);
/* fp16 KV cache rows of one layer (BlockMHA::key_cache/value_cache, src/model.h:344-347). which: 0 K, 1 V.
 * use_mla models: which 0 = kv_nope_cache (max_seq_len x kv_lora_rank), 1 = kv_rope_cache (max_seq_len x qk_rope_head_dim)
 * (BlockMLA, src/model.h:437-440). */
int dsk_kv_read(dsk_model* m, int layer, int which, uint16_t* dst, size_t n_halfs);
int dsk_kv_write(dsk_model* m, int layer, int which, const uint16_t* src, size_t n_halfs);

/* ---- forward ----------------------------------------------------------------------------------- */
/* Model::forward(state, token, pos, mode) — src/model.cpp:874-883 -> _forward_cpu src/infer.cpp:1265-1317.
 * host_logits (vocab floats, nullable) receives s.logits(); argmax (nullable) receives
 * Sampler::sample_argmax (src/sampler.cpp:28-39) computed on device. */
int dsk_forward(dsk_model* m, dsk_state* s, int token, int pos, int mode, float* host_logits, int* argmax);
/* Model::_copy_embedding — src/infer.cpp:1217-1263. */
int dsk_copy_embedding(dsk_model* m, dsk_state* s, int token);
/* Block::block(state, pos, kv_sink, kv_pos, kv_len) — src/model.cpp:290-322 -> _block_cpu src/infer.cpp:810-932. */
int dsk_block_forward(dsk_model* m, dsk_state* s, int layer, int pos, int kv_sink, int kv_pos, int kv_len);
/* Device-resident greedy decode (run_completion's sample->forward loop, src/main.cpp:324-335, -t 0):
 * starting from the logits already in `s` (the previous call must have been dsk_forward(..., DSK_OUTPUT_LOGITS)
 * or dsk_decode_greedy — anything else is rejected), generates n_steps tokens with on-device argmax feeding
 * the next forward.  The token loop runs INSIDE one persistent kernel launch: no host round trip and no
 * launch per token.  out_tokens (n_steps ints, nullable).  Returns the device time of the loop in milliseconds
 * through *elapsed_ms (CUDA events), nullable. */
int dsk_decode_greedy(dsk_model* m, dsk_state* s, int start_pos, int n_steps, int32_t* out_tokens,
                      float* elapsed_ms);
/* Number of kernels one forward launches (for bench.py's gpu_launches) */
int dsk_launches_per_forward(const dsk_model* m, int mode);

/* ---- on-device sampling: Sampler (src/sampler.cpp), reading the logits left in `s` by the last forward ----
 * dsk_sample = Sampler::sample(state, temperature, top_p) (src/sampler.cpp:41-75): temperature == 0 -> arg-max
 * (28-39); otherwise softmax at `temperature` and the first vocabulary index whose running probability sum
 * reaches r = coin * top_p — the reference sorts its index array for top_p < 1 but then walks the UNSORTED
 * logits (59-73), so the sort does not change the result and top_p only scales r; reproduced as is.
 * `coin` is the caller's std::rand() / (float)RAND_MAX, so the host keeps the reference's random stream.
 * Only the token id crosses PCIe (8 bytes instead of the vocab-sized logits).
 * dsk_sample_prob = Sampler::sample_prob(index, state) (12-26), used by the perplexity mode. */
int dsk_sample(dsk_model* m, dsk_state* s, float temperature, float top_p, float coin, int* token);
int dsk_sample_prob(dsk_model* m, dsk_state* s, int index, float* prob);

/* ---- multi-GPU ---------------------------------------------------------------------------
 * The reference has no distributed path.  Expert-only sharding has one real exchange step per MoE layer (the MoE partial
 * sum); the tensor-parallel model adds one after the wo projection of every layer (partial sums of the column-sharded wo),
 * one after the dense FFN layers, and one arg-max / logits exchange per token.  All of them run INSIDE the persistent
 * kernel over CUDA-IPC-mapped peer memory: stores straight into every peer's buffer over NVLink, a flag, a fixed-order sum.
 * nccl_unique_id: 128 bytes from dsk_comm_unique_id() on rank 0, broadcast by the launcher (one process
 * per GPU).  dsk_comm_init creates the communicator, then — unless DSK_P2P=0 or a rank lacks peer access —
 * maps one exchange buffer per rank into every peer through CUDA IPC: the decode kernel then stores its
 * partial sums straight into the peers' buffers (NVLink) and waits on flags, and a token stays ONE kernel
 * launch per GPU.  Fallback: ncclAllReduce between kernel segments.  Every rank must make the same
 * sequence of dsk_forward / dsk_block_forward / dsk_decode_greedy calls. */
int dsk_comm_unique_id(void* out128);
int dsk_comm_init(dsk_model* m, const void* nccl_unique_id128);

/* ---- kernel-level test hooks (mirror the statics reached by the reference's tests) ----------------
 * Every hook builds a one-stage program and runs it through the SAME persistent decode kernel that produces the
 * benchmark numbers (decode_kernel<quant>), so the bit-exact / index-exact parity tests cover the hot path itself. */
/* matmul / matmul_unscaled — src/infer.cpp:381-421.  w: raw payload (host), scale nullable. */
int dsk_gemv(int quant, int d, int n, const void* w, const float* scale, int bs0, int bs1, const float* x,
             float* out);
/* The activation vector exactly as the tile loop of a `model_quant` model reads it, after the fused RMSNorm
 * (norm_w nullable): K-quants -> out_q8 = n/256 block_q8_K records (292 B each; quantize_row_q8_K_ref,
 * src/quant.cpp:616-653, bit-exact); F32/F16 -> out_f32 = the fp32 vector (rmsnorm, src/infer.cpp:601-611);
 * F8E5M2 -> out_f32 = the exact fp16 hi/lo split of the tensor-core path, re-assembled ((hi + lo) * 2^-e). */
int dsk_stage_input(int model_quant, const float* x, const float* norm_w, int n, float eps, float* out_f32, void* out_q8);
/* quantize_row_q8_K_ref — src/quant.cpp:616-653 (= dsk_stage_input of a Q2_K model).  out: k/256 block_q8_K. */
int dsk_quantize_q8k(const float* x, int k, void* out);
/* dequantize_row_q{2,3}_K — src/quant.cpp:217-247, 384-432 (through the embedding stage). */
int dsk_dequantize_row(int quant, const void* blocks, int k, float* out);
/* rmsnorm — src/infer.cpp:601-611 (= dsk_stage_input of an F32 model) */
int dsk_rmsnorm(const float* x, const float* w, int n, float eps, float* out);
/* rope / rope_v3 — src/infer.cpp:648-685 (fp32), as the RoPE prologue of the attention stage; d == head_dim <= 128 */
int dsk_rope(float* vec, int d, int head_dim, int pos, float theta, int v3);
/* MoE gate logits of a `model_quant` model: matmul_unscaled(moegate) on rmsnorm(x) — src/infer.cpp:846-851.  gate_w is
 * F32 (n_experts x n) in every quant; out_xnorm (nullable) receives the normalised vector the rows are multiplied with. */
int dsk_gate_logits(int model_quant, int n_experts, int n, const float* gate_w, const float* x, const float* norm_w,
                    float eps, float* out_logits, float* out_xnorm);
/* moe_gate — src/infer.cpp:493-599.  logits (E) in/out (post-softmax/sigmoid+bias scores). */
int dsk_moe_gate(float* logits, const float* bias, int n_routed, int n_active, int norm_topk_prob,
                 float routed_scaling_factor, int scoring_sigmoid, int topk_method, int n_group, int topk_group,
                 int32_t* active_experts, float* weights);
/* attn over all heads — src/infer.cpp:728-762 / mha_cpu 1143-1164.  kcache/vcache fp16 (kv_len rows). */
int dsk_attn(const float* q, const uint16_t* kcache, const uint16_t* vcache, int n_heads, int head_dim,
             int v_head_dim, int kv_len, float* out);

/* ---- measurement hook (bench.py roofline): times `iters` back-to-back launches of ONE GEMV stage of the decode
 * interpreter (production tile plan, TMA ring, warp-per-tile reduction) on a (d x n) matrix of synthetic weights
 * resident in HBM, CUDA events on the launching stream, after `warmup` launches.  n_mats >= 1 distinct matrices
 * are cycled so consecutive launches never re-read L2-resident data.  Returns average milliseconds per launch
 * (launch overhead included) and the algorithmic bytes per launch. */
int dsk_bench_gemv(int quant, int d, int n, int n_mats, int warmup, int iters, float* avg_ms, double* bytes_per_launch);

/* Stage-level timeline of one token from the decode kernel's own globaltimer stamps (CTA 0): a text table (stage kind,
 * count, barrier / staging / tiles / arrive microseconds) written into `out`. */
int dsk_profile_token(dsk_model* m, dsk_state* s, int token, int pos, char* out, size_t cap);

#ifdef __cplusplus
}
#endif
#endif /* DSK_H */
