// dsk_engine.cu — C-ABI (include/dsk.h) over the sm_100a kernels: device model, state, the per-token
// launch sequence (one CUDA graph per token), expert-shard placement and the NCCL partial-sum all-reduce.
//
// Replaces, behind the reference's call surface (all /root/reference @ 8db9e56):
//   Model::forward -> _forward_cpu        src/model.cpp:874-883, src/infer.cpp:1265-1317   dsk_forward
//   Block::block   -> _block_cpu          src/model.cpp:290-322, src/infer.cpp:810-932     dsk_block_forward
//   BlockMHA::_attention_impl             src/infer.cpp:934-1049                           enqueue_layer (attention part)
//   Model::_copy_embedding                src/infer.cpp:1217-1263                          dsk_copy_embedding
//   Model::Model / Block ctors (binding)  src/model.cpp:149-515, 756-871                   dsk_upload_tensor / finalize
//   InferenceState                        src/model.cpp:677-754                            dsk_state_*
// There is no CPU fallback: every entry point fails if no CUDA device is bound.

#include "../../include/dsk.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>  // types only: the library is dlopen()ed lazily (see NcclApi)

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include <algorithm>
#include "dsk_kernels.cuh"
#include "dsk_mega.cuh"

using namespace dsk;

// ---------------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) return fail(-2, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)
#define CKN(call)                                                                                  \
  do {                                                                                             \
    ncclResult_t e_ = (call);                                                                      \
    if (e_ != ncclSuccess) return fail(-3, "%s failed: %s (%s:%d)", #call, g_nccl.GetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

// NCCL is bound at run time, not link time: a Python host usually has torch's bundled libnccl.so.2 loaded (or
// loads it later); hard-linking the system copy makes the two clash on the shared soname.  dlopen() returns the
// copy already in the process, else the system one.
struct NcclApi {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;
static int nccl_load() {
  if (g_nccl.h) return 0;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return fail(-3, "cannot dlopen libnccl.so.2: %s", dlerror());
  g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(h, "ncclGetUniqueId");
  g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(h, "ncclCommInitRank");
  g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(h, "ncclAllReduce");
  g_nccl.AllGather = (decltype(g_nccl.AllGather))dlsym(h, "ncclAllGather");
  g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(h, "ncclCommDestroy");
  g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(h, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy || !g_nccl.GetErrorString)
    return fail(-3, "libnccl.so.2 lacks a required symbol");
  g_nccl.h = h;
  return 0;
}


static int g_device = -1;
static int g_sm_count = 148;
static bool g_attrs_set = false;

// ---------------------------------------------------------------------------------------------------
// device tensors
// ---------------------------------------------------------------------------------------------------
struct DTensor {
  bool present = false;
  int quant = 0;
  int total_experts = 0;       // 0 = plain matrix
  int expert_first = 0, expert_count = 0;
  int rows = 0, cols = 0;
  uint8_t* w = nullptr;
  float* scale = nullptr;
  size_t row_bytes = 0;        // device bytes per row
  size_t expert_bytes = 0;     // device bytes per expert
  size_t scale_expert = 0;     // scale floats per expert
};

struct Layer {
  float *rms_att = nullptr, *rms_ffn = nullptr, *rms_q_a = nullptr, *rms_kv_a = nullptr;
  DTensor wq, wq_a, wq_b, wkv_a, wkv_b, wo, w1, w2, w3, sw1, sw2, sw3;
  DTensor wc, wq_rope_b, wv_b;   // true-MLA blocks (use_mla): absorbed query projection, rope query projection, per-head value up-projection
  float *gate = nullptr, *gate_bias = nullptr;
  bool is_moe = false;
  __half *kcache = nullptr, *vcache = nullptr;
};

// Host->device weight upload pipeline: the .dseek payload is an mmap of pageable memory, so a plain cudaMemcpy
// stages every tensor synchronously through the driver's bounce buffer.  Here two pinned staging buffers alternate: while
// chunk i travels over PCIe (cudaMemcpy2DAsync on the upload stream, which also re-pitches rows), the host thread copies
// chunk i+1 out of the page cache.  dsk_model_finalize() joins the stream.
struct Uploader {
  static constexpr size_t kChunk = 32u << 20;
  void* pin[2] = {nullptr, nullptr};
  cudaEvent_t ev[2] = {nullptr, nullptr};
  cudaStream_t st = nullptr;
  int cur = 0;
  std::vector<void*> staging;   // device staging buffers (Q3_K repack), freed at finalize
  size_t bytes = 0;
  double seconds = 0;
};

struct dsk_model {
  dsk_config c;
  int head_dim = 0;
  int rank = 0, n_ranks = 1;
  int expert_first = 0, expert_count = 0;
  std::vector<Layer> layers;
  DTensor embed, wcls;
  bool has_wcls = false;
  float* rms_final = nullptr;
  float* rope_freq = nullptr;  // qk_rope_head_dim/2 floats, tabulated on the host (src/infer.cpp:655)
  size_t resident = 0;
  ncclComm_t comm = nullptr;
  // peer-memory exchange of the MoE partial sums (n_ranks > 1): one buffer per rank, mapped into every peer through CUDA IPC
  bool p2p = false;
  int p2p_epoch = 0;
  float* xchg = nullptr;                 // [2 parities][n_ranks][dim] floats, then n_ranks arrival flags
  float* xchg_peer[kMaxRanks] = {};      // the same buffer of every rank, as seen from this process
  unsigned* xflag_peer[kMaxRanks] = {};
  // tensor parallelism (n_ranks > 1, peer memory available): attention heads, wo columns, shared-expert / dense-FFN hidden
  // units and LM-head rows are sharded too; their partial sums / slices travel through the same exchange buffer
  bool tp = false;
  int h0 = 0, nh_loc = 0;                // local attention heads [h0, h0 + nh_loc)
  int sh0 = 0, sh_loc = 0;               // local hidden units of the concatenated shared experts (multiples of 256)
  int hid0 = 0, hid_loc = 0;             // local hidden units of the dense FFN layers
  int v0 = 0, v_loc = 0;                 // local LM-head rows
  float* logits_full = nullptr;          // tp: full-vocabulary logits inside the exchange allocation
  float* logits_peer[kMaxRanks] = {};
  unsigned long long* amax_peer[kMaxRanks] = {};
  long long xchg_done = 0;               // exchanges completed so far through this model's buffer (sequence numbers are per model:
                                         // the buffer and its flags are shared by every dsk_state of the model)
  int n_states = 0;
  Uploader up;
  std::vector<void*> allocs;
};

struct dsk_state {
  dsk_model* m = nullptr;
  float *x = nullptr, *xb2 = nullptr, *hbk = nullptr, *hbs = nullptr, *q_a = nullptr, *q = nullptr, *kv_a = nullptr,
        *kv_b = nullptr, *q_c = nullptr, *moe_logits = nullptr, *moe_scores = nullptr, *act_w = nullptr, *logits = nullptr, *partial = nullptr;
  int* act = nullptr;
  Ctrl* ctrl = nullptr;       // device
  Ctrl* h_ctrl = nullptr;     // pinned host mirror
  int* token_log = nullptr;   // device
  int* step = nullptr;        // device
  float* sample_out = nullptr;   // device: [0] sampled token (as int bits), [1] its probability
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int last_pos = -1;
  bool have_logits = false;   // the last forward ran the LM-head stage: Ctrl::argmax_key / logits are valid
  size_t token_log_cap = 0;
  // persistent interpreter
  Program* prog = nullptr;              // device
  float* att_scratch = nullptr;
  unsigned int* sync_words = nullptr;   // [0] arrivals counter, [1] base
  unsigned long long* tstamp = nullptr;
  std::vector<std::string> stage_names;
  int n_stages = 0;
  size_t mega_smem = 0;
  std::vector<int> layer_begin, layer_end;  // stage ranges per layer
  std::vector<int> cut_after;           // multi-GPU: stage indices followed by the partial-sum all-reduce
  int built_epoch = 0;                  // m->p2p_epoch the program was built for
  int n_xchg = 0;                       // in-kernel exchanges per token (peer-memory mode)
  std::vector<int> xchg_before;         // exchanges preceding stage i within a token (size n_stages + 1)
  int n_tail = 1;                       // stages after the last layer (LM head [+ arg-max exchange]): skipped by HYDRATE_KV_CACHE
  float* logits_src = nullptr;          // where the full logits of the last forward live (tp: the exchange allocation)
};

static size_t disk_row_bytes(int quant, int cols) {
  switch (quant) {
    case DSK_F32: return (size_t)cols * 4;
    case DSK_F16: return (size_t)cols * 2;
    case DSK_F8E5M2: return (size_t)cols;
    case DSK_Q2_K: return (size_t)(cols / 256) * kQ2Bytes;
    case DSK_Q3_K: return (size_t)(cols / 256) * kQ3Disk;
  }
  return 0;
}
static size_t dev_row_bytes(int quant, int cols) {
  if (quant == DSK_Q3_K) return (size_t)(cols / 256) * kQ3Bytes;
  if (quant == DSK_F8E5M2) return f8_pitch((size_t)cols);   // re-pitched rows (see f8_pitch)
  return disk_row_bytes(quant, cols);
}
static int cdiv(int a, int b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------------------------------
// process / device
// ---------------------------------------------------------------------------------------------------
static constexpr int kSmemMax = 227 * 1024;   // opt-in dynamic shared memory per CTA on sm_100
static bool g_f8_mma_ok = true;   // f8 scale-block width is a power of two (the tensor-core loop shifts instead of dividing)
static bool g_coop_small = false, g_kq_small = true;   // DSK_COOP_SMALL=1 / DSK_KQ_SMALL=0: A/B switches of the tile planner
// (cooperative tiles for one-tile-per-CTA F8 stages won 2 us per stage in isolation but lose overall: their code is one more
//  cold path per layer — 393 vs 408 tok/s)
static bool g_use_mma = true;   // F8E5M2 tiles through mma.sync (DSK_NO_MMA=1: CUDA-core dequant path)
enum { ENG_MEGA = 0, ENG_STAGE = 1 };
static int g_engine = ENG_MEGA;

typedef void (*DecodeKernel)(const Program*, int, int, int, int);
static DecodeKernel decode_kernel_for(int quant) {
  switch (quant) {
    case DSK_F32: return decode_kernel<Q_F32>;
    case DSK_F16: return decode_kernel<Q_F16>;
    case DSK_F8E5M2: return decode_kernel<Q_F8>;
    case DSK_Q2_K: return decode_kernel<Q_Q2K>;
    default: return decode_kernel<Q_Q3K>;
  }
}

extern "C" int dsk_abi_version(void) { return DSK_ABI_VERSION; }
extern "C" const char* dsk_last_error(void) { return g_err; }

extern "C" int dsk_init(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(-1, "no CUDA device available (%s) — libdsk has no CPU fallback", cudaGetErrorString(e));
  if (device < 0 || device >= n) return fail(-1, "device %d out of range (have %d)", device, n);
  CK(cudaSetDevice(device));
  g_device = device;
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, device));
  g_sm_count = p.multiProcessorCount;
  int coop = 0;
  CK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device));
  if (!coop) return fail(-1, "device %d does not support cooperative launches (the decode kernel's grid barrier needs them)", device);
  if (!g_attrs_set) {
    g_use_mma = getenv("DSK_NO_MMA") == nullptr;
    if (const char* e = getenv("DSK_COOP_SMALL")) g_coop_small = atoi(e) != 0;
    if (const char* e = getenv("DSK_KQ_SMALL")) g_kq_small = atoi(e) != 0;
    if (const char* en = getenv("DSK_ENGINE")) g_engine = !strcmp(en, "stage") ? ENG_STAGE : ENG_MEGA;
    for (int q = DSK_F32; q <= DSK_Q3_K; q++)
      CK(cudaFuncSetAttribute(decode_kernel_for(q), cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax - 1024));
    g_attrs_set = true;
  }
  return 0;
}
static int need_device() {
  if (g_device < 0) return fail(-1, "dsk_init() has not bound a CUDA device — libdsk has no CPU fallback");
  return 0;
}
extern "C" int dsk_sync(void) {
  if (need_device()) return -1;
  CK(cudaDeviceSynchronize());
  return 0;
}
extern "C" int dsk_device_info(char* name128, int* sm_count, size_t* hbm_bytes) {
  if (need_device()) return -1;
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, g_device));
  if (name128) { strncpy(name128, p.name, 127); name128[127] = 0; }
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (hbm_bytes) *hbm_bytes = p.totalGlobalMem;
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// model
// ---------------------------------------------------------------------------------------------------
static int dmalloc(dsk_model* m, void** p, size_t bytes) {
  CK(cudaMalloc(p, bytes ? bytes : 16));
  m->allocs.push_back(*p);
  m->resident += bytes;
  return 0;
}

// freq_j = 1.0f / powf(theta, j / head_dim) for even j (src/infer.cpp:655, 675) — host libm, like the reference
static std::vector<float> rope_table(int rot, float theta) {
  std::vector<float> f(rot / 2);
  for (int t = 0; t < rot / 2; t++) f[t] = 1.0f / powf(theta, (float)(2 * t) / (float)rot);
  return f;
}

// Limits of the kernels, checked once here so that a checkpoint outside them fails at load time instead of computing garbage
// (the reference handles arbitrary values of some of these; see each message).
static int validate_config(const dsk_config& c) {
  if (c.dim <= 0 || c.n_layers <= 0 || c.n_heads <= 0 || c.vocab_size <= 0 || c.max_seq_len <= 0) return fail(-1, "bad model dimensions");
  if (c.dim % 256 != 0 && (c.quant == DSK_Q2_K || c.quant == DSK_Q3_K)) return fail(-1, "K-quants need dim %% 256 == 0 (src/quant.cpp:617)");
  if (c.n_routed_experts > 256) return fail(-1, "n_routed_experts > 256 unsupported (src/infer.cpp:527)");
  if (c.n_active_routed + 1 > kMaxJobs) return fail(-1, "n_active_routed > %d unsupported", kMaxJobs - 1);
  if (c.n_routed_experts > 0 && c.n_active_routed > c.n_routed_experts) return fail(-1, "n_active_routed > n_routed_experts");
  if (c.qk_rope_head_dim <= 0 || c.qk_rope_head_dim % 2 != 0 || c.qk_rope_head_dim > 128)
    return fail(-1, "qk_rope_head_dim %d unsupported: the attention stage rotates one pair per thread in 64-thread groups (needs an even value <= 128)", c.qk_rope_head_dim);
  if (c.qk_nope_head_dim < 0 || c.v_head_dim <= 0 || c.kv_lora_rank <= 0) return fail(-1, "bad attention dimensions");
  if (c.topk_method == DSK_TOPK_GROUP_LIMITED_GREEDY && c.n_routed_experts > 0) {
    if (c.n_group <= 0 || c.n_routed_experts % c.n_group != 0) return fail(-1, "n_routed_experts %d is not a multiple of n_group %d", c.n_routed_experts, c.n_group);
    if (c.topk_group <= 0 || c.topk_group > c.n_routed_experts / c.n_group) return fail(-1, "topk_group %d does not fit a group of %d experts", c.topk_group, c.n_routed_experts / c.n_group);
  }
  if (c.quant == DSK_F8E5M2) {
    // a weight tile (<= 32 rows, aligned to its height) carries ONE row of block scales: the tile must not straddle a
    // scale-block boundary.  The reference accepts any block size (src/infer.cpp:238-313); quantization_block_size_0 = 128
    // is what convert.py writes.
    if (c.bs0 <= 0 || c.bs1 <= 0) return fail(-1, "f8e5m2 needs quantization_block_size_{0,1} > 0 (got %d, %d)", c.bs0, c.bs1);
    if (c.bs0 % 32 != 0) return fail(-1, "f8e5m2 quantization_block_size_0 = %d unsupported: must be a multiple of 32 (weight tiles of up to 32 rows share one scale row)", c.bs0);
  }
  if (c.original_max_position <= 2) return fail(-1, "rope_scaling_original_max_position_embeddings must exceed the 2 attention sinks");
  if (c.use_mla) {
    if (c.q_lora_rank <= 0) return fail(-1, "use_mla requires q_lora_rank > 0 (src/infer.cpp:1057)");
    if (c.kv_lora_rank % 64 != 0) return fail(-1, "use_mla: kv_lora_rank %d must be a multiple of 64", c.kv_lora_rank);
    if (c.quant == DSK_F8E5M2 && c.bs0 > 0 && c.v_head_dim % c.bs0 != 0)
      return fail(-1, "use_mla with block scales: v_head_dim %d must be a multiple of block_size[0] %d (matmul_expert's per-head scale offset, src/infer.cpp:437)", c.v_head_dim, c.bs0);
    if ((c.quant == DSK_Q2_K || c.quant == DSK_Q3_K) && (c.kv_lora_rank % 256 != 0 || ((size_t)c.v_head_dim * dev_row_bytes(c.quant, c.kv_lora_rank)) % 16 != 0))
      return fail(-1, "use_mla with K-quants: kv_lora_rank %d must be a multiple of 256 and a head's wv_b slab a multiple of 16 bytes", c.kv_lora_rank);
    if (mla_floats(c.kv_lora_rank, c.qk_rope_head_dim, c.max_seq_len) * 4 > 96 * 1024)
      return fail(-1, "use_mla: max_seq_len %d does not fit the attention stage's shared memory (scores are kept on chip)", c.max_seq_len);
  }
  return 0;
}

extern "C" dsk_model* dsk_model_create(const dsk_config* cfg, int rank, int n_ranks) {
  if (need_device()) return nullptr;
  if (!cfg || n_ranks < 1 || n_ranks > kMaxRanks || rank < 0 || rank >= n_ranks) { fail(-1, "bad arguments"); return nullptr; }
  if (validate_config(*cfg)) return nullptr;
  dsk_model* m = new dsk_model();
  m->c = *cfg;
  m->head_dim = cfg->qk_nope_head_dim + cfg->qk_rope_head_dim;
  m->rank = rank;
  m->n_ranks = n_ranks;
  const int E = cfg->n_routed_experts;
  const int per = E > 0 ? cdiv(E, n_ranks) : 0;
  m->expert_first = std::min(E, rank * per);
  m->expert_count = std::max(0, std::min(per, E - m->expert_first));
  {
    std::vector<float> fr = rope_table(cfg->qk_rope_head_dim, cfg->rope_theta);
    if (cudaMalloc(&m->rope_freq, std::max<size_t>(fr.size(), 1) * 4) != cudaSuccess) { fail(-2, "alloc failed"); delete m; return nullptr; }
    m->allocs.push_back(m->rope_freq);
    if (!fr.empty()) cudaMemcpy(m->rope_freq, fr.data(), fr.size() * 4, cudaMemcpyHostToDevice);
  }
  {  // tensor-parallel shard of this rank (block splits: rank r keeps blocks [r B / N, (r + 1) B / N))
    const dsk_config& c = *cfg;
    auto lo = [&](int B, int r) { return (int)((long long)r * B / n_ranks); };
    const char* e_tp = getenv("DSK_TP");
    const char* e_p2p = getenv("DSK_P2P");
    const int sh = c.n_shared_experts * c.moe_intermediate_size;
    bool ok = n_ranks > 1 && !(e_tp && atoi(e_tp) == 0) && !(e_p2p && atoi(e_p2p) == 0) && !c.use_mla;   // (MLA blocks: experts only)
    ok = ok && c.n_heads % n_ranks == 0 && ((c.n_heads / n_ranks) * c.v_head_dim) % 256 == 0;   // wo column slices: whole 256-blocks
    ok = ok && sh % 256 == 0 && c.hidden_dim % 256 == 0;
    if (c.quant == DSK_F8E5M2)                                                                   // slices must not cut a scale block
      ok = ok && c.bs0 == 128 && c.bs1 == 128 && ((c.n_heads / n_ranks) * (c.qk_nope_head_dim + c.qk_rope_head_dim)) % 128 == 0 &&
           ((c.n_heads / n_ranks) * (c.qk_nope_head_dim + c.v_head_dim)) % 128 == 0;
    m->tp = ok;
    m->h0 = 0; m->nh_loc = c.n_heads; m->sh0 = 0; m->sh_loc = sh; m->hid0 = 0; m->hid_loc = c.hidden_dim; m->v0 = 0; m->v_loc = c.vocab_size;
    if (ok) {
      m->nh_loc = c.n_heads / n_ranks; m->h0 = rank * m->nh_loc;
      m->sh0 = lo(sh / 256, rank) * 256; m->sh_loc = lo(sh / 256, rank + 1) * 256 - m->sh0;
      m->hid0 = lo(c.hidden_dim / 256, rank) * 256; m->hid_loc = lo(c.hidden_dim / 256, rank + 1) * 256 - m->hid0;
      const int vb = cdiv(c.vocab_size, 128);
      m->v0 = lo(vb, rank) * 128; m->v_loc = std::min(c.vocab_size, lo(vb, rank + 1) * 128) - m->v0;
    }
  }
  m->layers.resize(cfg->n_layers);
  for (int l = 0; l < cfg->n_layers; l++) {
    Layer& L = m->layers[l];
    L.is_moe = E > 0 && l >= cfg->first_k_dense_replace;
    // MHA blocks: full K / V rows per head (src/model.cpp:459-460); MLA blocks: one latent row + one rope key per token (618-619)
    const size_t kb = (size_t)cfg->max_seq_len * (cfg->use_mla ? (size_t)cfg->kv_lora_rank : (size_t)m->nh_loc * m->head_dim) * sizeof(__half);
    const size_t vb = (size_t)cfg->max_seq_len * (cfg->use_mla ? (size_t)cfg->qk_rope_head_dim : (size_t)m->nh_loc * cfg->v_head_dim) * sizeof(__half);
    if (cudaMalloc(&L.kcache, kb) != cudaSuccess || cudaMalloc(&L.vcache, vb) != cudaSuccess) {
      fail(-2, "KV cache allocation failed (layer %d, %zu bytes)", l, kb + vb);
      delete m;
      return nullptr;
    }
    m->allocs.push_back(L.kcache);
    m->allocs.push_back(L.vcache);
    cudaMemset(L.kcache, 0, kb);
    cudaMemset(L.vcache, 0, vb);
  }
  return m;
}

static void uploader_release(Uploader& u) {
  if (u.st) cudaStreamSynchronize(u.st);
  for (void* p : u.staging) cudaFree(p);
  u.staging.clear();
  for (int i = 0; i < 2; i++) {
    if (u.pin[i]) { cudaFreeHost(u.pin[i]); u.pin[i] = nullptr; }
    if (u.ev[i]) { cudaEventDestroy(u.ev[i]); u.ev[i] = nullptr; }
  }
  if (u.st) { cudaStreamDestroy(u.st); u.st = nullptr; }
}

extern "C" void dsk_model_destroy(dsk_model* m) {
  if (!m) return;
  uploader_release(m->up);
  if (m->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(m->comm);
  for (int q = 0; q < kMaxRanks; q++) if (m->xchg_peer[q] && q != m->rank) cudaIpcCloseMemHandle(m->xchg_peer[q]);
  if (m->xchg) cudaFree(m->xchg);
  for (void* p : m->allocs) cudaFree(p);
  delete m;
}

// rows of `width` payload bytes at `spitch` (source) -> rows at `dpitch` (device).  Host sources go through the pinned
// double buffer on the upload stream; device sources (GPU-side minting) are copied on the legacy default stream, which
// keeps them ordered with the minting framework's own work on that stream.
static int copy_rows(dsk_model* m, void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t nrows, int on_dev) {
  if (nrows == 0 || width == 0) return 0;
  if (on_dev) {
    if (dpitch == spitch && spitch == width) { CK(cudaMemcpy(dst, src, width * nrows, cudaMemcpyDeviceToDevice)); return 0; }
    if (dpitch != width) CK(cudaMemset(dst, 0, dpitch * nrows));
    CK(cudaMemcpy2D(dst, dpitch, src, spitch, width, nrows, cudaMemcpyDeviceToDevice));
    return 0;
  }
  Uploader& u = m->up;
  if (!u.st) {
    CK(cudaStreamCreateWithFlags(&u.st, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) { CK(cudaMallocHost(&u.pin[i], Uploader::kChunk)); CK(cudaEventCreateWithFlags(&u.ev[i], cudaEventDisableTiming)); }
  }
  const bool flat = dpitch == width && spitch == width;
  if (!flat && dpitch != width) CK(cudaMemsetAsync(dst, 0, dpitch * nrows, u.st));
  if (!flat && width > Uploader::kChunk) return fail(-4, "row of %zu bytes exceeds the upload staging buffer", width);
  // flat payloads are cut into staging-buffer-sized pieces; pitched ones into whole rows
  const size_t unit = flat ? 1 : width, total = flat ? width * nrows : nrows;
  const size_t per = flat ? Uploader::kChunk : std::max<size_t>(1, Uploader::kChunk / width);
  for (size_t o = 0; o < total; o += per) {
    const size_t cnt = std::min(per, total - o);
    CK(cudaEventSynchronize(u.ev[u.cur]));   // the copy that last used this staging buffer has drained
    if (flat) {
      memcpy(u.pin[u.cur], (const char*)src + o, cnt);
      CK(cudaMemcpyAsync((char*)dst + o, u.pin[u.cur], cnt, cudaMemcpyHostToDevice, u.st));
    } else {
      const char* s = (const char*)src + o * spitch;
      if (spitch == width) memcpy(u.pin[u.cur], s, cnt * width);
      else for (size_t r = 0; r < cnt; r++) memcpy((char*)u.pin[u.cur] + r * width, s + r * spitch, width);
      CK(cudaMemcpy2DAsync((char*)dst + o * dpitch, dpitch, u.pin[u.cur], width, width, cnt, cudaMemcpyHostToDevice, u.st));
    }
    CK(cudaEventRecord(u.ev[u.cur], u.st));
    u.cur ^= 1;
  }
  (void)unit;
  u.bytes += width * nrows;
  return 0;
}


static bool parse_layer_name(const char* name, int* layer, std::string* rest) {
  const char* pre = "model.layers.";
  size_t n = strlen(pre);
  if (strncmp(name, pre, n) != 0) return false;
  char* end = nullptr;
  long l = strtol(name + n, &end, 10);
  if (end == name + n || *end != '.') return false;
  *layer = (int)l;
  *rest = std::string(end + 1);
  return true;
}

static const char* dtype_name(int dt) {
  switch (dt) { case 0: return "F32"; case 1: return "F16"; case 2: return "BF16"; case 3: return "F8_E5M2"; case 4: return "F8_E4M3";
                case 5: return "I32"; case 6: return "I16"; case 7: return "I8"; case 8: return "U8"; }
  return "?";
}
// QTensor::from_codec_tensor (src/codec.cpp:166-234): dtype must be the quant's codec dtype; K-quants are checked by byte
// count (their U8 payload shape is free), everything else by the exact 4-slot shape (unused slots zero).
static int check_tensor(const char* name, int dtype, const int64_t shape[4], size_t nbytes, int want_dtype, bool kquant,
                        const int64_t want[4], size_t want_bytes) {
  if (dtype != want_dtype)
    return fail(-4, "tensor mismatch for %s: expected dtype=%s, got dtype=%s", name, dtype_name(want_dtype), dtype_name(dtype));
  if (kquant) {
    if (nbytes != want_bytes) return fail(-4, "tensor mismatch for %s: expected dtype=U8, size=%zu; got size=%zu", name, want_bytes, nbytes);
    return 0;
  }
  for (int i = 0; i < 4; i++)
    if (shape[i] != want[i])
      return fail(-4, "tensor mismatch for %s: expected shape=[%lld,%lld,%lld,%lld], got shape=[%lld,%lld,%lld,%lld]", name,
                  (long long)want[0], (long long)want[1], (long long)want[2], (long long)want[3],
                  (long long)shape[0], (long long)shape[1], (long long)shape[2], (long long)shape[3]);
  if (nbytes != want_bytes) return fail(-4, "tensor %s: expected %zu bytes, got %zu", name, want_bytes, nbytes);
  return 0;
}

static int upload_f32(dsk_model* m, float** dst, int64_t d0, int64_t d1, int dtype, const int64_t shape[4], const void* data,
                      size_t nbytes, int on_dev, const char* name) {
  const int64_t want[4] = {d0, d1, 0, 0};
  const size_t n_expected = (size_t)d0 * (size_t)(d1 > 0 ? d1 : 1);
  if (check_tensor(name, dtype, shape, nbytes, DSK_DT_F32, false, want, n_expected * 4)) return -4;
  if (*dst) return fail(-4, "tensor %s uploaded twice", name);
  if (dmalloc(m, (void**)dst, nbytes)) return -2;
  return copy_rows(m, *dst, nbytes, data, nbytes, nbytes, 1, on_dev);
}

// bytes of `cols` consecutive columns of a disk row (cols: a multiple of the quant's block width)
static size_t disk_col_bytes(int quant, int cols) { return disk_row_bytes(quant, cols); }

// (rows x cols) logical tensor; this rank keeps rows [r0, r0 + rc) and columns [c0, c0 + cc) of it (tensor-parallel slices;
// the whole tensor when rc == rows and cc == cols) and, for expert stacks, its expert range.  dtype / shape are checked
// against the FULL tensor, like the reference does.
static int upload_weight(dsk_model* m, DTensor& t, int rows, int cols, bool expert, bool is_scale, int r0, int rc, int c0, int cc,
                         int dtype, const int64_t shape[4], const void* data, size_t nbytes, int on_dev, const char* name) {
  const dsk_config& c = m->c;
  const int E = expert ? c.n_routed_experts : 0;
  const int first = expert ? m->expert_first : 0;
  const int count = expert ? m->expert_count : 1;
  if (is_scale) {
    if (c.quant != DSK_F8E5M2) return fail(-4, "tensor %s: only f8e5m2 checkpoints carry scales (src/model.cpp:191)", name);
    const int sr = cdiv(rows, c.bs0), sc = cdiv(cols, c.bs1);
    const size_t per = (size_t)sr * sc;
    const size_t tot = per * (expert ? E : 1);
    const int64_t want_e[4] = {E, sr, sc, 0}, want_p[4] = {sr, sc, 0, 0};
    if (check_tensor(name, dtype, shape, nbytes, DSK_DT_F32, false, expert ? want_e : want_p, tot * 4)) return -4;
    if (t.scale) return fail(-4, "tensor %s uploaded twice", name);
    // local block range (slices start on scale-block boundaries: validated at model creation)
    const int sr0 = r0 / c.bs0, src = cdiv(rc, c.bs0), sc0 = c0 / c.bs1, scc = cdiv(cc, c.bs1);
    const size_t per_loc = (size_t)src * scc;
    t.scale_expert = per_loc;
    const size_t local = per_loc * count;
    if (dmalloc(m, (void**)&t.scale, local * 4)) return -2;
    if (local == 0) return 0;
    if (src == sr && scc == sc)
      return copy_rows(m, t.scale, local * 4, (const char*)data + (size_t)first * per * 4, local * 4, local * 4, 1, on_dev);
    return copy_rows(m, t.scale, (size_t)scc * 4, (const char*)data + ((size_t)sr0 * sc + sc0) * 4, (size_t)sc * 4, (size_t)scc * 4, src, on_dev);
  }
  const int q = c.quant;
  const bool kq = q == DSK_Q2_K || q == DSK_Q3_K;
  if (kq && cols % 256 != 0) return fail(-4, "tensor %s: cols %d not a multiple of 256", name, cols);
  if ((q == DSK_F16 || q == DSK_F8E5M2) && cols % 16 != 0) return fail(-4, "tensor %s: cols %d not a multiple of 16", name, cols);
  if (q == DSK_F32 && cols % 4 != 0) return fail(-4, "tensor %s: cols %d not a multiple of 4", name, cols);
  const size_t drb = disk_row_bytes(q, cols);
  const size_t tot = drb * rows * (expert ? E : 1);
  static const int codec_of_quant[5] = {DSK_DT_F32, DSK_DT_F16, DSK_DT_F8E5M2, DSK_DT_U8, DSK_DT_U8};   // quant_to_codec_dtype
  const int64_t want_e[4] = {E, rows, cols, 0}, want_p[4] = {rows, cols, 0, 0};
  if (check_tensor(name, dtype, shape, nbytes, codec_of_quant[q], kq, expert ? want_e : want_p, tot)) return -4;
  if (t.present) return fail(-4, "tensor %s uploaded twice", name);
  const size_t drb_loc = disk_col_bytes(q, cc), vrb = dev_row_bytes(q, cc);
  t.present = true;
  t.quant = q;
  t.total_experts = E;
  t.expert_first = first;
  t.expert_count = count;
  t.rows = rc;
  t.cols = cc;
  t.row_bytes = vrb;
  t.expert_bytes = vrb * rc;
  const size_t nrows_loc = (size_t)rc * count;
  const size_t local_disk = drb_loc * nrows_loc, local_dev = vrb * nrows_loc;
  if (dmalloc(m, (void**)&t.w, local_dev + 16)) return -2;
  const char* src = (const char*)data + (size_t)first * drb * rows + (size_t)r0 * drb + disk_col_bytes(q, c0);
  if (local_disk == 0) return 0;
  if (q == DSK_Q3_K) {
    // stage the 110-byte disk blocks (compacted to the local columns), repack to 112-byte aligned blocks on the device
    const unsigned char* dsrc = (const unsigned char*)src;
    cudaStream_t st = nullptr;
    if (!on_dev || drb_loc != drb) {
      void* staging = nullptr;
      CK(cudaMalloc(&staging, local_disk));
      m->up.staging.push_back(staging);
      if (copy_rows(m, staging, drb_loc, src, drb, drb_loc, nrows_loc, on_dev)) return -2;
      dsrc = (const unsigned char*)staging;
      st = on_dev ? nullptr : m->up.st;
    }
    const size_t nblocks = local_disk / kQ3Disk;
    q3k_repack_kernel<<<(unsigned)((nblocks + 7) / 8), 256, 0, st>>>(dsrc, t.w, nblocks);
    CK(cudaGetLastError());
    if (on_dev) CK(cudaDeviceSynchronize());
    return 0;
  }
  return copy_rows(m, t.w, vrb, src, drb, drb_loc, nrows_loc, on_dev);
}

extern "C" int dsk_upload_tensor(dsk_model* m, const char* name, int dtype, const int64_t shape[4], const void* data,
                                 size_t nbytes, int src_on_device) {
  if (need_device()) return -1;
  if (!m || !name || !data || !shape) return fail(-1, "bad arguments");
  const dsk_config& c = m->c;
  const int hd = m->head_dim, nope = c.qk_nope_head_dim, mi = c.moe_intermediate_size;
  std::string nm(name);
  if (nm == "tokenizer.tokens") return 0;  // host-side only
  auto W = [&](DTensor& t, int rows, int cols, bool expert, bool is_scale) {
    return upload_weight(m, t, rows, cols, expert, is_scale, 0, rows, 0, cols, dtype, shape, data, nbytes, src_on_device, name);
  };
  auto WS = [&](DTensor& t, int rows, int cols, bool is_scale, int r0, int rc, int c0, int cc) {
    return upload_weight(m, t, rows, cols, false, is_scale, r0, rc, c0, cc, dtype, shape, data, nbytes, src_on_device, name);
  };
  auto F = [&](float** dst, int64_t d0, int64_t d1) { return upload_f32(m, dst, d0, d1, dtype, shape, data, nbytes, src_on_device, name); };
  if (nm == "model.embed.weight") return W(m->embed, c.vocab_size, c.dim, false, false);
  if (nm == "model.embed.scale") return W(m->embed, c.vocab_size, c.dim, false, true);
  if (nm == "model.output.weight") { m->has_wcls = true; return WS(m->wcls, c.vocab_size, c.dim, false, m->v0, m->v_loc, 0, c.dim); }
  if (nm == "model.output.scale") return WS(m->wcls, c.vocab_size, c.dim, true, m->v0, m->v_loc, 0, c.dim);
  if (nm == "model.norm.weight") return F(&m->rms_final, c.dim, 0);
  int l = -1;
  std::string rest;
  if (!parse_layer_name(name, &l, &rest) || l < 0 || l >= c.n_layers) return fail(-4, "unknown tensor name %s", name);
  Layer& L = m->layers[l];
  if (rest == "attn.norm.weight") return F(&L.rms_att, c.dim, 0);
  if (rest == "mlp.norm.weight") return F(&L.rms_ffn, c.dim, 0);
  if (rest == "attn.kv_a_norm.weight") return F(&L.rms_kv_a, c.kv_lora_rank, 0);
  if (rest == "attn.q_a_norm.weight") return F(&L.rms_q_a, c.q_lora_rank, 0);
  if (rest == "moegate.weight") return F(&L.gate, c.n_routed_experts, c.dim);
  if (rest == "moegate.bias") return F(&L.gate_bias, c.n_routed_experts, 0);
  const size_t dot = rest.rfind('.');
  if (dot == std::string::npos) return fail(-4, "unknown tensor name %s", name);
  const std::string base = rest.substr(0, dot), kind = rest.substr(dot + 1);
  const bool is_scale = kind == "scale";
  if (!is_scale && kind != "weight") return fail(-4, "unknown tensor name %s", name);
  const int sh = c.n_shared_experts * mi;
  // tensor-parallel slices (the whole tensor when tp is off: h0 = 0, nh_loc = n_heads, ...)
  const int per_kv = nope + c.v_head_dim;
  const int hr0 = m->h0 * hd, hrc = m->nh_loc * hd;                    // wq / wq_b rows of the local heads
  const int kr0 = m->h0 * per_kv, krc = m->nh_loc * per_kv;            // wkv_b rows
  const int oc0 = m->h0 * c.v_head_dim, occ = m->nh_loc * c.v_head_dim;   // wo columns
  if (c.use_mla) {   // BlockMLA tensors (src/model.cpp:558-610)
    if (base == "attn.wc") return W(L.wc, c.n_heads * c.kv_lora_rank, c.q_lora_rank, false, is_scale);
    if (base == "attn.wq_rope_b") return W(L.wq_rope_b, c.n_heads * c.qk_rope_head_dim, c.q_lora_rank, false, is_scale);
    if (base == "attn.wv_b") return W(L.wv_b, c.n_heads * c.v_head_dim, c.kv_lora_rank, false, is_scale);
  }
  if (base == "attn.wq") return WS(L.wq, c.n_heads * hd, c.dim, is_scale, hr0, hrc, 0, c.dim);
  if (base == "attn.wq_a") return W(L.wq_a, c.q_lora_rank, c.dim, false, is_scale);
  if (base == "attn.wq_b") return WS(L.wq_b, c.n_heads * hd, c.q_lora_rank, is_scale, hr0, hrc, 0, c.q_lora_rank);
  if (base == "attn.wkv_a") return W(L.wkv_a, c.kv_lora_rank + c.qk_rope_head_dim, c.dim, false, is_scale);
  if (base == "attn.wkv_b") return WS(L.wkv_b, c.n_heads * per_kv, c.kv_lora_rank, is_scale, kr0, krc, 0, c.kv_lora_rank);
  if (base == "attn.wo") return WS(L.wo, c.dim, c.n_heads * c.v_head_dim, is_scale, 0, c.dim, oc0, occ);
  if (L.is_moe) {
    if (base == "mlp.w1") return W(L.w1, mi, c.dim, true, is_scale);
    if (base == "mlp.w2") return W(L.w2, c.dim, mi, true, is_scale);
    if (base == "mlp.w3") return W(L.w3, mi, c.dim, true, is_scale);
  } else {
    if (base == "mlp.w1") return WS(L.w1, c.hidden_dim, c.dim, is_scale, m->hid0, m->hid_loc, 0, c.dim);
    if (base == "mlp.w2") return WS(L.w2, c.dim, c.hidden_dim, is_scale, 0, c.dim, m->hid0, m->hid_loc);
    if (base == "mlp.w3") return WS(L.w3, c.hidden_dim, c.dim, is_scale, m->hid0, m->hid_loc, 0, c.dim);
  }
  if (base == "shared_mlp.w1") return WS(L.sw1, sh, c.dim, is_scale, m->sh0, m->sh_loc, 0, c.dim);
  if (base == "shared_mlp.w2") return WS(L.sw2, c.dim, sh, is_scale, 0, c.dim, m->sh0, m->sh_loc);
  if (base == "shared_mlp.w3") return WS(L.sw3, sh, c.dim, is_scale, m->sh0, m->sh_loc, 0, c.dim);
  return fail(-4, "unknown tensor name %s (MLA-mode tensors are not part of this path)", name);
}

extern "C" int dsk_model_finalize(dsk_model* m) {
  if (need_device()) return -1;
  if (!m) return fail(-1, "null model");
  const dsk_config& c = m->c;
  const bool f8 = c.quant == DSK_F8E5M2;
  auto need = [&](const DTensor& t, const char* what, int l) -> int {
    if (!t.present) return fail(-5, "missing tensor %s (layer %d)", what, l);
    if (f8 && !t.scale) return fail(-5, "missing scale for %s (layer %d)", what, l);
    return 0;
  };
  if (need(m->embed, "model.embed", -1)) return -5;
  if (!m->rms_final) return fail(-5, "missing model.norm.weight");
  if (!m->has_wcls) m->wcls = m->embed;  // tied embeddings (src/model.cpp:852-855)
  for (int l = 0; l < c.n_layers; l++) {
    Layer& L = m->layers[l];
    if (!L.rms_att || !L.rms_ffn || !L.rms_kv_a) return fail(-5, "missing norm weights (layer %d)", l);
    if (c.use_mla) {
      if (!L.rms_q_a) return fail(-5, "missing q_a_norm (layer %d)", l);
      if (need(L.wq_a, "attn.wq_a", l) || need(L.wc, "attn.wc", l) || need(L.wq_rope_b, "attn.wq_rope_b", l) || need(L.wv_b, "attn.wv_b", l)) return -5;
    } else if (c.q_lora_rank > 0) {
      if (!L.rms_q_a) return fail(-5, "missing q_a_norm (layer %d)", l);
      if (need(L.wq_a, "attn.wq_a", l) || need(L.wq_b, "attn.wq_b", l)) return -5;
    } else if (need(L.wq, "attn.wq", l)) return -5;
    if (need(L.wkv_a, "attn.wkv_a", l) || (!c.use_mla && need(L.wkv_b, "attn.wkv_b", l)) || need(L.wo, "attn.wo", l)) return -5;
    if (need(L.w1, "mlp.w1", l) || need(L.w2, "mlp.w2", l) || need(L.w3, "mlp.w3", l)) return -5;
    if (L.is_moe) {
      if (!L.gate) return fail(-5, "missing moegate.weight (layer %d)", l);
      if (c.is_v3 && !L.gate_bias) return fail(-5, "missing moegate.bias (layer %d)", l);
      if (c.n_shared_experts > 0 && (need(L.sw1, "shared_mlp.w1", l) || need(L.sw2, "shared_mlp.w2", l) || need(L.sw3, "shared_mlp.w3", l)))
        return -5;
    }
  }
  // join the upload pipeline and give its pinned staging buffers back
  if (m->up.st) {
    cudaError_t e = cudaStreamSynchronize(m->up.st);
    if (e != cudaSuccess) return fail(-2, "weight upload failed: %s", cudaGetErrorString(e));
  }
  uploader_release(m->up);
  CK(cudaDeviceSynchronize());
  return 0;
}

extern "C" size_t dsk_model_resident_bytes(const dsk_model* m) { return m ? m->resident : 0; }
extern "C" int dsk_model_sharding(const dsk_model* m, int* tensor_parallel, int* local_heads, int* local_experts) {
  if (!m) return fail(-1, "null model");
  if (tensor_parallel) *tensor_parallel = m->tp ? 1 : 0;
  if (local_heads) *local_heads = m->nh_loc;
  if (local_experts) *local_experts = m->expert_count;
  return 0;
}

// Algorithmic weight bytes per decoded token: sum over executed GEMVs of rows*cols*bpw,
// bpw = 4 / 2 / 1+4/(bs0*bs1) / 84/256 / 110/256, plus the F32 gate, gate bias and norm weights.
extern "C" double dsk_model_active_bytes_per_token(const dsk_model* m) {
  if (!m) return 0;
  const dsk_config& c = m->c;
  double bpw = 4;
  switch (c.quant) {
    case DSK_F16: bpw = 2; break;
    case DSK_F8E5M2: bpw = 1.0 + 4.0 / ((double)c.bs0 * c.bs1); break;
    case DSK_Q2_K: bpw = 84.0 / 256; break;
    case DSK_Q3_K: bpw = 110.0 / 256; break;
    default: break;
  }
  const double hd = m->head_dim, nope = c.qk_nope_head_dim, mi = c.moe_intermediate_size, dim = c.dim;
  double total = 0;
  for (int l = 0; l < c.n_layers; l++) {
    double w = 0;
    if (c.use_mla) w += (double)c.q_lora_rank * dim + (double)c.n_heads * (c.kv_lora_rank + c.qk_rope_head_dim) * c.q_lora_rank;   // wq_a, wc, wq_rope_b
    else if (c.q_lora_rank > 0) w += (double)c.q_lora_rank * dim + (double)c.n_heads * hd * c.q_lora_rank;
    else w += (double)c.n_heads * hd * dim;
    w += (double)(c.kv_lora_rank + c.qk_rope_head_dim) * dim;
    if (c.use_mla) w += (double)c.n_heads * c.v_head_dim * c.kv_lora_rank;                                                             // wv_b
    else w += (double)c.n_heads * (nope + c.v_head_dim) * c.kv_lora_rank;
    w += dim * c.n_heads * c.v_head_dim;
    double f32 = 2 * dim + c.kv_lora_rank + c.q_lora_rank;  // norm weights
    if (m->layers[l].is_moe) {
      w += 3.0 * c.n_active_routed * mi * dim + 3.0 * c.n_shared_experts * mi * dim;
      f32 += (double)c.n_routed_experts * dim + (c.is_v3 ? c.n_routed_experts : 0);
    } else {
      w += 3.0 * c.hidden_dim * dim;
    }
    total += w * bpw + f32 * 4;
  }
  total += dim * bpw;                          // embedding row
  total += (double)c.vocab_size * dim * bpw;   // LM head
  total += dim * 4;                            // final norm
  return total;
}

// ---------------------------------------------------------------------------------------------------
// state
// ---------------------------------------------------------------------------------------------------
static int build_program(dsk_model* m, dsk_state* s);

extern "C" dsk_state* dsk_state_create(dsk_model* m) {
  if (need_device() || !m) return nullptr;
  const dsk_config& c = m->c;
  dsk_state* s = new dsk_state();
  s->m = m;
  auto fa = [&](float** p, size_t n) { cudaMalloc((void**)p, std::max<size_t>(n, 4) * 4); cudaMemset(*p, 0, std::max<size_t>(n, 4) * 4); };
  const int mi = c.moe_intermediate_size;
  fa(&s->x, c.dim);
  fa(&s->xb2, std::max(std::max(c.dim, c.n_heads * c.v_head_dim), c.use_mla ? c.n_heads * c.kv_lora_rank : 0));
  fa(&s->q_c, c.use_mla ? (size_t)c.n_heads * c.kv_lora_rank : 4);
  fa(&s->hbk, (size_t)std::max(1, c.n_active_routed) * std::max(1, mi));
  fa(&s->hbs, std::max(c.hidden_dim, c.n_shared_experts * mi));
  fa(&s->q_a, std::max(1, c.q_lora_rank));
  fa(&s->q, (size_t)c.n_heads * m->head_dim);
  fa(&s->kv_a, c.kv_lora_rank + c.qk_rope_head_dim);
  fa(&s->kv_b, (size_t)c.n_heads * (c.qk_nope_head_dim + c.v_head_dim));
  fa(&s->moe_logits, std::max(1, c.n_routed_experts));
  fa(&s->moe_scores, std::max(1, c.n_routed_experts));
  fa(&s->act_w, 16);
  fa(&s->logits, c.vocab_size);
  fa(&s->partial, c.dim);
  cudaMalloc((void**)&s->act, 16 * sizeof(int));
  cudaMemset(s->act, 0, 16 * sizeof(int));
  cudaMalloc((void**)&s->ctrl, sizeof(Ctrl));
  cudaMemset(s->ctrl, 0, sizeof(Ctrl));
  cudaMallocHost((void**)&s->h_ctrl, sizeof(Ctrl));
  memset(s->h_ctrl, 0, sizeof(Ctrl));
  s->token_log_cap = 1 << 16;
  cudaMalloc((void**)&s->token_log, s->token_log_cap * sizeof(int));
  cudaMalloc((void**)&s->step, sizeof(int));
  cudaMemset(s->step, 0, sizeof(int));
  cudaMalloc((void**)&s->sample_out, 4 * sizeof(float));
  cudaMemset(s->sample_out, 0, 4 * sizeof(float));
  cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
  cudaEventCreate(&s->ev0);
  cudaEventCreate(&s->ev1);
  if (cudaGetLastError() != cudaSuccess) { fail(-2, "state allocation failed"); return nullptr; }
  if (build_program(m, s)) { return nullptr; }
  m->n_states++;
  return s;
}

extern "C" void dsk_state_destroy(dsk_state* s) {
  if (!s) return;
  if (s->m) s->m->n_states--;
  float* bufs[] = {s->x, s->xb2, s->hbk, s->hbs, s->q_a, s->q, s->q_c, s->kv_a, s->kv_b, s->moe_logits, s->moe_scores, s->act_w, s->logits, s->partial};
  for (float* b : bufs) cudaFree(b);
  cudaFree(s->prog); cudaFree(s->att_scratch); cudaFree(s->sync_words); cudaFree(s->tstamp);
  cudaFree(s->sample_out); cudaFree(s->act); cudaFree(s->ctrl); cudaFreeHost(s->h_ctrl); cudaFree(s->token_log); cudaFree(s->step);
  cudaEventDestroy(s->ev0); cudaEventDestroy(s->ev1);
  cudaStreamDestroy(s->stream);
  delete s;
}

static float* state_buf(dsk_state* s, const char* name, size_t* cap) {
  const dsk_config& c = s->m->c;
  std::string k(name);
  if (k == "x") { *cap = c.dim; return s->x; }
  if (k == "xb2") { *cap = std::max(std::max(c.dim, c.n_heads * c.v_head_dim), c.use_mla ? c.n_heads * c.kv_lora_rank : 0); return s->xb2; }
  if (k == "q_c" && c.use_mla) { *cap = (size_t)c.n_heads * c.kv_lora_rank; return s->q_c; }
  if (k == "hb") { *cap = std::max(c.hidden_dim, c.n_shared_experts * c.moe_intermediate_size); return s->hbs; }
  if (k == "hb_routed") { *cap = (size_t)c.n_active_routed * c.moe_intermediate_size; return s->hbk; }
  if (k == "q_a") { *cap = c.q_lora_rank; return s->q_a; }
  if (k == "q") { *cap = (size_t)c.n_heads * s->m->head_dim; return s->q; }
  if (k == "kv_a") { *cap = c.kv_lora_rank + c.qk_rope_head_dim; return s->kv_a; }
  if (k == "kv_b") { *cap = (size_t)c.n_heads * (c.qk_nope_head_dim + c.v_head_dim); return s->kv_b; }
  if (k == "moe_weights") { *cap = c.n_routed_experts; return s->moe_scores; }
  if (k == "active_experts_weights") { *cap = c.n_active_routed; return s->act_w; }
  if (k == "logits") { *cap = c.vocab_size; return s->logits_src ? s->logits_src : s->logits; }
  return nullptr;
}
extern "C" int dsk_state_read(dsk_state* s, const char* buffer, float* dst, size_t n) {
  if (need_device() || !s) return -1;
  size_t cap = 0;
  float* p = state_buf(s, buffer, &cap);
  if (!p) return fail(-4, "unknown state buffer %s", buffer);
  if (n > cap) return fail(-4, "state buffer %s holds %zu floats, asked %zu", buffer, cap, n);
  CK(cudaStreamSynchronize(s->stream));
  CK(cudaMemcpy(dst, p, n * 4, cudaMemcpyDeviceToHost));
  return 0;
}
extern "C" int dsk_state_write(dsk_state* s, const char* buffer, const float* src, size_t n) {
  if (need_device() || !s) return -1;
  size_t cap = 0;
  float* p = state_buf(s, buffer, &cap);
  if (!p) return fail(-4, "unknown state buffer %s", buffer);
  if (n > cap) return fail(-4, "state buffer %s holds %zu floats, asked %zu", buffer, cap, n);
  CK(cudaStreamSynchronize(s->stream));
  CK(cudaMemcpy(p, src, n * 4, cudaMemcpyHostToDevice));
  return 0;
}
extern "C" int dsk_state_read_i32(dsk_state* s, const char* buffer, int32_t* dst, size_t n) {
  if (need_device() || !s) return -1;
  if (std::string(buffer) != "active_experts") return fail(-4, "unknown int buffer %s", buffer);
  if (n > 16) return fail(-4, "active_experts holds at most 16 ints");
  CK(cudaStreamSynchronize(s->stream));
  CK(cudaMemcpy(dst, s->act, n * 4, cudaMemcpyDeviceToHost));
  return 0;
}
static int kv_ptr(dsk_model* m, int layer, int which, __half** p, size_t* cap) {
  if (!m || layer < 0 || layer >= m->c.n_layers) return fail(-4, "bad layer");
  Layer& L = m->layers[layer];
  *p = which == 0 ? L.kcache : L.vcache;
  *cap = m->c.use_mla ? (size_t)m->c.max_seq_len * (which == 0 ? m->c.kv_lora_rank : m->c.qk_rope_head_dim)
                      : (size_t)m->c.max_seq_len * m->nh_loc * (which == 0 ? m->head_dim : m->c.v_head_dim);
  return 0;
}
extern "C" int dsk_kv_read(dsk_model* m, int layer, int which, uint16_t* dst, size_t n) {
  if (need_device()) return -1;
  __half* p = nullptr; size_t cap = 0;
  if (kv_ptr(m, layer, which, &p, &cap)) return -4;
  if (n > cap) return fail(-4, "kv cache holds %zu halfs", cap);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(dst, p, n * 2, cudaMemcpyDeviceToHost));
  return 0;
}
extern "C" int dsk_kv_write(dsk_model* m, int layer, int which, const uint16_t* src, size_t n) {
  if (need_device()) return -1;
  __half* p = nullptr; size_t cap = 0;
  if (kv_ptr(m, layer, which, &p, &cap)) return -4;
  if (n > cap) return fail(-4, "kv cache holds %zu halfs", cap);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(p, src, n * 2, cudaMemcpyHostToDevice));
  return 0;
}

#define CKL(call)                                                                                   \
  do {                                                                                              \
    cudaError_t e_ = (call);                                                                        \
    if (e_ != cudaSuccess) return fail(-2, "launch failed: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

static size_t xvec_bytes_q(int quant, int n) {
  return (quant == DSK_Q2_K || quant == DSK_Q3_K) ? xvec_bytes<Q_Q2K>(n) : xvec_bytes<Q_F32>(n);
}

// ---------------------------------------------------------------------------------------------------
// persistent interpreter: build the per-state program (stages, tile plans, column pieces)
// ---------------------------------------------------------------------------------------------------
static int granules(int quant, int n) {
  switch (quant) { case DSK_F32: return n / 4; case DSK_F16: return n / 8; case DSK_F8E5M2: return n / 16; default: return n / 256; }
}
static bool kq_quant(int q) { return q == DSK_Q2_K || q == DSK_Q3_K; }

static int g_wp_rows = 16;
static int g_slot_data = kSlotData, g_slot_scale = kSlotScale;   // set per program: 16 KB tiles for the warp-per-tile tensor-core path

static int plan_gemv_stage(Stage& st, int quant, int G) {
  const size_t rb = dev_row_bytes(quant, st.n);
  const int parts = st.epi == EPI_GLU ? 2 : 1;
  if (quant == DSK_F8E5M2 && st.n % 64 == 0 && g_use_mma && g_f8_mma_ok) {
    // tensor-core tiles: 16 weight rows per mma row group (8 + 8 for the gate/up pair), K split into 64-column pieces
    st.use_mma = 1;
    int total_rows = 0;
    for (int j = 0; j < st.njobs; j++) total_rows += st.job[j].rows;
    int RT = parts == 2 ? g_wp_rows / 2 : g_wp_rows;
    while (RT > 1 && align_up((size_t)RT * rb, 128) * parts > (size_t)g_slot_data) RT >>= 1;
    st.rows_per_tile = RT;
    st.rpass = 1;
    st.wp = 1;   // one warp reduces the whole tile (rows at a padded pitch)
    const int rpg = parts == 2 ? 8 : 16, groups = cdiv(RT, rpg), gran = st.n / 64;
    int csplit = std::max(1, std::min(std::min(8 / std::max(1, groups), gran), 8));
    st.npieces = csplit;
    for (int cs = 0; cs < csplit; cs++) st.piece[cs] = Piece{0, (int)((long long)cs * gran / csplit), (int)((long long)(cs + 1) * gran / csplit), 0};
    int t = 0;
    st.has_dyn = 0;
    for (int j = 0; j < st.njobs; j++) { st.job[j].tile_begin = t; t += cdiv(st.job[j].rows, RT); if (st.job[j].expert_slot >= 0) st.has_dyn = 1; }
    st.ntiles = t;
    // short stages (at most ~2 tiles per CTA of long rows): one warp per tile would leave 6-7 warps idle behind a single
    // 5 us tile, so all eight warps share each tile instead (column pieces, combined in a fixed order)
    if (g_coop_small && st.ntiles <= 2 * G && st.n >= 1024 && csplit > 1) st.wp = 0;
    return 0;
  }
  int total_rows = 0;
  for (int j = 0; j < st.njobs; j++) total_rows += st.job[j].rows;
  if (kq_quant(quant)) {
    if ((st.n / 256) * 4 > 32 * kKqMaxPass) return fail(-4, "K-quant row of %d columns exceeds the warp-per-tile limit (%d)", st.n, 8 * 256 * kKqMaxPass);
    // warp-per-tile K-quant tiles: up to 32 rows (lane r keeps row r), as many as fit one slot, TMA-aligned
    const int nb = st.n / 256;
    const int step = quant == DSK_Q2_K ? ((nb % 4 == 0) ? 1 : (nb % 2 == 0 ? 2 : 4)) : 1;
    int RT = (int)std::min<size_t>(32, ((size_t)g_slot_data / parts) / rb);
    RT = RT / std::max(step, 4) * std::max(step, 4);
    if (g_kq_small && RT >= 4) {   // short stages: enough tiles for every consumer warp of every CTA (tile >= 2 KB)
      const int unit = std::max(step, 4);
      int want = cdiv(cdiv(total_rows, G * 8), unit) * unit;
      while ((size_t)want * rb * parts < 2048) want += unit;
      RT = std::max(unit, std::min(RT, want));
    }
    if (RT >= 4) {
      st.wp = 1; st.rows_per_tile = RT; st.rpass = 1; st.npieces = 1;
      st.piece[0] = Piece{0, 0, nb, 0};
      int t = 0;
      st.has_dyn = 0;
      for (int j = 0; j < st.njobs; j++) { st.job[j].tile_begin = t; t += cdiv(st.job[j].rows, RT); if (st.job[j].expert_slot >= 0) st.has_dyn = 1; }
      st.ntiles = t;
      return 0;
    }
    return fail(-4, "K-quant row of %d columns (%zu bytes) does not fit a ring slot", st.n, rb);
  }
  int RT = 32;
  while (RT > 1 && align_up((size_t)RT * rb, 128) * parts > (size_t)g_slot_data) RT >>= 1;
  int min_rt = 1;
  if (quant == DSK_Q2_K) { const int nb = st.n / 256; min_rt = (nb % 4 == 0) ? 1 : (nb % 2 == 0 ? 2 : 4); }  // 16-byte TMA source alignment
  while (RT > std::max(min_rt, 4) && cdiv(total_rows, RT) < G) RT >>= 1;   // at least one tile per CTA when possible
  RT = std::max(RT, min_rt);
  st.rows_per_tile = RT;
  const bool kq = kq_quant(quant);
  int R = kq ? 1 : (RT >= 32 ? 4 : (RT >= 16 ? 2 : 1));
  if (parts == 2) R = std::min(R, 2);
  st.rpass = R;
  const int groups = cdiv(RT, R), gran = granules(quant, st.n);
  int csplit = groups >= 8 ? 1 : 8 / groups;
  csplit = std::max(1, std::min(csplit, std::max(1, gran / 2)));
  csplit = std::min(csplit, 8);
  st.npieces = csplit;
  for (int cs = 0; cs < csplit; cs++) st.piece[cs] = Piece{0, (int)((long long)cs * gran / csplit), (int)((long long)(cs + 1) * gran / csplit), 0};
  int t = 0;
  st.has_dyn = 0;
  for (int j = 0; j < st.njobs; j++) { st.job[j].tile_begin = t; t += cdiv(st.job[j].rows, RT); if (st.job[j].expert_slot >= 0) st.has_dyn = 1; }
  st.ntiles = t;
  return 0;
}

// partial-sum buffers of a warp-per-tile DOWN stage: with at most W consecutive pieces outstanding in the ring, up to
// (W - 2) / np + 2 row groups are in flight; 512 floats of shared memory hold their partial sums ([buffer][piece][16 rows]).
// Picks the largest window (16, 8, 4) whose buffers fit; returns 0 if none does.
static int down_window(int np, int* nbuf) {
  for (int W = kRingEntries; W >= 4; W >>= 1) {
    const int need = (W - 2) / np + 2;
    int nb = 2;
    while (nb < need) nb <<= 1;
    if (nb <= 16 && nb * np * 16 <= 512) { *nbuf = nb; return W; }
  }
  return 0;
}

static int plan_down_stage(Stage& st, int quant, int dim) {
  const size_t rb_mi = dev_row_bytes(quant, st.mi), rb_sh = dev_row_bytes(quant, st.sh);
  if (quant == DSK_F8E5M2 && st.mi % 64 == 0 && st.sh % 64 == 0 && g_use_mma && g_f8_mma_ok) {
    // warp-per-tile pieces: (segment, rows [g0, g0+g1) of an 8-row output group), whole rows, <= one slot each
    st.use_mma = 1; st.wp = 1; st.down_rows = g_wp_rows; st.rows_per_tile = g_wp_rows; st.seg_stride = 0;
    int np = 0;
    for (int k = 0; k <= st.K; k++) {
      const int n = k < st.K ? st.mi : st.sh;
      if (n == 0) continue;
      int pr = g_wp_rows;
      while (pr > 1 && (size_t)pr * f8_pitch((size_t)n) > (size_t)g_slot_data) pr >>= 1;
      if ((size_t)pr * f8_pitch((size_t)n) > (size_t)g_slot_data) return fail(-4, "down-projection row (%d bytes) does not fit a ring slot", n);
      for (int r0 = 0; r0 < g_wp_rows; r0 += pr) {
        if (np >= 16) return fail(-4, "too many down-projection pieces");
        st.piece[np++] = Piece{k, r0, pr, 0};
      }
    }
    st.npieces = np;
    st.ntiles = cdiv(dim, g_wp_rows) * np;
    st.max_inflight = down_window(np, &st.down_nbuf);
    if (!st.max_inflight) return fail(-4, "down projection: %d pieces per row group exceed the partial-sum buffers", np);
    return 0;
  }
  if (kq_quant(quant)) {
    if ((std::max(st.mi, st.sh) / 256) * 4 > 32 * kKqMaxPass) return fail(-4, "K-quant down-projection row exceeds the warp-per-tile limit");
    // K-quants: one tile per row group = DR output rows x all K + 1 segments, reduced by one warp (see kq_down_group).
    // DR = the smallest of {16, 8, 4} that leaves every CTA at most 8 row groups (one per consumer warp: a single round),
    // as long as the tile fits the ring comfortably.
    const int G = g_sm_count;
    const size_t rb_mi = dev_row_bytes(quant, st.mi), rb_sh = dev_row_bytes(quant, st.sh);
    int DR = 0;
    for (int cand : {4, 8, 16}) {
      if (cdiv(cdiv(dim, cand), G) > 8 && cand != 16) continue;
      DR = cand;
      break;
    }
    if (const char* e = getenv("DSK_DOWN_ROWS")) { const int v = atoi(e); if (v == 4 || v == 8 || v == 16) DR = v; }
    while (DR > 4 && (size_t)DR * (rb_mi * st.K + rb_sh) > 64 * 1024) DR >>= 1;
    if ((size_t)DR * (rb_mi * st.K + rb_sh) <= 64 * 1024) {
      st.wp = 1; st.down_rows = DR; st.rows_per_tile = DR; st.npieces = 1;
      st.seg_stride = (int)((size_t)DR * rb_mi);          // 16-byte multiple for every DR >= 4 (84- / 112-byte blocks)
      st.ntiles = cdiv(dim, DR);
      st.max_inflight = kRingEntries; st.down_nbuf = 1;
      return 0;
    }
    return fail(-4, "K-quant down-projection rows do not fit a ring slot");
  }
  int RT = 8;
  auto bytes = [&](int r) { return align_up((size_t)r * rb_mi, 128) * st.K + align_up((size_t)r * rb_sh, 128); };
  while (RT > 1 && bytes(RT) > (size_t)g_slot_data) RT >>= 1;
  if (bytes(RT) > (size_t)g_slot_data) return fail(-4, "down-projection row (%zu bytes) does not fit a ring slot", bytes(1));
  if (quant == DSK_Q2_K) {
    auto ok = [&](int n) { const int nb = n / 256; return n == 0 || (RT * nb) % 4 == 0; };
    if (!ok(st.mi) || !ok(st.sh)) return fail(-4, "Q2_K down projection: tile rows x blocks not 16-byte aligned");
  }
  st.rows_per_tile = RT;
  st.seg_stride = (int)align_up((size_t)RT * rb_mi, 128);
  st.ntiles = cdiv(dim, RT);
  st.use_mma = 0;
  const int g_mi = st.mi ? (st.use_mma ? st.mi / 64 : granules(quant, st.mi)) : 0, g_sh = st.sh ? (st.use_mma ? st.sh / 64 : granules(quant, st.sh)) : 0;
  const long long total = (long long)g_mi * st.K + g_sh;
  const int want = st.use_mma ? 8 : std::max(1, cdiv(16, RT));
  const long long L = std::max<long long>(1, cdiv((int)total, want));
  int np = 0;
  for (int k = 0; k <= st.K; k++) {
    const int g = k < st.K ? g_mi : g_sh;
    if (g == 0) continue;
    int cs = (int)std::max<long long>(1, (g + L / 2) / L);
    cs = std::min(cs, g);
    while (np + cs + (st.K - k) > kMaxPieces && cs > 1) cs--;
    for (int c = 0; c < cs; c++) st.piece[np++] = Piece{k, (int)((long long)c * g / cs), (int)((long long)(c + 1) * g / cs), 0};
  }
  st.npieces = np;
  return 0;
}


static MJob mjob(const DTensor& t, float* out) {
  MJob j{};
  j.w = t.w; j.scale = t.scale; j.out = out; j.rows = t.rows; j.expert_slot = -1;
  return j;
}

// Ring-slot payload size of a program, chosen before its stages are planned: the tensor-core F8 path wants 16 rows of
// 2048 + 64 bytes, K-quant tiles at least four rows of the longest row any stage streams (a 4-row interleave is the unit of
// kq_tile_rows), and a quantised MoE model one whole F32 gate row.
static void choose_slot_geometry(int q, const std::vector<Stage>& S, int gate_dim) {
  const bool wp_model = q == DSK_F8E5M2 && g_use_mma && g_f8_mma_ok;
  const bool kq_model = kq_quant(q);
  g_wp_rows = getenv("DSK_WP_ROWS") ? atoi(getenv("DSK_WP_ROWS")) : 16;
  if (g_wp_rows != 8) g_wp_rows = 16;
  g_slot_data = wp_model ? (g_wp_rows == 16 ? 33 * 1024 + 256 : 16 * 1024 + 512) : (kq_model ? 16 * 1024 + 512 : kSlotData);
  if (kq_model) {
    size_t need = 0;
    for (const Stage& st : S) {
      if (st.kind == ST_GEMV && kq_quant(st.quant)) need = std::max(need, 4 * dev_row_bytes(st.quant, st.n) * (st.epi == EPI_GLU ? 2 : 1));
      else if (st.kind == ST_DOWN) need = std::max(need, 4 * dev_row_bytes(st.quant, std::max(st.mi, st.sh)));
    }
    if (need <= 48 * 1024) g_slot_data = std::max(g_slot_data, (int)align_up(need, 128));   // longer rows: the generic (cooperative) path
  }
  if (gate_dim > 0) g_slot_data = std::max(g_slot_data, (int)align_up((size_t)gate_dim * 4, 128));   // a whole F32 gate row per slot
  g_slot_scale = kSlotScale;
}

struct Geometry { int ring_bytes = 0, slot_data = 0, slot_scale = 0, slot_bytes = 0; size_t xreg = 0, smem = 0; };

// shared-memory budget of a planned stage list: activation region = max over stages, the rest is ring slots
static int program_geometry(const std::vector<Stage>& S, int q, int hd, int max_seq, int bs1_cfg, Geometry* g) {
  size_t xreg = 8192;
  bool has_attn = false;
  for (const Stage& st : S) {
    if (st.kind == ST_GEMV) xreg = std::max(xreg, st.use_mma ? x16_bytes(st.n) : xvec_bytes_q(st.quant, st.n));
    else if (st.kind == ST_DOWN) {
      size_t b = 0;
      if (st.use_mma) b = down_x16_bytes(st.K, st.mi, st.sh);
      else if (kq_quant(st.quant)) b = down_q8_bytes<Q_Q2K>(st.K, st.mi, st.sh);
      else for (int k = 0; k <= st.K; k++) { const int n = k < st.K ? st.mi : st.sh; if (n) b += xvec_bytes_q(st.quant, n); }
      xreg = std::max(xreg, b);
    } else if (st.kind == ST_ATTN) has_attn = true;
    else if (st.kind == ST_MLA_CACHE) xreg = std::max(xreg, (size_t)4096);
    else if (st.kind == ST_ATTN_MLA) {   // fp32 scratch (scores stay on chip) + the head's wv_b slab and staged latent
      const int kvl = st.n, vh = st.job[0].rows;
      size_t b = align_up(mla_floats(kvl, st.K /* rope dim */, max_seq) * 4, 128);
      if (st.quant == DSK_Q2_K) b += mla_weight_bytes<Q_Q2K>(kvl, vh);
      else if (st.quant == DSK_Q3_K) b += mla_weight_bytes<Q_Q3K>(kvl, vh);
      else if (st.quant == DSK_F8E5M2 && st.use_mma) b += mla_weight_bytes<Q_F8>(kvl, vh);
      xreg = std::max(xreg, b);
    }
  }
  if (has_attn) {
    const size_t attn_need = (size_t)(512 + ((hd + 3) & ~3) + ((max_seq + 3) & ~3) + kConsumers + 16) * 4;
    if (attn_need <= 64 * 1024) xreg = std::max(xreg, attn_need);
    xreg = std::max(xreg, (size_t)(512 + ((hd + 3) & ~3) + 64) * 4);
  }
  xreg = align_up(xreg, 128);
  const size_t budget = (size_t)kSmemMax - 2048;
  {  // scale-row area of a ring slot: as small as this program's f8 scale rows allow (a 5th 34 KB slot fits for V2-Lite)
    size_t need = 0;
    bool generic = false;
    const int bs1 = bs1_cfg > 0 ? bs1_cfg : 1;
    for (const Stage& st : S) {
      if (st.kind == ST_GEMV && st.quant == DSK_F8E5M2) need = std::max(need, (size_t)(cdiv(st.n, bs1) * 4 + 16) * (st.epi == EPI_GLU ? 2 : 1));
      else if (st.kind == ST_DOWN && st.quant == DSK_F8E5M2) {
        if (!st.wp) generic = true;
        need = std::max(need, (size_t)(cdiv(std::max(st.mi, st.sh), bs1) * 4 + 16));
      }
    }
    if (q != DSK_F8E5M2) g_slot_scale = 128;                       // no scale rows at all (F32 / F16 / K-quants)
    else if (!generic && need <= 512) g_slot_scale = 512;
    else g_slot_scale = kSlotScale;
  }
  const size_t slot_bytes = (size_t)g_slot_data + g_slot_scale;
  size_t max_tile = align_up(slot_bytes, 128);
  for (const Stage& st : S)
    if (st.kind == ST_DOWN && st.wp && kq_quant(st.quant))
      max_tile = std::max(max_tile, align_up((size_t)st.K * st.seg_stride + (size_t)st.down_rows * dev_row_bytes(st.quant, st.sh), 128));
  if (kMegaHdr + xreg + 2 * max_tile > budget) return fail(-4, "activation staging (%zu bytes) leaves no room for the TMA ring", xreg);
  g->ring_bytes = (int)((budget - kMegaHdr - xreg) / 128 * 128);   // everything left: tiles take what they need out of it
  g->slot_data = g_slot_data; g->slot_scale = g_slot_scale; g->slot_bytes = (int)slot_bytes;
  g->xreg = xreg;
  g->smem = kMegaHdr + xreg + (size_t)g->ring_bytes;
  // the grid barrier needs every CTA resident: verify that one CTA of this footprint fits an SM (the launch is cooperative,
  // so anything less would be a launch error, not a hang)
  int nb = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, decode_kernel_for(q), kMegaThreads, g->smem));
  if (nb < 1) return fail(-4, "decode kernel (%zu bytes of shared memory) does not fit an SM", g->smem);
  return 0;
}

static void plan_stages(std::vector<Stage>& S, int q, int dim, int G, int* err) {
  *err = 0;
  for (Stage& st : S) {
    if (st.max_inflight == 0) st.max_inflight = kRingEntries;
    if (st.kind == ST_GEMV && st.ntiles == 0) { if (plan_gemv_stage(st, st.quant, G)) { *err = -4; return; } }
    else if (st.kind == ST_DOWN && st.ntiles == 0) { if (plan_down_stage(st, q, dim)) { *err = -4; return; } }
  }
}

static void fill_program_header(Program* P, const dsk_config& c, int hd, const Geometry& g) {
  P->dim = c.dim; P->n_heads = c.n_heads; P->hd = hd; P->nope = c.qk_nope_head_dim; P->rope = c.qk_rope_head_dim; P->vh = c.v_head_dim;
  P->kv_lora = c.kv_lora_rank; P->is_v3 = c.is_v3; P->bs0 = c.bs0 > 0 ? c.bs0 : 1; P->bs1 = c.bs1 > 0 ? c.bs1 : 1;
  P->act_silu = c.act_silu; P->max_seq = c.max_seq_len;
  { int b1 = P->bs1, sh = 0; while ((1 << sh) < b1) sh++; P->bs1_shift = ((1 << sh) == b1) ? sh : -1; }
  P->E = c.n_routed_experts; P->K = c.n_active_routed;
  P->norm_topk_prob = c.norm_topk_prob; P->sigmoid = c.scoring_sigmoid; P->topk_method = c.topk_method;
  P->n_group = std::max(1, c.n_group); P->topk_group = c.topk_group; P->original_max = c.original_max_position;
  P->eps = c.norm_eps; P->routed_scale = c.routed_scaling_factor;
  P->ring_bytes = g.ring_bytes; P->xregion_bytes = (int)g.xreg;
  P->slot_data = g.slot_data; P->slot_scale = g.slot_scale; P->slot_bytes = g.slot_bytes;
  P->n_ranks = 1; P->rank = 0;
}

static int build_program(dsk_model* m, dsk_state* s) {
  const dsk_config& c = m->c;
  const int hd = m->head_dim, mi = c.moe_intermediate_size, q = c.quant, G = g_sm_count;
  const bool tp = m->tp && m->p2p;   // slices are decided at model creation; without the peer mapping dsk_comm_init() refuses
  g_f8_mma_ok = c.bs1 > 0 && (c.bs1 & (c.bs1 - 1)) == 0;
  std::vector<Stage> S;
  int n_xchg = 0;
  s->cut_after.clear();
  auto gemv = [&](int quant, const float* in, const float* norm_w, int n, int epi, int layer) {
    Stage st{};
    st.kind = ST_GEMV; st.quant = quant; st.epi = epi; st.in = in; st.norm_w = norm_w; st.n = n; st.layer = layer;
    return st;
  };
  s->layer_begin.assign(c.n_layers, 0);
  s->layer_end.assign(c.n_layers, 0);
  { Stage st{}; st.kind = ST_EMBED; st.quant = q; S.push_back(st); }
  for (int l = 0; l < c.n_layers; l++) {
    Layer& L = m->layers[l];
    s->layer_begin[l] = (int)S.size();
    {  // S1
      Stage st = gemv(q, s->x, L.rms_att, c.dim, EPI_STORE, l);
      st.job[0] = c.q_lora_rank > 0 ? mjob(L.wq_a, s->q_a) : mjob(L.wq, s->q);
      st.job[1] = mjob(L.wkv_a, s->kv_a);
      st.njobs = 2;
      S.push_back(st);
    }
    if (c.use_mla) {
      // BlockMLA (src/infer.cpp:1051-1141): cache update | q_rope = wq_rope_b . q_a', q_c = wc . q_a' | per-head latent attention + wv_b
      { Stage st{}; st.kind = ST_MLA_CACHE; st.quant = q; st.layer = l; st.norm_w = L.rms_kv_a; st.kcache = L.kcache; st.vcache = L.vcache; S.push_back(st); }
      {
        Stage st = gemv(q, s->q_a, L.rms_q_a, c.q_lora_rank, EPI_STORE, l);
        st.job[0] = mjob(L.wq_rope_b, s->q); st.job[1] = mjob(L.wc, s->q_c); st.njobs = 2;
        S.push_back(st);
      }
      {
        Stage st{}; st.kind = ST_ATTN_MLA; st.quant = q; st.layer = l; st.kcache = L.kcache; st.vcache = L.vcache;
        MJob j{}; j.w = L.wv_b.w; j.scale = L.wv_b.scale; j.rows = c.v_head_dim; j.expert_slot = -1;
        j.w_stride = (long long)((size_t)c.v_head_dim * L.wv_b.row_bytes);                       // bytes per head
        j.s_stride = (long long)((size_t)cdiv(c.v_head_dim, c.bs0 > 0 ? c.bs0 : 1) * cdiv(c.kv_lora_rank, c.bs1 > 0 ? c.bs1 : 1));   // scale floats per head (matmul_expert)
        st.job[0] = j; st.njobs = 1; st.n = c.kv_lora_rank; st.K = c.qk_rope_head_dim;
        st.use_mma = (q == DSK_F8E5M2 && g_use_mma && g_f8_mma_ok && c.kv_lora_rank % 64 == 0 && c.v_head_dim % 8 == 0 &&
                      c.bs0 % 16 == 0 && j.s_stride <= 256) ? 1 : 0;
        S.push_back(st);
      }
      {  // wo on the up-projected values (src/infer.cpp:1140)
        Stage st = gemv(q, s->kv_b, nullptr, c.n_heads * c.v_head_dim, EPI_RESID, l);
        st.job[0] = mjob(L.wo, s->x); st.njobs = 1;
        S.push_back(st);
      }
    } else {
    if (c.q_lora_rank > 0) {
      Stage st = gemv(q, s->q_a, L.rms_q_a, c.q_lora_rank, EPI_STORE, l);
      st.job[0] = mjob(L.wq_b, s->q); st.njobs = 1;
      S.push_back(st);
    }
    {  // S2
      Stage st = gemv(q, s->kv_a, L.rms_kv_a, c.kv_lora_rank, EPI_KVB, l);
      st.job[0] = mjob(L.wkv_b, s->kv_b); st.njobs = 1; st.kcache = L.kcache; st.vcache = L.vcache;
      S.push_back(st);
    }
    { Stage st{}; st.kind = ST_ATTN; st.quant = q; st.layer = l; st.kcache = L.kcache; st.vcache = L.vcache; S.push_back(st); }
    {  // S4: x += wo . xb2.  Tensor parallel: this rank holds the wo columns of its heads -> partial sum, exchanged right after
      Stage st = gemv(q, s->xb2, nullptr, m->nh_loc * c.v_head_dim, tp ? EPI_PARTIAL : EPI_RESID, l);
      st.job[0] = mjob(L.wo, s->x); st.njobs = 1;
      st.xchg_ord = n_xchg;
      S.push_back(st);
      if (tp) { Stage xs{}; xs.kind = ST_XCHG; xs.quant = q; xs.layer = l; xs.xchg_ord = n_xchg++; S.push_back(xs); }
    }
    }   // (MHA blocks)
    if (L.is_moe) {
      const int sh = m->sh_loc;   // (local slice of the concatenated shared experts; all of it without tensor parallelism)
      {  // S5 gate logits (F32 weights in every quant)
        Stage st = gemv(DSK_F32, s->x, L.rms_ffn, c.dim, EPI_STORE, l);
        MJob j{}; j.w = (const uint8_t*)L.gate; j.out = s->moe_logits; j.rows = c.n_routed_experts; j.expert_slot = -1;
        st.job[0] = j; st.njobs = 1;
        if (q != DSK_F32) {
          // quantised model: the dedicated gate stage (gate_f32_stage) takes one row per tile
          st.rows_per_tile = 1; st.rpass = 1; st.npieces = 1; st.wp = 0; st.use_mma = 0;
          st.piece[0] = Piece{0, 0, c.dim / 4, 0};
          st.job[0].tile_begin = 0; st.ntiles = c.n_routed_experts; st.has_dyn = 0;
        }
        S.push_back(st);
      }
      {  // S56: routing + shared (static, first: streams before the routing is known) + routed experts
        Stage st = gemv(q, s->x, L.rms_ffn, c.dim, EPI_GLU, l);
        st.need_topk = 1; st.gate_logits = s->moe_logits; st.gate_bias = L.gate_bias;
        int nj = 0;
        if (sh > 0 && L.sw1.rows > 0) { MJob j = mjob(L.sw1, s->hbs); j.w_b = L.sw3.w; j.scale_b = L.sw3.scale; st.job[nj++] = j; }
        for (int k = 0; k < c.n_active_routed; k++) {
          MJob j{};
          j.w = L.w1.w; j.scale = L.w1.scale; j.w_b = L.w3.w; j.scale_b = L.w3.scale;
          j.out = s->hbk + (size_t)k * mi; j.rows = mi; j.expert_slot = k;
          j.w_stride = (long long)L.w1.expert_bytes; j.s_stride = (long long)L.w1.scale_expert;
          st.job[nj++] = j;
        }
        st.njobs = nj;
        S.push_back(st);
      }
      {  // S7
        Stage st{};
        st.kind = ST_DOWN; st.quant = q; st.layer = l;
        st.w2 = L.w2.w; st.s2 = L.w2.scale; st.w2_stride = (long long)L.w2.expert_bytes; st.s2_stride = (long long)L.w2.scale_expert;
        st.sw2 = sh > 0 ? L.sw2.w : nullptr; st.ss2 = sh > 0 ? L.sw2.scale : nullptr;   // (sh == 0: this rank holds no slice)
        st.K = c.n_active_routed; st.mi = mi; st.sh = sh;
        st.add_shared = (m->n_ranks == 1 || m->rank == 0 || tp) ? 1 : 0;   // tp: every rank adds its slice of the shared experts
        st.xchg_ord = n_xchg;
        S.push_back(st);
        if (m->n_ranks > 1 && m->p2p) {   // in-kernel exchange over peer memory: the token stays ONE kernel
          Stage xs{};
          xs.kind = ST_XCHG; xs.quant = q; xs.layer = l; xs.xchg_ord = n_xchg++;
          S.push_back(xs);
        } else if (m->n_ranks > 1) s->cut_after.push_back((int)S.size() - 1);
      }
    } else {
      {
        Stage st = gemv(q, s->x, L.rms_ffn, c.dim, EPI_GLU, l);
        MJob j = mjob(L.w1, s->hbs); j.w_b = L.w3.w; j.scale_b = L.w3.scale;
        st.job[0] = j; st.njobs = 1;
        S.push_back(st);
      }
      {
        Stage st{};
        st.kind = ST_DOWN; st.quant = q; st.layer = l;
        st.sw2 = L.w2.w; st.ss2 = L.w2.scale; st.K = 0; st.mi = 0; st.sh = m->hid_loc; st.add_shared = 1;
        st.xchg_ord = n_xchg;
        S.push_back(st);
        if (tp) { Stage xs{}; xs.kind = ST_XCHG; xs.quant = q; xs.layer = l; xs.xchg_ord = n_xchg++; S.push_back(xs); }
      }
    }
    s->layer_end[l] = (int)S.size();
  }
  s->logits_src = (tp && m->logits_full) ? m->logits_full : s->logits;
  s->n_tail = 1;
  {  // LM head + argmax (tensor parallel: this rank's rows; logits and arg-max keys are exchanged by the ST_AMAX stage)
    Stage st = gemv(q, s->x, m->rms_final, c.dim, EPI_LOGITS, -1);
    st.job[0] = mjob(m->wcls, s->logits_src + m->v0); st.job[0].row_base = m->v0; st.njobs = 1;
    S.push_back(st);
    if (tp) { Stage xs{}; xs.kind = ST_AMAX; xs.quant = q; xs.layer = -1; xs.xchg_ord = n_xchg++; S.push_back(xs); s->n_tail = 2; }
  }
  if (m->n_ranks > 1 && m->p2p)
    for (Stage& st : S) {
      st.peer_stores = (st.kind == ST_GEMV && st.epi == EPI_LOGITS && tp) ? 1 : 0;   // (partial sums go through the local vector)
    }
  const bool has_gate = q != DSK_F32 && c.n_routed_experts > 0;
  choose_slot_geometry(q, S, has_gate ? c.dim : 0);
  if (has_gate && (size_t)c.dim * 4 > (size_t)g_slot_data) return fail(-4, "gate row (%d bytes) does not fit a ring slot", c.dim * 4);
  int perr = 0;
  plan_stages(S, q, c.dim, G, &perr);
  if (perr) return perr;
  Geometry geo;
  if (program_geometry(S, q, hd, c.max_seq_len, c.bs1, &geo)) return -4;
  s->mega_smem = geo.smem;
  s->n_stages = (int)S.size();

  std::vector<unsigned char> buf(sizeof(Program) + (S.size() - 1) * sizeof(Stage));
  Program* P = reinterpret_cast<Program*>(buf.data());
  memset(P, 0, sizeof(Program));
  fill_program_header(P, c, hd, geo);
  P->n_heads = m->nh_loc;
  P->expert_first = m->expert_first; P->expert_count = m->expert_count;
  P->embed_quant = q; P->n_stages = (int)S.size();
  P->embed_w = m->embed.w; P->embed_scale = m->embed.scale; P->rope_freq = m->rope_freq;
  P->q_c = s->q_c;
  P->x = s->x; P->q = s->q; P->q_a = s->q_a; P->kv_a = s->kv_a; P->kv_b = s->kv_b; P->xb2 = s->xb2; P->hbk = s->hbk; P->hbs = s->hbs;
  P->moe_logits = s->moe_logits; P->moe_scores = s->moe_scores; P->act_w = s->act_w; P->logits = s->logits_src; P->partial = m->n_ranks > 1 ? s->partial : nullptr;
  P->act = s->act; P->ctrl = s->ctrl; P->token_log = s->token_log; P->step = s->step;
  CK(cudaMalloc((void**)&s->att_scratch, (size_t)m->nh_loc * (c.max_seq_len + kConsumers + 8) * 4));
  CK(cudaMalloc((void**)&s->sync_words, 64));
  CK(cudaMemset(s->sync_words, 0, 64));
  P->att_scratch = s->att_scratch; P->sync_counter = s->sync_words; P->sync_base = s->sync_words + 1;
  CK(cudaMalloc((void**)&s->tstamp, (S.size() * 8 + 8) * sizeof(unsigned long long)));
  CK(cudaMemset(s->tstamp, 0, (S.size() * 8 + 8) * sizeof(unsigned long long)));
  P->n_ranks = m->n_ranks; P->rank = m->rank; P->n_xchg = n_xchg; P->tp = tp ? 1 : 0;
  for (int qq = 0; qq < kMaxRanks; qq++) {
    P->xchg_peer[qq] = m->xchg_peer[qq]; P->xflag_peer[qq] = m->xflag_peer[qq];
    P->logits_peer[qq] = m->logits_peer[qq]; P->amax_peer[qq] = m->amax_peer[qq];
  }
  s->n_xchg = n_xchg;
  s->xchg_before.assign(S.size() + 1, 0);
  for (size_t i = 0; i < S.size(); i++) s->xchg_before[i + 1] = s->xchg_before[i] + ((S[i].kind == ST_XCHG || S[i].kind == ST_AMAX) ? 1 : 0);
  s->built_epoch = m->p2p_epoch;
  P->tstamp = getenv("DSK_TSTAMP") ? s->tstamp : nullptr;   // the per-stage timeline costs a few globaltimer reads per stage: opt-in
  P->route_prof = reinterpret_cast<long long*>(s->tstamp + S.size() * 8);
  s->stage_names.clear();
  for (const Stage& st : S) {
    char nm[96];
    const char* kind = st.kind == ST_EMBED ? "embed" : st.kind == ST_XCHG ? "xchg" : st.kind == ST_AMAX ? "amax" : st.kind == ST_MLA_CACHE ? "mla_kv" : st.kind == ST_ATTN_MLA ? "attn_mla" : st.kind == ST_ATTN ? "attn" : st.kind == ST_DOWN ? "down" : (st.epi == EPI_GLU ? "glu" : st.epi == EPI_KVB ? "kv_b" : st.epi == EPI_RESID ? "wo" : st.epi == EPI_LOGITS ? "lm_head" : (st.quant == DSK_F32 && q != DSK_F32 ? "gate" : "proj"));
    snprintf(nm, sizeof(nm), "%-8s n=%5d tiles=%5d rt=%2d wp=%d pieces=%2d", kind, st.kind == ST_DOWN ? st.K * st.mi + st.sh : st.n, st.ntiles, st.rows_per_tile, st.wp, st.npieces);
    s->stage_names.push_back(nm);
  }
  memcpy(P->stage, S.data(), S.size() * sizeof(Stage));
  CK(cudaMalloc((void**)&s->prog, buf.size()));
  CK(cudaMemcpy(s->prog, buf.data(), buf.size(), cudaMemcpyHostToDevice));
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// launches: ONE cooperative launch runs stages [b, e) for n_tokens tokens (cudaLaunchAttributeCooperative: all g_sm_count
// CTAs are guaranteed co-resident, which the in-kernel grid barrier relies on — a busy device yields a launch error)
// ---------------------------------------------------------------------------------------------------
static int g_launch_count = 0;

static cudaError_t launch_decode_raw(int quant, const Program* prog, size_t smem, int s_begin, int s_end, int from_argmax,
                                     int n_tokens, cudaStream_t st) {
  if (s_end <= s_begin || n_tokens <= 0) return cudaSuccess;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)g_sm_count);
  cfg.blockDim = dim3((unsigned)kMegaThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;
  at[0].val.cooperative = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  g_launch_count++;
  return cudaLaunchKernelEx(&cfg, decode_kernel_for(quant), prog, s_begin, s_end, from_argmax, n_tokens);
}
static cudaError_t launch_decode(dsk_model* m, dsk_state* s, int s_begin, int s_end, int from_argmax, int n_tokens, cudaStream_t st) {
  return launch_decode_raw(m->c.quant, s->prog, s->mega_smem, s_begin, s_end, from_argmax, n_tokens, st);
}

// The state (and its program) may have been created before dsk_comm_init() mapped the peers: rebuild it once.
static int ensure_program(dsk_model* m, dsk_state* s) {
  if (s->built_epoch == m->p2p_epoch) return 0;
  CK(cudaStreamSynchronize(s->stream));
  cudaFree(s->prog); cudaFree(s->att_scratch); cudaFree(s->sync_words); cudaFree(s->tstamp);
  s->prog = nullptr; s->att_scratch = nullptr; s->sync_words = nullptr; s->tstamp = nullptr;
  s->stage_names.clear(); s->layer_begin.clear(); s->layer_end.clear();
  return build_program(m, s);
}
// Ctrl::pad[0] = exchanges completed before the token the launch [b, e) belongs to (peer-memory mode)
static void set_xchg_base(dsk_model* m, dsk_state* s, int b, int e) {
  if (s->n_xchg == 0) return;
  s->h_ctrl->pad[0] = (int)(m->xchg_done - s->xchg_before[b]);
  m->xchg_done += s->xchg_before[e] - s->xchg_before[b];
}

// stages [b, e) of ONE token: one resident grid (ENG_MEGA) or one launch per stage (ENG_STAGE); the NCCL fallback of the
// multi-GPU path cuts at the all-reduce points
static int run_stages(dsk_model* m, dsk_state* s, int b, int e, int from_argmax, cudaStream_t st) {
  const dsk_config& c = m->c;
  int cur = b;
  auto flush = [&](int upto) -> int {
    if (g_engine == ENG_STAGE) { for (int i = cur; i < upto; i++) CKL(launch_decode(m, s, i, i + 1, from_argmax, 1, st)); }
    else CKL(launch_decode(m, s, cur, upto, from_argmax, 1, st));
    cur = upto;
    return 0;
  };
  for (int cut : s->cut_after) {
    if (cut < b || cut >= e) continue;
    if (flush(cut + 1)) return -2;
    if (!m->comm) return fail(-3, "n_ranks > 1 but dsk_comm_init() was not called");
    CKN(g_nccl.AllReduce(s->partial, s->partial, c.dim, ncclFloat, ncclSum, m->comm, st));
    add_vec_kernel<<<cdiv(c.dim, 256), 256, 0, st>>>(s->x, s->partial, c.dim);
    g_launch_count += 2;
    CKL(cudaGetLastError());
  }
  return flush(e);
}

static void fill_ctrl(Ctrl* h, const dsk_config& c, int token, int pos) {
  // src/infer.cpp:1274-1277
  const int omp = c.original_max_position;
  h->token = token;
  h->pos = pos;
  h->kv_sink = pos >= omp ? 2 : 0;
  h->kv_pos = h->kv_sink + (pos - h->kv_sink) % (omp - h->kv_sink);
  h->kv_len = pos >= omp ? omp : pos + 1;
  h->argmax_key = 0ull;
}

extern "C" int dsk_forward(dsk_model* m, dsk_state* s, int token, int pos, int mode, float* host_logits, int* argmax) {
  if (need_device()) return -1;
  if (!m || !s) return fail(-1, "null model/state");
  const dsk_config& c = m->c;
  if (token < 0 || token >= c.vocab_size) return fail(-4, "token %d out of range", token);
  if (pos < 0) return fail(-4, "negative pos");
  mode = mode ? 1 : 0;
  if (m->n_ranks > 1 && !m->comm) return fail(-3, "n_ranks > 1 but dsk_comm_init() was not called");
  if (ensure_program(m, s)) return -2;
  fill_ctrl(s->h_ctrl, c, token, pos);
  if (s->h_ctrl->kv_pos >= c.max_seq_len || s->h_ctrl->kv_len > c.max_seq_len)
    return fail(-4, "pos %d does not fit the KV cache (max_seq_len %d; the reference would overrun it)", pos, c.max_seq_len);
  const int e = mode == DSK_HYDRATE_KV_CACHE ? s->n_stages - s->n_tail : s->n_stages;
  set_xchg_base(m, s, 0, e);
  CK(cudaMemcpyAsync(s->ctrl, s->h_ctrl, sizeof(Ctrl), cudaMemcpyHostToDevice, s->stream));
  if (run_stages(m, s, 0, e, 0, s->stream)) return -2;
  if (mode && host_logits) CK(cudaMemcpyAsync(host_logits, s->logits_src, (size_t)c.vocab_size * 4, cudaMemcpyDeviceToHost, s->stream));
  if (mode && argmax) CK(cudaMemcpyAsync(s->h_ctrl, s->ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, s->stream));
  CK(cudaStreamSynchronize(s->stream));
  if (mode && argmax) *argmax = (int)(0xFFFFFFFFu - (unsigned)(s->h_ctrl->argmax_key & 0xFFFFFFFFull));
  s->last_pos = pos;
  s->have_logits = mode != 0;
  return 0;
}

extern "C" int dsk_copy_embedding(dsk_model* m, dsk_state* s, int token) {
  if (need_device()) return -1;
  if (!m || !s) return fail(-1, "null model/state");
  if (token < 0 || token >= m->c.vocab_size) return fail(-4, "token %d out of range", token);
  if (ensure_program(m, s)) return -2;
  fill_ctrl(s->h_ctrl, m->c, token, 0);
  set_xchg_base(m, s, 0, 1);
  CK(cudaMemcpyAsync(s->ctrl, s->h_ctrl, sizeof(Ctrl), cudaMemcpyHostToDevice, s->stream));
  if (run_stages(m, s, 0, 1, 0, s->stream)) return -2;
  CK(cudaStreamSynchronize(s->stream));
  s->have_logits = false;
  return 0;
}

extern "C" int dsk_block_forward(dsk_model* m, dsk_state* s, int layer, int pos, int kv_sink, int kv_pos, int kv_len) {
  if (need_device()) return -1;
  if (!m || !s) return fail(-1, "null model/state");
  const dsk_config& c = m->c;
  if (layer < 0 || layer >= c.n_layers) return fail(-4, "bad layer %d", layer);
  if (pos < 0 || kv_sink < 0 || kv_pos < 0 || kv_pos >= c.max_seq_len || kv_len < 1 || kv_len > c.max_seq_len || kv_sink > kv_len)
    return fail(-4, "block(pos %d, kv_sink %d, kv_pos %d, kv_len %d) does not fit the KV cache (max_seq_len %d)", pos, kv_sink, kv_pos, kv_len, c.max_seq_len);
  if (m->n_ranks > 1 && !m->comm) return fail(-3, "n_ranks > 1 but dsk_comm_init() was not called");
  if (ensure_program(m, s)) return -2;
  Ctrl* h = s->h_ctrl;
  h->token = 0; h->pos = pos; h->kv_sink = kv_sink; h->kv_pos = kv_pos; h->kv_len = kv_len; h->argmax_key = 0;
  set_xchg_base(m, s, s->layer_begin[layer], s->layer_end[layer]);
  CK(cudaMemcpyAsync(s->ctrl, h, sizeof(Ctrl), cudaMemcpyHostToDevice, s->stream));
  if (run_stages(m, s, s->layer_begin[layer], s->layer_end[layer], 0, s->stream)) return -2;
  CK(cudaStreamSynchronize(s->stream));
  s->have_logits = false;
  return 0;
}

extern "C" int dsk_decode_greedy(dsk_model* m, dsk_state* s, int start_pos, int n_steps, int32_t* out_tokens,
                                 float* elapsed_ms) {
  if (need_device()) return -1;
  if (!m || !s) return fail(-1, "null model/state");
  if (n_steps <= 0) return fail(-4, "n_steps must be positive");
  if (s->last_pos < 0 || start_pos != s->last_pos + 1)
    return fail(-4, "dsk_decode_greedy: start_pos %d must follow the last forward (pos %d)", start_pos, s->last_pos);
  // the first token is taken from the on-device arg-max of the previous LM-head stage: a hydrate-only forward, a block
  // call or an embedding copy leaves no such key behind (the reference's run_completion samples from s.logits() here)
  if (!s->have_logits)
    return fail(-4, "dsk_decode_greedy: the last forward did not produce logits (mode DSK_HYDRATE_KV_CACHE, block or embedding call) — run dsk_forward(..., DSK_OUTPUT_LOGITS) first");
  if ((size_t)n_steps > s->token_log_cap) return fail(-4, "n_steps too large");
  const dsk_config& c = m->c;
  if (c.original_max_position > c.max_seq_len && start_pos + n_steps > c.max_seq_len)
    return fail(-4, "decode would run past the KV cache (%d)", c.max_seq_len);
  if (m->n_ranks > 1 && !m->comm) return fail(-3, "n_ranks > 1 but dsk_comm_init() was not called");
  if (ensure_program(m, s)) return -2;
  if (s->n_xchg > 0) {   // the embedding stage of every token advances the exchange base by n_xchg
    s->h_ctrl->pad[0] = (int)(m->xchg_done - s->n_xchg);
    CK(cudaMemcpyAsync(&s->ctrl->pad[0], &s->h_ctrl->pad[0], sizeof(int), cudaMemcpyHostToDevice, s->stream));
    m->xchg_done += (long long)n_steps * s->n_xchg;
  }
  CK(cudaMemsetAsync(s->step, 0, sizeof(int), s->stream));
  CK(cudaEventRecord(s->ev0, s->stream));
  if (g_engine == ENG_MEGA && s->cut_after.empty()) {
    CKL(launch_decode(m, s, 0, s->n_stages, 1, n_steps, s->stream));   // the token loop runs inside the persistent kernel
  } else {
    for (int i = 0; i < n_steps; i++) if (run_stages(m, s, 0, s->n_stages, 1, s->stream)) return -2;
  }
  CK(cudaEventRecord(s->ev1, s->stream));
  CK(cudaStreamSynchronize(s->stream));
  if (elapsed_ms) CK(cudaEventElapsedTime(elapsed_ms, s->ev0, s->ev1));
  if (out_tokens) CK(cudaMemcpy(out_tokens, s->token_log, (size_t)n_steps * 4, cudaMemcpyDeviceToHost));
  s->last_pos = start_pos + n_steps - 1;
  s->have_logits = true;
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// on-device sampling — Sampler::sample / sample_prob (src/sampler.cpp:12-75) over the logits the LM-head stage left in HBM
// ---------------------------------------------------------------------------------------------------
// One CTA of 1024 threads; thread t owns the contiguous index range [t*chunk, (t+1)*chunk) so that the running
// probability sum is formed in vocabulary order like the reference's loop.  The reference accumulates in fp32 (its own
// result depends on the compiler's -ffast-math vectorisation of that loop); here the per-element probabilities are the same
// fp32 values expf((l - max) / T) / sum, accumulated in fp64, so the selected index can only differ from a given build of
// the reference when r lies within fp32 summation error of a bucket edge.
__global__ void __launch_bounds__(1024) sample_kernel(const float* __restrict__ logits, int vocab, float temperature, float r,
                                                       int index, float* __restrict__ out) {
  __shared__ double s_d[1024];
  __shared__ float s_f[32];
  __shared__ int s_who;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int chunk = (vocab + 1023) / 1024;
  const int i0 = min(vocab, tid * chunk), i1 = min(vocab, i0 + chunk);
  float mx = -3.402823466e38f;
  for (int i = i0; i < i1; i++) mx = fmaxf(mx, logits[i]);
  mx = warp_max(mx);
  if (lane == 0) s_f[warp] = mx;
  if (tid == 0) s_who = 1024;
  __syncthreads();
  mx = s_f[0];
  for (int w = 1; w < 32; w++) mx = fmaxf(mx, s_f[w]);
  double part = 0.0;
  for (int i = i0; i < i1; i++) part += (double)expf((logits[i] - mx) / temperature);
  s_d[tid] = part;
  __syncthreads();
  for (int o = 512; o; o >>= 1) { if (tid < o) s_d[tid] += s_d[tid + o]; __syncthreads(); }
  const float sum = (float)s_d[0];
  __syncthreads();
  if (index >= 0) {   // sample_prob
    if (tid == 0) { out[0] = __int_as_float(index); out[1] = expf((logits[index] - mx) / temperature) / sum; }
    return;
  }
  double loc = 0.0;
  for (int i = i0; i < i1; i++) loc += (double)(expf((logits[i] - mx) / temperature) / sum);
  s_d[tid] = loc;
  __syncthreads();
  // inclusive scan over the 1024 per-thread sums (Hillis-Steele in shared memory, fp64)
  for (int o = 1; o < 1024; o <<= 1) {
    const double add = tid >= o ? s_d[tid - o] : 0.0;
    __syncthreads();
    s_d[tid] += add;
    __syncthreads();
  }
  const double incl = s_d[tid], excl = incl - loc;
  if (incl >= (double)r && i1 > i0) atomicMin(&s_who, tid);
  __syncthreads();
  if (s_who == 1024) {   // the running sum never reaches r: the reference returns vocab_size - 1
    if (tid == 0) { out[0] = __int_as_float(vocab - 1); out[1] = 0.f; }
    return;
  }
  if (tid == s_who) {
    double cum = excl;
    int pick = i1 - 1;
    for (int i = i0; i < i1; i++) {
      cum += (double)(expf((logits[i] - mx) / temperature) / sum);
      if (cum >= (double)r) { pick = i; break; }
    }
    out[0] = __int_as_float(pick);
    out[1] = expf((logits[pick] - mx) / temperature) / sum;
  }
}

static int need_logits(dsk_state* s, const char* who) {
  if (!s->have_logits) return fail(-4, "%s: the last forward did not produce logits (run dsk_forward(..., DSK_OUTPUT_LOGITS) first)", who);
  return 0;
}

extern "C" int dsk_sample(dsk_model* m, dsk_state* s, float temperature, float top_p, float coin, int* token) {
  if (need_device()) return -1;
  if (!m || !s || !token) return fail(-1, "bad arguments");
  if (need_logits(s, "dsk_sample")) return -4;
  if (temperature == 0.0f) {   // Sampler::sample_argmax: the LM-head stage already reduced it (lowest index on ties)
    CK(cudaMemcpyAsync(s->h_ctrl, s->ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    *token = (int)(0xFFFFFFFFu - (unsigned)(s->h_ctrl->argmax_key & 0xFFFFFFFFull));
    return 0;
  }
  if (!(temperature > 0.0f)) return fail(-4, "temperature must be >= 0");
  sample_kernel<<<1, 1024, 0, s->stream>>>(s->logits_src, m->c.vocab_size, temperature, coin * top_p, -1, s->sample_out);
  CKL(cudaGetLastError());
  float h[2];
  CK(cudaMemcpyAsync(h, s->sample_out, sizeof(h), cudaMemcpyDeviceToHost, s->stream));
  CK(cudaStreamSynchronize(s->stream));
  memcpy(token, &h[0], sizeof(int));
  return 0;
}

extern "C" int dsk_sample_prob(dsk_model* m, dsk_state* s, int index, float* prob) {
  if (need_device()) return -1;
  if (!m || !s || !prob) return fail(-1, "bad arguments");
  if (index < 0 || index >= m->c.vocab_size) return fail(-4, "index %d out of range", index);
  if (need_logits(s, "dsk_sample_prob")) return -4;
  sample_kernel<<<1, 1024, 0, s->stream>>>(s->logits_src, m->c.vocab_size, 1.0f, 0.f, index, s->sample_out);
  CKL(cudaGetLastError());
  float h[2];
  CK(cudaMemcpyAsync(h, s->sample_out, sizeof(h), cudaMemcpyDeviceToHost, s->stream));
  CK(cudaStreamSynchronize(s->stream));
  *prob = h[1];
  return 0;
}

extern "C" int dsk_launches_per_forward(const dsk_model* m, int mode) {
  if (!m) return 0;
  (void)mode;
  const dsk_config& c = m->c;
  int cuts = 0;
  if (m->n_ranks > 1 && !m->p2p) for (int l = 0; l < c.n_layers; l++) cuts += m->layers[l].is_moe ? 1 : 0;   // peer-memory mode: one kernel
  return 1 + cuts * 3;
}

// ---------------------------------------------------------------------------------------------------
// multi-GPU
// ---------------------------------------------------------------------------------------------------
extern "C" int dsk_comm_unique_id(void* out128) {
  if (nccl_load()) return -3;
  ncclUniqueId id;
  CKN(g_nccl.GetUniqueId(&id));
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  memcpy(out128, &id, 128);
  return 0;
}
extern "C" int dsk_comm_init(dsk_model* m, const void* nccl_unique_id128) {
  if (need_device()) return -1;
  if (!m) return fail(-1, "null model");
  if (m->n_ranks == 1) return 0;
  if (nccl_load()) return -3;
  ncclUniqueId id;
  memcpy(&id, nccl_unique_id128, 128);
  CKN(g_nccl.CommInitRank(&m->comm, m->n_ranks, id, m->rank));
  // Peer-memory exchange (DSK_P2P=0 keeps the NCCL all-reduce between kernel segments): every rank allocates one exchange
  // buffer and maps every peer's through CUDA IPC; the handles travel through the NCCL communicator just created.
  const char* p2p_env = getenv("DSK_P2P");
  if (p2p_env && atoi(p2p_env) == 0) return 0;
  if (m->n_ranks > kMaxRanks || !g_nccl.AllGather) {
    if (m->tp) return fail(-3, "tensor-parallel model needs the peer-memory exchange (ncclAllGather missing): create it with DSK_TP=0");
    return 0;
  }
  const int N = m->n_ranks;
  // layout: [2 parities][N][dim] {partial, seq} 8-byte words | flags (4 KB) | tp: [2][N] arg-max keys (256 B) | tp: vocab logits
  const size_t data_bytes = (size_t)2 * N * m->c.dim * sizeof(unsigned long long);   // {value, sequence number} words
  const size_t flag_bytes = 4096;   // [N][dim / 256] chunk flags, then N arg-max flags (u32 each)
  if (((size_t)N * (size_t)cdiv(m->c.dim, 256) + (size_t)N) * 4 > flag_bytes) return fail(-4, "dim %d too large for the exchange flag area", m->c.dim);
  const size_t amax_off = data_bytes + flag_bytes, logits_off = amax_off + 256;
  const size_t total = logits_off + (m->tp ? (size_t)m->c.vocab_size * sizeof(float) : 0);
  cudaIpcMemHandle_t mine;
  unsigned char* d_handles = nullptr;
  std::vector<cudaIpcMemHandle_t> all(N);
  memset(&mine, 0, sizeof(mine));
  bool ok = cudaMalloc((void**)&m->xchg, total) == cudaSuccess && cudaMemset(m->xchg, 0, total) == cudaSuccess &&
            cudaIpcGetMemHandle(&mine, m->xchg) == cudaSuccess;
  // every rank takes part in the collectives below whatever happened locally (a failed rank is voted out afterwards)
  cudaStream_t st = nullptr;
  CK(cudaMalloc((void**)&d_handles, (size_t)N * sizeof(mine)));
  CK(cudaMemcpy(d_handles + (size_t)m->rank * sizeof(mine), &mine, sizeof(mine), cudaMemcpyHostToDevice));
  CKN(g_nccl.AllGather(d_handles + (size_t)m->rank * sizeof(mine), d_handles, sizeof(mine), ncclChar, m->comm, st));
  CK(cudaStreamSynchronize(st));
  CK(cudaMemcpy(all.data(), d_handles, (size_t)N * sizeof(mine), cudaMemcpyDeviceToHost));
  for (int q = 0; ok && q < N; q++) {
    void* ptr = m->xchg;
    if (q != m->rank && cudaIpcOpenMemHandle(&ptr, all[q], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = false; break; }
    m->xchg_peer[q] = (float*)ptr;
    m->xflag_peer[q] = (unsigned*)((unsigned char*)ptr + data_bytes);
    if (m->tp) {
      m->amax_peer[q] = (unsigned long long*)((unsigned char*)ptr + amax_off);
      m->logits_peer[q] = (float*)((unsigned char*)ptr + logits_off);
    }
  }
  if (d_handles) cudaFree(d_handles);
  cudaGetLastError();
  // agree on the mode: one rank without peer access sends everybody back to the NCCL path (all-reduce of a flag)
  float* flag = nullptr;
  CK(cudaMalloc((void**)&flag, sizeof(float)));
  const float mine_ok = ok ? 0.f : 1.f;
  CK(cudaMemcpy(flag, &mine_ok, sizeof(float), cudaMemcpyHostToDevice));
  CKN(g_nccl.AllReduce(flag, flag, 1, ncclFloat, ncclSum, m->comm, st));
  CK(cudaStreamSynchronize(st));
  float bad = 0.f;
  CK(cudaMemcpy(&bad, flag, sizeof(float), cudaMemcpyDeviceToHost));
  cudaFree(flag);
  m->p2p = bad == 0.f;
  if (m->p2p && m->tp) m->logits_full = m->logits_peer[m->rank];
  if (!m->p2p && m->tp)
    return fail(-3, "tensor-parallel weights were sliced at upload but the peer-memory mapping is unavailable on rank %d: create the model with DSK_TP=0 (expert-only sharding, NCCL exchange)", m->rank);
  if (!m->p2p) { fprintf(stderr, "[dsk] rank %d: peer-memory exchange unavailable, using NCCL all-reduce between kernel segments\n", m->rank); return 0; }
  m->p2p_epoch++;
  return 0;
}


// ---------------------------------------------------------------------------------------------------
// kernel-level test hooks.  Every hook builds a one-stage Program and runs it through decode_kernel<Q> — the persistent
// interpreter that produces every benchmark number — so the bit-exact / index-exact parity tests exercise the hot path
// itself (stage_q8 / q8_block_nf, kq_tile_rows, mma_rows_f8, route_all, c_attention, c_embed, gate_f32_stage).
// ---------------------------------------------------------------------------------------------------
struct Tmp {
  std::vector<void*> p;
  ~Tmp() { for (void* q : p) cudaFree(q); }
  template <typename T> T* up(const T* host, size_t n) {
    T* d = nullptr;
    if (cudaMalloc((void**)&d, std::max<size_t>(n * sizeof(T), 16)) != cudaSuccess) return nullptr;
    p.push_back(d);
    if (host && n) cudaMemcpy(d, host, n * sizeof(T), cudaMemcpyHostToDevice);
    else cudaMemset(d, 0, std::max<size_t>(n * sizeof(T), 16));
    return d;
  }
};

struct Mini {
  Tmp t;
  std::vector<Stage> S;
  dsk_config c{};            // only the fields the stages under test read
  int hd = 0;
  Program hdr{};             // extra header fields set by the hook (buffers, taps); geometry filled by run()
  Program* d_prog = nullptr;
  size_t smem = 0;
  Ctrl h_ctrl{};
  Mini() {
    memset(&hdr, 0, sizeof(hdr));
    c.dim = 256; c.n_heads = 1; c.vocab_size = 1; c.max_seq_len = 1; c.norm_eps = 1e-6f; c.act_silu = 1;
    c.qk_rope_head_dim = 2; c.v_head_dim = 1; c.kv_lora_rank = 1; c.bs0 = 128; c.bs1 = 128; c.original_max_position = 1 << 30;
    c.routed_scaling_factor = 1.0f; c.n_group = 1;
  }
  int prepare() {
    const int q = c.quant;
    g_f8_mma_ok = c.bs1 > 0 && (c.bs1 & (c.bs1 - 1)) == 0;
    int gate_dim = 0;
    for (const Stage& st : S) if (st.kind == ST_GEMV && st.quant == DSK_F32 && q != DSK_F32) gate_dim = std::max(gate_dim, st.n);
    choose_slot_geometry(q, S, gate_dim);
    int perr = 0;
    plan_stages(S, q, c.dim, g_sm_count, &perr);
    if (perr) return perr;
    Geometry geo;
    if (program_geometry(S, q, hd, c.max_seq_len, c.bs1, &geo)) return -4;
    smem = geo.smem;
    std::vector<unsigned char> buf(sizeof(Program) + (S.size() - 1) * sizeof(Stage));
    Program* P = reinterpret_cast<Program*>(buf.data());
    *P = hdr;
    fill_program_header(P, c, hd, geo);
    P->expert_first = 0; P->expert_count = c.n_routed_experts;
    P->embed_quant = q; P->n_stages = (int)S.size();
    P->ctrl = t.up<Ctrl>(&h_ctrl, 1);
    unsigned* sync = t.up<unsigned>(nullptr, 16);
    P->sync_counter = sync; P->sync_base = sync + 1;
    if (!P->ctrl || !sync) return fail(-2, "allocation failed");
    memcpy(P->stage, S.data(), S.size() * sizeof(Stage));
    d_prog = (Program*)t.up<unsigned char>(buf.data(), buf.size());
    if (!d_prog) return fail(-2, "allocation failed");
    return 0;
  }
  int launch(int b, int e) {
    CKL(launch_decode_raw(c.quant, d_prog, smem, b, e, 0, 1, 0));
    return 0;
  }
  int run() {
    if (prepare()) return -4;
    if (launch(0, (int)S.size())) return -2;
    CK(cudaDeviceSynchronize());
    return 0;
  }
};

static Stage gemv_stage(int quant, const float* in, const float* norm_w, int n, int epi) {
  Stage st{};
  st.kind = ST_GEMV; st.quant = quant; st.epi = epi; st.in = in; st.norm_w = norm_w; st.n = n; st.layer = 0;
  return st;
}
// device copy of a (d x n) weight payload in the layout the kernels stream (Q3_K re-pitched to 112-byte blocks, F8 rows padded)
static uint8_t* upload_test_weight(Tmp& t, int quant, int d, int n, const void* w) {
  const size_t drb = disk_row_bytes(quant, n), vrb = dev_row_bytes(quant, n);
  uint8_t* dw = t.up<uint8_t>(nullptr, vrb * d + 16);
  if (!dw) return nullptr;
  if (!w) return dw;
  if (quant == DSK_Q3_K) {
    uint8_t* st = t.up<uint8_t>((const uint8_t*)w, drb * d);
    const size_t nblocks = drb * d / kQ3Disk;
    q3k_repack_kernel<<<(unsigned)((nblocks + 7) / 8), 256>>>(st, dw, nblocks);
  } else if (vrb == drb) {
    cudaMemcpy(dw, w, drb * d, cudaMemcpyHostToDevice);
  } else {
    cudaMemcpy2D(dw, vrb, w, drb, drb, d, cudaMemcpyHostToDevice);
  }
  return dw;
}
static int check_gemv_shape(int quant, int n) {
  if (quant < 0 || quant > 4) return fail(-4, "bad quant");
  if ((quant >= DSK_Q2_K && n % 256) || (quant == DSK_F8E5M2 && n % 16) || (quant == DSK_F16 && n % 16) || (quant == DSK_F32 && n % 4) || n <= 0)
    return fail(-4, "n=%d not supported for quant %d (src/infer.cpp:169,246; src/quant.cpp:617)", n, quant);
  return 0;
}

extern "C" int dsk_gemv(int quant, int d, int n, const void* w, const float* scale, int bs0, int bs1, const float* x,
                        float* out) {
  if (need_device()) return -1;
  if (check_gemv_shape(quant, n)) return -4;
  if (d <= 0 || !w || !x || !out) return fail(-4, "bad arguments");
  if (quant == DSK_F8E5M2 && scale && (bs0 <= 0 || bs1 <= 0 || bs0 % 32 != 0)) return fail(-4, "f8e5m2 block size (%d, %d) unsupported (block_size_0 %% 32 != 0)", bs0, bs1);
  Mini mp;
  mp.c.quant = quant; mp.c.dim = d; mp.c.bs0 = bs0 > 0 ? bs0 : 128; mp.c.bs1 = bs1 > 0 ? bs1 : 128;
  uint8_t* dw = upload_test_weight(mp.t, quant, d, n, w);
  float* ds = scale ? mp.t.up<float>(scale, (size_t)cdiv(d, mp.c.bs0) * cdiv(n, mp.c.bs1)) : nullptr;
  float* dx = mp.t.up<float>(x, n);
  float* dout = mp.t.up<float>(nullptr, d);
  if (!dw || !dx || !dout) return fail(-2, "allocation failed");
  Stage st = gemv_stage(quant, dx, nullptr, n, EPI_STORE);
  MJob j{}; j.w = dw; j.scale = ds; j.out = dout; j.rows = d; j.expert_slot = -1;
  st.job[0] = j; st.njobs = 1;
  mp.S.push_back(st);
  if (mp.run()) return -2;
  CK(cudaMemcpy(out, dout, (size_t)d * 4, cudaMemcpyDeviceToHost));
  return 0;
}

extern "C" int dsk_stage_input(int model_quant, const float* x, const float* norm_w, int n, float eps, float* out_f32, void* out_q8) {
  if (need_device()) return -1;
  if (check_gemv_shape(model_quant, n)) return -4;
  if (!x) return fail(-4, "bad arguments");
  const bool kq = kq_quant(model_quant);
  if (kq && !out_q8) return fail(-4, "K-quant staging produces Q8_K blocks: out_q8 is required");
  if (!kq && !out_f32) return fail(-4, "out_f32 is required");
  Mini mp;
  mp.c.quant = model_quant; mp.c.dim = 256; mp.c.norm_eps = eps;
  uint8_t* dw = upload_test_weight(mp.t, model_quant, 16, n, nullptr);   // 16 zero rows: the tile loop runs, its result is unused
  std::vector<float> ones((size_t)cdiv(n, 128), 1.0f);
  float* ds = model_quant == DSK_F8E5M2 ? mp.t.up<float>(ones.data(), ones.size()) : nullptr;
  float* dx = mp.t.up<float>(x, n);
  float* dn = norm_w ? mp.t.up<float>(norm_w, n) : nullptr;
  float* dout = mp.t.up<float>(nullptr, 16);
  unsigned char* dq8 = kq ? mp.t.up<unsigned char>(nullptr, (size_t)(n / 256) * 292) : nullptr;
  float* dxf = kq ? nullptr : mp.t.up<float>(nullptr, n);
  if (!dw || !dx || !dout) return fail(-2, "allocation failed");
  mp.hdr.dbg_q8 = dq8; mp.hdr.dbg_x = dxf;
  Stage st = gemv_stage(model_quant, dx, dn, n, EPI_STORE);
  MJob j{}; j.w = dw; j.scale = ds; j.out = dout; j.rows = 16; j.expert_slot = -1;
  st.job[0] = j; st.njobs = 1;
  mp.S.push_back(st);
  if (mp.run()) return -2;
  if (kq) CK(cudaMemcpy(out_q8, dq8, (size_t)(n / 256) * 292, cudaMemcpyDeviceToHost));
  else CK(cudaMemcpy(out_f32, dxf, (size_t)n * 4, cudaMemcpyDeviceToHost));
  return 0;
}

extern "C" int dsk_quantize_q8k(const float* x, int k, void* out) {
  if (k <= 0 || k % 256) return fail(-4, "k must be a positive multiple of 256 (src/quant.cpp:617)");
  return dsk_stage_input(DSK_Q2_K, x, nullptr, k, 0.f, nullptr, out);
}

extern "C" int dsk_rmsnorm(const float* x, const float* w, int n, float eps, float* out) {
  if (n <= 0 || n % 4) return fail(-4, "n must be a positive multiple of 4");
  return dsk_stage_input(DSK_F32, x, w, n, eps, out, nullptr);
}

extern "C" int dsk_gate_logits(int model_quant, int n_experts, int n, const float* gate_w, const float* x, const float* norm_w,
                               float eps, float* out_logits, float* out_xnorm) {
  if (need_device()) return -1;
  if (model_quant < 0 || model_quant > 4 || n_experts <= 0 || n <= 0 || n % 4 || !gate_w || !x || !out_logits) return fail(-4, "bad arguments");
  Mini mp;
  mp.c.quant = model_quant; mp.c.dim = n; mp.c.norm_eps = eps; mp.c.n_routed_experts = n_experts;
  float* dw = mp.t.up<float>(gate_w, (size_t)n_experts * n);
  float* dx = mp.t.up<float>(x, n);
  float* dn = norm_w ? mp.t.up<float>(norm_w, n) : nullptr;
  float* dout = mp.t.up<float>(nullptr, n_experts);
  float* dxf = out_xnorm ? mp.t.up<float>(nullptr, n) : nullptr;
  if (!dw || !dx || !dout) return fail(-2, "allocation failed");
  mp.hdr.dbg_x = dxf;
  Stage st = gemv_stage(DSK_F32, dx, dn, n, EPI_STORE);
  MJob j{}; j.w = (const uint8_t*)dw; j.out = dout; j.rows = n_experts; j.expert_slot = -1;
  st.job[0] = j; st.njobs = 1;
  if (model_quant != DSK_F32) {   // the dedicated gate stage of a quantised model: one row per tile (build_program does the same)
    st.rows_per_tile = 1; st.rpass = 1; st.npieces = 1; st.wp = 0; st.use_mma = 0;
    st.piece[0] = Piece{0, 0, n / 4, 0};
    st.job[0].tile_begin = 0; st.ntiles = n_experts; st.has_dyn = 0;
  }
  mp.S.push_back(st);
  if (mp.run()) return -2;
  CK(cudaMemcpy(out_logits, dout, (size_t)n_experts * 4, cudaMemcpyDeviceToHost));
  if (out_xnorm) CK(cudaMemcpy(out_xnorm, dxf, (size_t)n * 4, cudaMemcpyDeviceToHost));
  return 0;
}

extern "C" int dsk_dequantize_row(int quant, const void* blocks, int k, float* out) {
  if (need_device()) return -1;
  if ((quant != DSK_Q2_K && quant != DSK_Q3_K) || k <= 0 || k % 256 || !blocks || !out) return fail(-4, "bad arguments");
  Mini mp;
  mp.c.quant = quant; mp.c.dim = k;
  uint8_t* dw = upload_test_weight(mp.t, quant, 1, k, blocks);
  float* dx = mp.t.up<float>(nullptr, k);
  if (!dw || !dx) return fail(-2, "allocation failed");
  mp.hdr.embed_w = dw; mp.hdr.x = dx;
  Stage st{}; st.kind = ST_EMBED; st.quant = quant;
  mp.S.push_back(st);
  mp.h_ctrl.token = 0;
  if (mp.run()) return -2;
  CK(cudaMemcpy(out, dx, (size_t)k * 4, cudaMemcpyDeviceToHost));
  return 0;
}

// rope / rope_v3 on one head's rotary slice (the reference always calls them with d == head_dim == qk_rope_head_dim,
// src/infer.cpp:956-972): run as the RoPE prologue of the attention stage on a head with no `nope` part
extern "C" int dsk_rope(float* vec, int d, int head_dim, int pos, float theta, int v3) {
  if (need_device()) return -1;
  if (!vec || d <= 0 || d % 2 || d > 128 || d != head_dim) return fail(-4, "rope hook: d must equal head_dim, be even and <= 128 (got d=%d, head_dim=%d)", d, head_dim);
  if (pos < 0) return fail(-4, "negative pos");
  Mini mp;
  mp.c.quant = DSK_F32; mp.c.n_heads = 1; mp.c.qk_nope_head_dim = 0; mp.c.qk_rope_head_dim = d; mp.c.v_head_dim = 4;
  mp.c.kv_lora_rank = 4; mp.c.is_v3 = v3 ? 1 : 0; mp.c.max_seq_len = 1;
  mp.hd = d;
  std::vector<float> fr = rope_table(d, theta);
  float* df = mp.t.up<float>(fr.data(), fr.size());
  float* dq = mp.t.up<float>(vec, d);
  float* dkva = mp.t.up<float>(nullptr, 4 + d);
  float* dout = mp.t.up<float>(nullptr, 4);
  __half* kc = mp.t.up<__half>(nullptr, d);
  __half* vc = mp.t.up<__half>(nullptr, 4);
  float* scratch = mp.t.up<float>(nullptr, 1 + kConsumers + 8);
  if (!df || !dq || !dkva || !dout || !kc || !vc || !scratch) return fail(-2, "allocation failed");
  mp.hdr.rope_freq = df; mp.hdr.q = dq; mp.hdr.kv_a = dkva; mp.hdr.xb2 = dout; mp.hdr.att_scratch = scratch;
  Stage st{}; st.kind = ST_ATTN; st.quant = DSK_F32; st.kcache = kc; st.vcache = vc;
  mp.S.push_back(st);
  mp.h_ctrl.pos = pos; mp.h_ctrl.kv_pos = 0; mp.h_ctrl.kv_len = 1; mp.h_ctrl.kv_sink = 0;
  if (mp.run()) return -2;
  CK(cudaMemcpy(vec, dq, (size_t)d * 4, cudaMemcpyDeviceToHost));
  return 0;
}

extern "C" int dsk_moe_gate(float* logits, const float* bias, int n_routed, int n_active, int norm_topk_prob,
                            float routed_scaling_factor, int scoring_sigmoid, int topk_method, int n_group, int topk_group,
                            int32_t* active_experts, float* weights) {
  if (need_device()) return -1;
  if (!logits || !active_experts || !weights || n_routed <= 0 || n_routed > 256 || n_active <= 0 || n_active > 16 || n_active + 1 > kMaxJobs)
    return fail(-4, "E <= 256 (src/infer.cpp:527), K <= %d", kMaxJobs - 1);
  Mini mp;
  dsk_config& c = mp.c;
  c.quant = DSK_Q2_K; c.dim = 256; c.n_routed_experts = n_routed; c.n_active_routed = n_active; c.norm_topk_prob = norm_topk_prob;
  c.routed_scaling_factor = routed_scaling_factor; c.scoring_sigmoid = scoring_sigmoid; c.topk_method = topk_method;
  c.n_group = std::max(1, n_group); c.topk_group = topk_group;
  if (topk_method == DSK_TOPK_GROUP_LIMITED_GREEDY && (n_routed % c.n_group != 0 || topk_group <= 0)) return fail(-4, "bad group configuration");
  float* dl = mp.t.up<float>(logits, n_routed);
  float* db = bias ? mp.t.up<float>(bias, n_routed) : nullptr;
  float* dscores = mp.t.up<float>(nullptr, 256);
  int* dact = mp.t.up<int>(nullptr, 16);
  float* dactw = mp.t.up<float>(nullptr, 16);
  // the routing runs in the prologue of the S56 stage: give that stage one (all-zero) shared-expert tile to stream
  uint8_t* dw = upload_test_weight(mp.t, DSK_Q2_K, 16, 256, nullptr);
  float* dx = mp.t.up<float>(nullptr, 256);
  float* dh = mp.t.up<float>(nullptr, 16);
  if (!dl || !dscores || !dact || !dactw || !dw || !dx || !dh) return fail(-2, "allocation failed");
  mp.hdr.moe_scores = dscores; mp.hdr.act = dact; mp.hdr.act_w = dactw;
  Stage st = gemv_stage(DSK_Q2_K, dx, nullptr, 256, EPI_GLU);
  st.need_topk = 1; st.gate_logits = dl; st.gate_bias = db;
  MJob j{}; j.w = dw; j.w_b = dw; j.out = dh; j.rows = 16; j.expert_slot = -1;
  st.job[0] = j; st.njobs = 1;
  mp.S.push_back(st);
  if (mp.run()) return -2;
  CK(cudaMemcpy(logits, dscores, (size_t)n_routed * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(active_experts, dact, (size_t)n_active * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(weights, dactw, (size_t)n_active * 4, cudaMemcpyDeviceToHost));
  return 0;
}

extern "C" int dsk_attn(const float* q, const uint16_t* kcache, const uint16_t* vcache, int n_heads, int head_dim,
                        int v_head_dim, int kv_len, float* out) {
  if (need_device()) return -1;
  if (!q || !kcache || !vcache || !out || n_heads <= 0 || head_dim <= 0 || v_head_dim <= 0 || kv_len <= 0) return fail(-4, "bad arguments");
  Mini mp;
  dsk_config& c = mp.c;
  c.quant = DSK_Q2_K; c.n_heads = n_heads; c.qk_nope_head_dim = head_dim; c.qk_rope_head_dim = 0; c.v_head_dim = v_head_dim;
  c.kv_lora_rank = 4; c.max_seq_len = kv_len;
  mp.hd = head_dim;
  float* dq = mp.t.up<float>(q, (size_t)n_heads * head_dim);
  uint16_t* dk = mp.t.up<uint16_t>(kcache, (size_t)kv_len * n_heads * head_dim);
  uint16_t* dv = mp.t.up<uint16_t>(vcache, (size_t)kv_len * n_heads * v_head_dim);
  float* dout = mp.t.up<float>(nullptr, (size_t)n_heads * v_head_dim);
  float* dkva = mp.t.up<float>(nullptr, 8);
  float* df = mp.t.up<float>(nullptr, 4);
  float* scratch = mp.t.up<float>(nullptr, (size_t)n_heads * (kv_len + kConsumers + 8));
  if (!dq || !dk || !dv || !dout || !dkva || !df || !scratch) return fail(-2, "allocation failed");
  mp.hdr.q = dq; mp.hdr.kv_a = dkva; mp.hdr.xb2 = dout; mp.hdr.rope_freq = df; mp.hdr.att_scratch = scratch;
  Stage st{}; st.kind = ST_ATTN; st.quant = DSK_Q2_K; st.kcache = (__half*)dk; st.vcache = (__half*)dv;
  mp.S.push_back(st);
  mp.h_ctrl.pos = kv_len - 1; mp.h_ctrl.kv_pos = kv_len - 1; mp.h_ctrl.kv_len = kv_len; mp.h_ctrl.kv_sink = 0;
  if (mp.run()) return -2;
  CK(cudaMemcpy(out, dout, (size_t)n_heads * v_head_dim * 4, cudaMemcpyDeviceToHost));
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// measurement hook: ONE interpreter GEMV stage (the production tile plan, TMA ring, warp-per-tile reduction) timed alone
// ---------------------------------------------------------------------------------------------------
__global__ void fill_pattern_kernel(uint32_t* p, size_t n_words, uint32_t seed, uint32_t and_mask, uint32_t or_mask) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n_words; i += stride) {
    uint32_t x = (uint32_t)i * 2654435761u + seed;
    x ^= x >> 15; x *= 2246822519u; x ^= x >> 13;
    p[i] = (x & and_mask) | or_mask;
  }
}

extern "C" int dsk_bench_gemv(int quant, int d, int n, int n_mats, int warmup, int iters, float* avg_ms,
                              double* bytes_per_launch) {
  if (need_device()) return -1;
  if (check_gemv_shape(quant, n) || n_mats < 1 || n_mats > 64 || d <= 0) return fail(-4, "bad arguments");
  Mini mp;
  mp.c.quant = quant; mp.c.dim = 256; mp.c.bs0 = 128; mp.c.bs1 = 128;
  const size_t vrb = dev_row_bytes(quant, n);
  const size_t mat = (vrb * d + 255) & ~(size_t)255;
  uint8_t* w = mp.t.up<uint8_t>(nullptr, mat * n_mats + 256);
  if (!w) return fail(-2, "allocation of %zu bytes failed", mat * n_mats);
  // finite synthetic payload: f8/f16 exponent bits kept small; K-quant bytes arbitrary (d/dmin patched by mask)
  uint32_t and_mask = 0xFFFFFFFFu, or_mask = 0;
  if (quant == DSK_F8E5M2) and_mask = 0xBFBFBFBFu;       // clear the top exponent bit of every byte
  if (quant == DSK_F16) and_mask = 0xBFFFBFFFu;
  if (quant == DSK_F32) { and_mask = 0xBFFFFFFFu; }
  if (quant >= DSK_Q2_K) and_mask = 0x3F3F3F3Fu;          // keeps the fp16 d/dmin fields finite wherever they fall
  fill_pattern_kernel<<<1024, 256>>>((uint32_t*)w, mat * n_mats / 4, 12345u, and_mask, or_mask);
  const int srows = cdiv(d, 128), scols = cdiv(n, 128);
  std::vector<float> hs((size_t)srows * scols, 1e-3f);
  float* ds = quant == DSK_F8E5M2 ? mp.t.up<float>(hs.data(), hs.size()) : nullptr;
  std::vector<float> hx(n, 0.5f);
  float* dx = mp.t.up<float>(hx.data(), n);
  float* dout = mp.t.up<float>(nullptr, d);
  if (!dx || !dout) return fail(-2, "allocation failed");
  for (int i = 0; i < n_mats; i++) {   // one stage per matrix: consecutive launches never re-read L2-resident weights
    Stage st = gemv_stage(quant, dx, nullptr, n, EPI_STORE);
    MJob j{}; j.w = w + (size_t)i * mat; j.scale = ds; j.out = dout; j.rows = d; j.expert_slot = -1;
    st.job[0] = j; st.njobs = 1;
    mp.S.push_back(st);
  }
  if (mp.prepare()) return -4;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int i = 0; i < warmup + iters; i++) {
    if (i == warmup) CK(cudaEventRecord(e0, 0));
    if (mp.launch(i % n_mats, i % n_mats + 1)) return -2;
  }
  CK(cudaEventRecord(e1, 0));
  CK(cudaEventSynchronize(e1));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (avg_ms) *avg_ms = ms / iters;
  double bpw = quant == DSK_F32 ? 4 : quant == DSK_F16 ? 2 : quant == DSK_F8E5M2 ? 1.0 + 4.0 / 16384 : quant == DSK_Q2_K ? 84.0 / 256 : 110.0 / 256;
  if (bytes_per_launch) *bytes_per_launch = (double)d * n * bpw;
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// stage-level timeline of one token from the interpreter's own globaltimer stamps (CTA 0)
// ---------------------------------------------------------------------------------------------------
extern "C" int dsk_profile_token(dsk_model* m, dsk_state* s, int token, int pos, char* out, size_t cap) {
  if (need_device()) return -1;
  if (!m || !s || !out) return fail(-1, "bad arguments");
  if (ensure_program(m, s)) return -2;
  {  // switch the stamps on for this token only (they cost a few globaltimer reads per stage)
    unsigned long long* on = s->tstamp;
    CK(cudaMemcpy(reinterpret_cast<unsigned char*>(s->prog) + offsetof(Program, tstamp), &on, sizeof(on), cudaMemcpyHostToDevice));
  }
  const int rc = dsk_forward(m, s, token, pos, 1, nullptr, nullptr);
  if (!getenv("DSK_TSTAMP")) {
    unsigned long long* off = nullptr;
    CK(cudaMemcpy(reinterpret_cast<unsigned char*>(s->prog) + offsetof(Program, tstamp), &off, sizeof(off), cudaMemcpyHostToDevice));
  }
  if (rc) return -2;
  std::vector<unsigned long long> ts((size_t)s->n_stages * 8 + 8);
  CK(cudaMemcpy(ts.data(), s->tstamp, ts.size() * 8, cudaMemcpyDeviceToHost));
  struct Agg { int n = 0; double wait = 0, stage = 0, tiles = 0, arrive = 0, cw = 0, ct = 0, cs = 0, ce = 0; };
  std::map<std::string, Agg> agg;
  double total = 0;
  for (int i = 0; i < s->n_stages; i++) {
    const double t0 = (double)ts[i * 8], t1 = (double)ts[i * 8 + 1], t2 = (double)ts[i * 8 + 2], t3 = (double)ts[i * 8 + 3];
    const double prev_end = i > 0 ? (double)ts[(i - 1) * 8 + 3] : t0;
    Agg& a = agg[s->stage_names[i]];
    a.n++;
    a.wait += (t0 - prev_end) / 1e3;                      // grid barrier wait
    a.stage += (t1 > 0 ? (t1 - t0) : 0) / 1e3;           // routing + activation staging
    a.tiles += (t2 - (t1 > 0 ? t1 : t0)) / 1e3;          // tile loop (or attention / embed body)
    a.arrive += (t3 - t2) / 1e3;                         // fence + arrive
    a.cw += (double)ts[i * 8 + 4]; a.ct += (double)ts[i * 8 + 5]; a.cs += (double)ts[i * 8 + 6]; a.ce += (double)ts[i * 8 + 7];
    if (i + 1 == s->n_stages) total = ((double)ts[i * 8 + 3] - (double)ts[0]) / 1e3;
  }
  size_t off = 0;
  for (auto& kv : agg) {
    const Agg& a = kv.second;
    int n = snprintf(out + off, cap - off, "%s x%3d  barrier %6.2f  stage-in %6.2f  tiles %6.2f  arrive %5.2f us | kcyc/stage: wait %6.1f task %6.1f sync %6.1f epi %6.1f | sum %8.1f us\n",
                     kv.first.c_str(), a.n, a.wait / a.n, a.stage / a.n, a.tiles / a.n, a.arrive / a.n, a.cw / a.n / 1e3, a.ct / a.n / 1e3, a.cs / a.n / 1e3, a.ce / a.n / 1e3, a.wait + a.stage + a.tiles + a.arrive);
    if (n < 0 || (size_t)n >= cap - off) break;
    off += n;
  }
  snprintf(out + off, cap - off, "token total %.1f us over %d stages (CTA 0 timeline)\n", total, s->n_stages);
  return 0;
}
