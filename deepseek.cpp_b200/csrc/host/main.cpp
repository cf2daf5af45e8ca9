// main.cpp — `main <checkpoint_dir> [options]`: the reference's CLI surface (src/main.cpp:18-43, 594-691) over libdsk.so.
// Host side in C++ calling CUDA through the thin C-ABI (include/dsk.h): .dseek loader, trie tokenizer
// (src/tokenizer.cpp:3-94), completion + perplexity modes.  The forward pass is dsk_forward(); sampling (greedy, temperature /
// top-p, perplexity probabilities) runs on the device through dsk_sample / dsk_sample_prob, so no logits cross PCIe.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../../include/dsk.h"
#include "dseek_loader.h"

static void error_usage() {
  fprintf(stderr, "Usage:   main <checkpoint_dir> [options]\n");
  fprintf(stderr, "Example: main model_weights_dir/ -i \"Q: What is the meaning of life?\"\n");
  fprintf(stderr, "Options:\n");
  fprintf(stderr, "  -h Display this help message\n");
  fprintf(stderr, "  -L Locks model weights to RAM while they are uploaded\n");
  fprintf(stderr, "  -m [completion,perplexity] which mode to run in (default - completion)\n");
  fprintf(stderr, "  -T <int> sliding window context length (0 - max)\n");
  fprintf(stderr, "  -g <int> CUDA device (default 0)\n");
  fprintf(stderr, "Completion mode options:\n");
  fprintf(stderr, "  -n <int>    number of steps to run for in completion mode, default 256. 0 = max_seq_len, -1 = infinite\n");
  fprintf(stderr, "  -i <string> input prompt | -f <filepath> input file with prompt\n");
  fprintf(stderr, "  -t <float> temperature (default - 1.0; 0 = greedy)   -p <float> p for top-p sampling (default - 0.95)\n");
  fprintf(stderr, "Perplexity mode options:  -i <string> | -f <filepath>\n");
  exit(1);
}

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// ---- tokenizer: greedy longest-match over a trie, byte fallback (src/tokenizer.cpp:3-94) ----------------------------
struct TokenTrie { std::unordered_map<char, std::unique_ptr<TokenTrie>> children; int token_id = -1; };
struct Tokenizer {
  std::vector<std::string> vocab;
  TokenTrie root;
  int bos_id = -1, eos_id = -1, eot_id = -1, byte_fallback_start = -1;
  Tokenizer(const DseekData& d) {
    bos_id = std::stoi(d.metadata.at("bos_token_id"));
    eos_id = std::stoi(d.metadata.at("eos_token_id"));
    const DseekTensor& t = d.tensors.at("tokenizer.tokens");
    const char* p = (const char*)t.data; const char* end = p + t.size;
    for (; p < end; p++) { const char* s = p; while (p < end && *p != '\0') p++; vocab.emplace_back(s, p - s); }
    for (size_t i = 0; i < vocab.size(); i++) {
      if (vocab[i] == "<0x00>") byte_fallback_start = (int)i;
      else if (vocab[i] == "<|eot_id|>" || vocab[i] == "<|end|>" || vocab[i] == "<|im_end|>") eot_id = (int)i;
      TokenTrie* n = &root;
      for (char c : vocab[i]) { auto& ch = n->children[c]; if (!ch) ch = std::make_unique<TokenTrie>(); n = ch.get(); }
      n->token_id = (int)i;
    }
  }
  std::vector<int> encode(const std::string& text, bool bos) const {
    std::vector<int> out;
    if (bos) out.push_back(bos_id);
    for (size_t i = 0; i < text.size();) {
      size_t l = 0, valid_l = 0; const TokenTrie* n = &root; const TokenTrie* valid = nullptr;
      while (i + l < text.size()) {
        auto it = n->children.find(text[i + l]);
        if (it == n->children.end()) break;
        n = it->second.get(); l++;
        if (n->token_id >= 0) { valid = n; valid_l = l; }
      }
      if (!valid) { if (byte_fallback_start >= 0) out.push_back((unsigned char)text[i] + byte_fallback_start); i++; }
      else { out.push_back(valid->token_id); i += valid_l; }
    }
    return out;
  }
  std::string decode_one(int prev, int tok) const {
    const std::string& piece = vocab[tok];
    if (prev == bos_id && !piece.empty() && piece[0] == ' ') return piece.substr(1);
    if (byte_fallback_start >= 0 && tok >= byte_fallback_start && tok - byte_fallback_start < 256) return std::string(1, (char)(tok - byte_fallback_start));
    return piece;
  }
};

// Sampling runs on the device (dsk_sample / dsk_sample_prob: Sampler::sample / sample_prob, src/sampler.cpp:12-75); the host
// only draws the random number, with the reference's generator and seed policy (std::srand / std::rand, src/sampler.cpp:7-10,65).
static int meta_i(const DseekData& d, const char* k, int def, bool required = false) {
  auto it = d.metadata.find(k);
  if (it == d.metadata.end()) { if (required) { fprintf(stderr, "FATAL: missing metadata %s\n", k); exit(1); } return def; }
  return std::stoi(it->second);
}
static float meta_f(const DseekData& d, const char* k, float def) { auto it = d.metadata.find(k); return it == d.metadata.end() ? def : std::stof(it->second); }
static std::string meta_s(const DseekData& d, const char* k, const char* def) { auto it = d.metadata.find(k); return it == d.metadata.end() ? def : it->second; }

// Config::from_yalm (src/model.cpp:22-127)
static dsk_config config_from(const DseekData& d, int context) {
  dsk_config c{};
  c.dim = meta_i(d, "dim", 0, true); c.hidden_dim = meta_i(d, "hidden_dim", 0, true); c.n_layers = meta_i(d, "n_layers", 0, true);
  c.n_heads = meta_i(d, "n_heads", 0, true); c.vocab_size = meta_i(d, "vocab_size", 0, true); c.max_seq_len = meta_i(d, "max_seq_len", 0, true);
  if (context) c.max_seq_len = std::min(c.max_seq_len, context);
  c.rope_theta = meta_f(d, "rope_theta", 10000.f); c.norm_eps = meta_f(d, "norm_eps", 1e-5f);
  c.act_silu = meta_s(d, "act_type", "gelu") == "silu";
  c.first_k_dense_replace = meta_i(d, "first_k_dense_replace", 0);
  c.n_shared_experts = meta_i(d, "n_shared_experts", 0); c.n_routed_experts = meta_i(d, "n_routed_experts", 0);
  c.n_active_routed = meta_i(d, "n_active_routed", 0); c.moe_intermediate_size = meta_i(d, "moe_intermediate_size", 0);
  c.routed_scaling_factor = meta_f(d, "routed_scaling_factor", 1.0f); c.n_group = meta_i(d, "n_group", 1);
  c.norm_topk_prob = meta_s(d, "norm_topk_prob", "False") == "True";
  c.scoring_sigmoid = meta_s(d, "scoring_func", "softmax") == "sigmoid";
  c.topk_group = meta_i(d, "topk_group", 0);
  std::string tm = meta_s(d, "topk_method", "");
  if (tm == "noaux_tc") { fprintf(stderr, "FATAL: topk_method noaux_tc unsupported (src/model.cpp:51-53)\n"); exit(1); }
  c.topk_method = tm == "group_limited_greedy";
  c.is_v3 = meta_s(d, "arch", "") == "DeepseekV3ForCausalLM";
  c.kv_lora_rank = meta_i(d, "kv_lora_rank", 0); c.q_lora_rank = meta_i(d, "q_lora_rank", 0);
  c.qk_nope_head_dim = meta_i(d, "qk_nope_head_dim", 0); c.qk_rope_head_dim = meta_i(d, "qk_rope_head_dim", 0); c.v_head_dim = meta_i(d, "v_head_dim", 0);
  std::string q = meta_s(d, "quant", "");
  if (q == "fp32") c.quant = DSK_F32; else if (q == "fp16") c.quant = DSK_F16; else if (q == "f8e5m2") c.quant = DSK_F8E5M2;
  else if (q == "q2_k") c.quant = DSK_Q2_K; else if (q == "q3_k") c.quant = DSK_Q3_K;
  else { fprintf(stderr, "FATAL: unsupported quant: %s\n", q.c_str()); exit(1); }
  c.bs0 = meta_i(d, "quantization_block_size_0", 0); c.bs1 = meta_i(d, "quantization_block_size_1", 0);
  c.original_max_position = meta_i(d, "rope_scaling_original_max_position_embeddings", 4096);
  c.use_mla = meta_i(d, "use_mla", 0) ? 1 : 0;
  return c;
}

#define DSK_OK(call) do { if ((call) != 0) { fprintf(stderr, "FATAL: %s: %s\n", #call, dsk_last_error()); exit(1); } } while (0)

int main(int argc, char* argv[]) {
  if (argc < 2) error_usage();
  std::string dir = argv[1], mode = "completion", prompt, prompt_path;
  bool lock = false; int context = 0, device = 0, num_steps = 256; float temperature = 1.0f, top_p = 0.95f;
  for (int i = 2; i < argc;) {
    if (argv[i][0] != '-' || strlen(argv[i]) != 2) error_usage();
    char f = argv[i][1];
    if (f == 'h') error_usage();
    if (f == 'L') { lock = true; i++; continue; }
    if (i + 1 >= argc) error_usage();
    const char* v = argv[i + 1];
    switch (f) {
      case 'm': if (std::string("completion").rfind(v, 0) == 0) mode = "completion"; else if (std::string("perplexity").rfind(v, 0) == 0) mode = "perplexity"; else error_usage(); break;
      case 'T': context = atoi(v); break; case 'g': device = atoi(v); break; case 'n': num_steps = atoi(v); break;
      case 'i': prompt = v; break; case 'f': prompt_path = v; break; case 't': temperature = (float)atof(v); break; case 'p': top_p = (float)atof(v); break;
      default: error_usage();
    }
    i += 2;
  }
  if (!prompt_path.empty()) { std::ifstream fs(prompt_path); std::stringstream ss; ss << fs.rdbuf(); prompt = ss.str(); }
  if (prompt.empty()) { fprintf(stderr, "No prompt provided\n"); error_usage(); }

  DseekData data;
  std::string err = data.load(dir, lock);
  if (!err.empty()) { fprintf(stderr, "failed to load checkpoint: %s\n", err.c_str()); return 1; }
  dsk_config cfg = config_from(data, context);
  std::cout << "loading model with quant: " << data.metadata["quant"] << std::endl;
  DSK_OK(dsk_init(device));
  double t0 = now_s();
  dsk_model* model = dsk_model_create(&cfg, 0, 1);
  if (!model) { fprintf(stderr, "FATAL: %s\n", dsk_last_error()); return 1; }
  size_t up = 0;
  for (auto& kv : data.tensors) {
    const DseekTensor& t = kv.second;
    static const char* names[] = {"F32", "F16", "BF16", "F8_E5M2", "F8_E4M3", "I32", "I16", "I8", "U8"};   // CodecDType order
    int dt = -1;
    for (int k = 0; k < 9; k++) if (t.dtype == names[k]) dt = k;
    if (dt < 0) { fprintf(stderr, "FATAL: tensor %s has unknown dtype %s\n", t.name.c_str(), t.dtype.c_str()); return 1; }
    DSK_OK(dsk_upload_tensor(model, t.name.c_str(), dt, t.shape, t.data, t.size, 0));
    up += t.size;
  }
  DSK_OK(dsk_model_finalize(model));
  dsk_state* state = dsk_state_create(model);
  if (!state) { fprintf(stderr, "FATAL: %s\n", dsk_last_error()); return 1; }
  std::cout << "uploaded " << up / 1e9 << " GB in " << now_s() - t0 << " s (" << dsk_model_resident_bytes(model) / 1e9 << " GB resident)" << std::endl;
  std::cout << "Model active bytes per token (algorithmic): " << dsk_model_active_bytes_per_token(model) << std::endl;

  Tokenizer tokenizer(data);
  srand((unsigned)(uint64_t)(now_s() * 1000));
  if (num_steps == 0) num_steps = cfg.max_seq_len;
  DSK_OK(dsk_forward(model, state, 0, 0, DSK_OUTPUT_LOGITS, nullptr, nullptr));   // warm-up (graph capture), like src/main.cpp:299-303
  std::vector<int> enc = tokenizer.encode(prompt, true);
  std::cout << "[";
  for (size_t i = 0; i < enc.size(); i++) std::cout << (i ? "," : "") << enc[i];
  std::cout << "]" << std::endl;

  if (mode == "perplexity") {  // src/main.cpp:371-431
    double sum_nll = 0, ss = 0; size_t n = 0;
    for (size_t pos = 0; pos + 1 < enc.size(); pos++) {
      DSK_OK(dsk_forward(model, state, enc[pos], (int)pos, DSK_OUTPUT_LOGITS, nullptr, nullptr));
      float pr = 0.f;
      DSK_OK(dsk_sample_prob(model, state, enc[pos + 1], &pr));
      double lp = std::log((double)pr);
      sum_nll += -lp; ss += lp * lp; n++;
    }
    double mean = sum_nll / n, var = ss / n - mean * mean;
    std::cout << "perplexity: " << std::exp(mean) << " ± " << std::exp(mean) * std::sqrt(std::max(0.0, var) / n) << " over " << n << " tokens" << std::endl;
    return 0;
  }
  double start = now_s();
  for (size_t pos = 0; pos < enc.size(); pos++) {   // hydrate (src/main.cpp:312-319)
    bool last = pos + 1 == enc.size();
    DSK_OK(dsk_forward(model, state, enc[pos], (int)pos, last ? DSK_OUTPUT_LOGITS : DSK_HYDRATE_KV_CACHE, nullptr, nullptr));
  }
  double end_hydrate = now_s();
  for (int i = 0; i < num_steps || num_steps == -1; i++) {   // src/main.cpp:324-335
    int tok = -1;
    const float coin = temperature == 0.0f ? 0.f : rand() / (float)RAND_MAX;
    DSK_OK(dsk_sample(model, state, temperature, top_p, coin, &tok));
    std::cout << tokenizer.decode_one(enc.back(), tok) << std::flush;
    enc.push_back(tok);
    if (tok == tokenizer.eos_id || tok == tokenizer.eot_id) break;
    if ((int)enc.size() - 1 >= cfg.max_seq_len && cfg.original_max_position > cfg.max_seq_len) break;
    DSK_OK(dsk_forward(model, state, tok, (int)enc.size() - 1, DSK_OUTPUT_LOGITS, nullptr, nullptr));
  }
  double elapsed = now_s() - start;
  std::cout << "\n\nGeneration stats:\n  " << enc.size() << " tokens\n  throughput: " << enc.size() / elapsed << "tok/s\n  latency: " << elapsed / enc.size()
            << "s/tok\n  hydrate: " << end_hydrate - start << "s\n  bandwidth: " << dsk_model_active_bytes_per_token(model) * enc.size() / 1e9 / elapsed
            << "GB/s\n  total: " << elapsed << "s\n" << std::endl;
  std::cout << "generated ids:";
  for (size_t i = 0; i < enc.size(); i++) std::cout << " " << enc[i];
  std::cout << std::endl;
  dsk_state_destroy(state);
  dsk_model_destroy(model);
  return 0;
}
