// dseek_loader.h — `.dseek` checkpoint directory reader for the C++ host (mmap + a 150-line JSON header parser).
// Byte-compatible with the reference's container (src/codec.cpp:262-377): every directory entry is a shard
// `u64 LE header_len | JSON header | raw tensor bytes`; `__metadata__` (flat string->string) comes from the first shard
// in sorted order; tensors are merged from all shards.  Payloads are handed to dsk_upload_tensor() untouched.
#pragma once
#include <cstddef>
#include <cstdint>
#include <map>
#include <string>
#include <vector>

struct DseekTensor {
  std::string name, dtype;
  int64_t shape[4] = {0, 0, 0, 0};
  const void* data = nullptr;
  size_t size = 0;
};

struct DseekData {
  std::map<std::string, std::string> metadata;
  std::map<std::string, DseekTensor> tensors;
  std::vector<std::pair<void*, size_t>> maps;
  // returns "" on success, else an error message
  std::string load(const std::string& dirname, bool lock_weights);
  ~DseekData();
};
