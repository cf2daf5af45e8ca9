#include "dseek_loader.h"

#include <algorithm>
#include <cstring>
#include <dirent.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace {
// Minimal JSON reader for safetensors headers: objects, arrays, strings, integers.
struct JParser {
  const char* p; const char* end; std::string err;
  void ws() { while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) p++; }
  bool lit(char c) { ws(); if (p < end && *p == c) { p++; return true; } return false; }
  bool str(std::string& out) {
    ws();
    if (p >= end || *p != '"') { err = "expected string"; return false; }
    p++; out.clear();
    while (p < end && *p != '"') {
      if (*p == '\\' && p + 1 < end) {
        p++;
        switch (*p) {
          case 'n': out += '\n'; break; case 't': out += '\t'; break; case 'r': out += '\r'; break;
          case 'b': out += '\b'; break; case 'f': out += '\f'; break;
          case 'u': {  // \uXXXX -> UTF-8 (BMP only; enough for metadata)
            if (p + 4 >= end) { err = "bad \\u"; return false; }
            unsigned cp = (unsigned)strtoul(std::string(p + 1, 4).c_str(), nullptr, 16); p += 4;
            if (cp < 0x80) out += (char)cp;
            else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
            else { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
            break;
          }
          default: out += *p;
        }
        p++;
      } else out += *p++;
    }
    if (p >= end) { err = "unterminated string"; return false; }
    p++; return true;
  }
  bool integer(int64_t& v) {
    ws();
    char* e = nullptr;
    v = strtoll(p, &e, 10);
    if (e == p) { err = "expected integer"; return false; }
    p = e; return true;
  }
  bool skip_value();
};
bool JParser::skip_value() {
  ws();
  if (p >= end) return false;
  if (*p == '"') { std::string s; return str(s); }
  if (*p == '{' || *p == '[') {
    char open = *p, close = open == '{' ? '}' : ']';
    p++;
    if (lit(close)) return true;
    for (;;) {
      if (open == '{') { std::string k; if (!str(k) || !lit(':')) return false; }
      if (!skip_value()) return false;
      if (lit(',')) continue;
      return lit(close);
    }
  }
  while (p < end && *p != ',' && *p != '}' && *p != ']') p++;
  return true;
}
}  // namespace

DseekData::~DseekData() { for (auto& m : maps) munmap(m.first, m.second); }

static std::string load_file(DseekData& d, const std::string& path, bool read_md, bool lock_weights) {
  int fd = open(path.c_str(), O_RDONLY);
  if (fd < 0) return "cannot open " + path;
  struct stat st;
  if (fstat(fd, &st) != 0) { close(fd); return "cannot stat " + path; }
  size_t size = (size_t)st.st_size;
  int flags = MAP_PRIVATE | (lock_weights ? MAP_POPULATE : 0);
  void* data = mmap(nullptr, size, PROT_READ, flags, fd, 0);
  close(fd);
  if (data == MAP_FAILED) return "mmap failed for " + path;
  if (lock_weights) mlock(data, size);
  d.maps.push_back({data, size});
  if (size < 8) return "shard too small: " + path;
  uint64_t hlen;
  memcpy(&hlen, data, 8);
  if (hlen == 0 || hlen > size - 8) return "bad header length in " + path;
  const char* js = (const char*)data + 8;
  const char* bytes = js + hlen;
  const size_t bytes_size = size - 8 - hlen;
  JParser jp{js, js + hlen, ""};
  if (!jp.lit('{')) return "header is not an object: " + path;
  if (jp.lit('}')) return "";
  for (;;) {
    std::string key;
    if (!jp.str(key) || !jp.lit(':')) return "bad header key in " + path + ": " + jp.err;
    if (key == "__metadata__") {
      if (!jp.lit('{')) return "bad __metadata__";
      if (!jp.lit('}')) for (;;) {
        std::string k, v;
        if (!jp.str(k) || !jp.lit(':') || !jp.str(v)) return "bad __metadata__ entry: " + jp.err;
        if (read_md) d.metadata[k] = v;
        if (jp.lit(',')) continue;
        if (!jp.lit('}')) return "bad __metadata__ end";
        break;
      }
    } else {
      DseekTensor t;
      t.name = key;
      int64_t off[2] = {0, 0};
      bool have_off = false;
      if (!jp.lit('{')) return "bad tensor entry " + key;
      for (;;) {
        std::string f;
        if (!jp.str(f) || !jp.lit(':')) return "bad tensor field in " + key;
        if (f == "dtype") { if (!jp.str(t.dtype)) return "bad dtype in " + key; }
        else if (f == "shape") {
          if (!jp.lit('[')) return "bad shape in " + key;
          int n = 0;
          if (!jp.lit(']')) for (;;) {
            int64_t v;
            if (!jp.integer(v)) return "bad shape value in " + key;
            if (n < 4) t.shape[n] = v;
            n++;
            if (jp.lit(',')) continue;
            if (!jp.lit(']')) return "bad shape end in " + key;
            break;
          }
          if (n > 4) return "shape exceeds 4 dimensions: " + key;
        } else if (f == "data_offsets") {
          if (!jp.lit('[') || !jp.integer(off[0]) || !jp.lit(',') || !jp.integer(off[1]) || !jp.lit(']')) return "bad offsets in " + key;
          have_off = true;
        } else if (!jp.skip_value()) return "bad field in " + key;
        if (jp.lit(',')) continue;
        if (!jp.lit('}')) return "bad tensor end " + key;
        break;
      }
      if (!have_off || off[0] < 0 || off[1] <= off[0] || (size_t)off[1] > bytes_size) return "bad offsets for " + key;
      t.data = bytes + off[0];
      t.size = (size_t)(off[1] - off[0]);
      d.tensors[key] = t;
    }
    if (jp.lit(',')) continue;
    if (!jp.lit('}')) return "bad header end in " + path;
    break;
  }
  return "";
}

std::string DseekData::load(const std::string& dirname, bool lock_weights) {
  DIR* dir = opendir(dirname.c_str());
  if (!dir) return "failed to open directory " + dirname;
  std::vector<std::string> files;
  while (struct dirent* e = readdir(dir)) {
    std::string f = e->d_name;
    if (f != "." && f != "..") files.push_back(dirname + "/" + f);
  }
  closedir(dir);
  if (files.empty()) return "no files found in " + dirname;
  std::sort(files.begin(), files.end());
  for (size_t i = 0; i < files.size(); i++) {
    std::string err = load_file(*this, files[i], i == 0, lock_weights);
    if (!err.empty()) return err;
  }
  if (metadata.empty()) return "first shard carries no __metadata__";
  return "";
}
