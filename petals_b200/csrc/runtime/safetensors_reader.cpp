// Minimal safetensors reader: mmap the file, parse the JSON header, hand out tensor extents and do
// multi-threaded copies into (pinned) staging buffers. Native replacement for the Rust `safetensors`
// dependency on the per-block weight loading path (src/petals/server/from_pretrained.py:216-224).
#include "runtime.h"

#include <fcntl.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <string>
#include <thread>
#include <vector>

namespace {
thread_local std::string g_err;

struct TensorInfo {
  std::string name, dtype;
  std::vector<int64_t> shape;
  int64_t begin = 0, end = 0;
};
struct StFile {
  int fd = -1;
  uint8_t* map = nullptr;
  size_t size = 0;
  size_t data_off = 0;
  std::vector<TensorInfo> tensors;
};

// --- tiny JSON scanner (objects / arrays / strings / integers are all the header contains) ---------
struct Scanner {
  const char* p;
  const char* e;
  void ws() { while (p < e && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p; }
  bool eat(char c) { ws(); if (p < e && *p == c) { ++p; return true; } return false; }
  bool str(std::string& out) {
    ws();
    if (p >= e || *p != '"') return false;
    ++p;
    out.clear();
    while (p < e && *p != '"') {
      if (*p == '\\' && p + 1 < e) {
        ++p;
        switch (*p) {
          case 'n': out.push_back('\n'); break;
          case 't': out.push_back('\t'); break;
          case 'u': out.push_back('?'); p += 4; break;
          default: out.push_back(*p);
        }
        ++p;
      } else {
        out.push_back(*p++);
      }
    }
    if (p >= e) return false;
    ++p;
    return true;
  }
  bool integer(int64_t& v) {
    ws();
    const char* s = p;
    bool neg = false;
    if (p < e && *p == '-') { neg = true; ++p; }
    int64_t x = 0;
    const char* d = p;
    while (p < e && *p >= '0' && *p <= '9') x = x * 10 + (*p++ - '0');
    if (p == d) { p = s; return false; }
    v = neg ? -x : x;
    return true;
  }
  bool skip_value() {  // skip any JSON value
    ws();
    if (p >= e) return false;
    if (*p == '"') { std::string t; return str(t); }
    if (*p == '{' || *p == '[') {
      const char open = *p, close = open == '{' ? '}' : ']';
      ++p;
      ws();
      if (eat(close)) return true;
      while (true) {
        if (open == '{') { std::string k; if (!str(k) || !eat(':')) return false; }
        if (!skip_value()) return false;
        if (eat(',')) continue;
        return eat(close);
      }
    }
    while (p < e && *p != ',' && *p != '}' && *p != ']') ++p;
    return true;
  }
};

bool parse_header(StFile* f, const char* js, size_t n) {
  Scanner s{js, js + n};
  if (!s.eat('{')) return false;
  if (s.eat('}')) return true;
  while (true) {
    std::string key;
    if (!s.str(key) || !s.eat(':')) return false;
    if (key == "__metadata__") {
      if (!s.skip_value()) return false;
    } else {
      TensorInfo t;
      t.name = key;
      if (!s.eat('{')) return false;
      while (true) {
        std::string k;
        if (!s.str(k) || !s.eat(':')) return false;
        if (k == "dtype") {
          if (!s.str(t.dtype)) return false;
        } else if (k == "shape") {
          if (!s.eat('[')) return false;
          if (!s.eat(']')) {
            while (true) {
              int64_t v;
              if (!s.integer(v)) return false;
              t.shape.push_back(v);
              if (s.eat(',')) continue;
              if (!s.eat(']')) return false;
              break;
            }
          }
        } else if (k == "data_offsets") {
          if (!s.eat('[') || !s.integer(t.begin) || !s.eat(',') || !s.integer(t.end) || !s.eat(']')) return false;
        } else if (!s.skip_value()) {
          return false;
        }
        if (s.eat(',')) continue;
        if (!s.eat('}')) return false;
        break;
      }
      f->tensors.push_back(std::move(t));
    }
    if (s.eat(',')) continue;
    return s.eat('}');
  }
}
}  // namespace

extern "C" const char* pb_st_error(void) { return g_err.c_str(); }

extern "C" void* pb_st_open(const char* path) {
  auto* f = new StFile();
  f->fd = ::open(path, O_RDONLY);
  if (f->fd < 0) { g_err = std::string("cannot open ") + path; delete f; return nullptr; }
  struct stat st;
  if (fstat(f->fd, &st) != 0 || st.st_size < 8) { g_err = "bad file size"; ::close(f->fd); delete f; return nullptr; }
  f->size = static_cast<size_t>(st.st_size);
  void* m = mmap(nullptr, f->size, PROT_READ, MAP_PRIVATE, f->fd, 0);
  if (m == MAP_FAILED) { g_err = "mmap failed"; ::close(f->fd); delete f; return nullptr; }
  f->map = static_cast<uint8_t*>(m);
  uint64_t hlen = 0;
  memcpy(&hlen, f->map, 8);
  if (f->size < 8 || hlen > f->size - 8) { g_err = "header length exceeds file"; pb_st_close(f); return nullptr; }
  f->data_off = 8 + hlen;
  if (!parse_header(f, reinterpret_cast<const char*>(f->map + 8), hlen)) {
    g_err = "malformed safetensors header";
    pb_st_close(f);
    return nullptr;
  }
  for (auto& t : f->tensors)
    if (t.begin < 0 || t.end < t.begin || f->data_off + static_cast<size_t>(t.end) > f->size) {
      g_err = "tensor extent outside file: " + t.name;
      pb_st_close(f);
      return nullptr;
    }
  std::sort(f->tensors.begin(), f->tensors.end(), [](const TensorInfo& a, const TensorInfo& b) { return a.name < b.name; });
  return f;
}
extern "C" void pb_st_close(void* h) {
  auto* f = static_cast<StFile*>(h);
  if (!f) return;
  if (f->map) munmap(f->map, f->size);
  if (f->fd >= 0) ::close(f->fd);
  delete f;
}
extern "C" int pb_st_num_tensors(void* h) { return static_cast<int>(static_cast<StFile*>(h)->tensors.size()); }
extern "C" int pb_st_tensor_info(void* h, int idx, char* name, int name_cap, char* dtype, int dtype_cap, int64_t* shape,
                                 int64_t* data_offset, int64_t* nbytes) {
  auto* f = static_cast<StFile*>(h);
  if (idx < 0 || idx >= static_cast<int>(f->tensors.size())) return -1;
  const TensorInfo& t = f->tensors[idx];
  snprintf(name, name_cap, "%s", t.name.c_str());
  snprintf(dtype, dtype_cap, "%s", t.dtype.c_str());
  const int nd = static_cast<int>(std::min<size_t>(t.shape.size(), 8));
  for (int i = 0; i < nd; ++i) shape[i] = t.shape[i];
  *data_offset = t.begin;
  *nbytes = t.end - t.begin;
  return nd;
}
extern "C" int pb_st_find(void* h, const char* name) {
  auto* f = static_cast<StFile*>(h);
  auto it = std::lower_bound(f->tensors.begin(), f->tensors.end(), std::string(name),
                             [](const TensorInfo& a, const std::string& n) { return a.name < n; });
  if (it == f->tensors.end() || it->name != name) return -1;
  return static_cast<int>(it - f->tensors.begin());
}
extern "C" const void* pb_st_data(void* h) {
  auto* f = static_cast<StFile*>(h);
  return f->map + f->data_off;
}
extern "C" int pb_st_read(void* h, int idx, void* dst, int64_t cap, int threads) {
  auto* f = static_cast<StFile*>(h);
  if (idx < 0 || idx >= static_cast<int>(f->tensors.size())) return -1;
  const TensorInfo& t = f->tensors[idx];
  const int64_t n = t.end - t.begin;
  if (n > cap) return -2;
  const uint8_t* src = f->map + f->data_off + t.begin;
  if (threads <= 1 || n < (8 << 20)) {
    memcpy(dst, src, static_cast<size_t>(n));
    return 0;
  }
  std::vector<std::thread> pool;
  const int64_t chunk = (n + threads - 1) / threads;
  for (int i = 0; i < threads; ++i) {
    const int64_t b = i * chunk, e = std::min<int64_t>(n, b + chunk);
    if (b >= e) break;
    pool.emplace_back([=] { memcpy(static_cast<uint8_t*>(dst) + b, src + b, static_cast<size_t>(e - b)); });
  }
  for (auto& th : pool) th.join();
  return 0;
}
