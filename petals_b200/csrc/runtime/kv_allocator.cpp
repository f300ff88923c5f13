// Refcounted KV-cache page allocator + token-budget reservation with timeout.
// Native replacement for the reference's cross-process byte-budget MemoryCache
// (src/petals/server/memory_cache.py:71-167): sessions first *reserve* their max_length worth of pages
// (admission control with alloc_timeout, FIFO fair), then bind physical pages lazily as tokens arrive.
// Refcounts let beam-search hypotheses share full pages (copy-on-write of the last partial page only),
// replacing the whole-cache gather of src/petals/server/backend.py:154-158.
#include "runtime.h"

#include <chrono>
#include <condition_variable>
#include <mutex>
#include <set>
#include <vector>

namespace {
struct KvAllocator {
  std::mutex mu;
  std::condition_variable cv;
  std::vector<int> free_list;
  std::vector<int> refs;
  long reserved = 0;  // pages promised to sessions (>= pages actually bound)
  long total = 0;
  uint64_t next_ticket = 0, serving = 0;  // FIFO among waiting reservations
  std::set<uint64_t> abandoned;           // tickets whose owner timed out before their turn: skipped when `serving` reaches them
  void advance() {                        // retire the ticket being served and every abandoned one right behind it
    ++serving;
    for (auto it = abandoned.find(serving); it != abandoned.end(); it = abandoned.find(serving)) { abandoned.erase(it); ++serving; }
    cv.notify_all();
  }
};
}  // namespace

extern "C" void* pb_kv_create(int num_pages) {
  if (num_pages < 0) return nullptr;  // a C ABI must not let std::length_error escape
  auto* a = new KvAllocator();
  a->total = num_pages;
  a->refs.assign(num_pages, 0);
  a->free_list.reserve(num_pages);
  for (int i = num_pages - 1; i >= 0; --i) a->free_list.push_back(i);
  return a;
}
extern "C" void pb_kv_destroy(void* h) { delete static_cast<KvAllocator*>(h); }

extern "C" int pb_kv_alloc(void* h, int n, int* out) {
  auto* a = static_cast<KvAllocator*>(h);
  std::lock_guard<std::mutex> g(a->mu);
  if (static_cast<int>(a->free_list.size()) < n) return -1;
  for (int i = 0; i < n; ++i) {
    int p = a->free_list.back();
    a->free_list.pop_back();
    a->refs[p] = 1;
    out[i] = p;
  }
  return 0;
}
extern "C" void pb_kv_incref(void* h, const int* pages, int n) {
  auto* a = static_cast<KvAllocator*>(h);
  std::lock_guard<std::mutex> g(a->mu);
  for (int i = 0; i < n; ++i) {
    const int p = pages[i];
    if (p < 0 || p >= static_cast<int>(a->refs.size()) || a->refs[p] <= 0) continue;  // only live pages can gain an owner
    a->refs[p]++;
  }
}
extern "C" void pb_kv_free(void* h, const int* pages, int n) {
  auto* a = static_cast<KvAllocator*>(h);
  std::lock_guard<std::mutex> g(a->mu);
  for (int i = 0; i < n; ++i) {
    int p = pages[i];
    if (p < 0 || p >= static_cast<int>(a->refs.size()) || a->refs[p] <= 0) continue;
    if (--a->refs[p] == 0) a->free_list.push_back(p);
  }
}
extern "C" int pb_kv_num_free(void* h) {
  auto* a = static_cast<KvAllocator*>(h);
  std::lock_guard<std::mutex> g(a->mu);
  return static_cast<int>(a->free_list.size());
}
extern "C" int pb_kv_refcount(void* h, int page) {
  auto* a = static_cast<KvAllocator*>(h);
  std::lock_guard<std::mutex> g(a->mu);
  if (page < 0 || page >= a->total) return -1;
  return a->refs[page];
}
extern "C" int pb_kv_reserve(void* h, long pages, double timeout_s) {
  auto* a = static_cast<KvAllocator*>(h);
  std::unique_lock<std::mutex> lk(a->mu);
  if (pages < 0 || pages > a->total) return -1;
  const uint64_t ticket = a->next_ticket++;
  const auto deadline = std::chrono::steady_clock::now() + std::chrono::duration<double>(timeout_s < 0 ? 0 : timeout_s);
  auto ready = [&] { return a->serving == ticket && a->reserved + pages <= a->total; };
  bool ok = ready();
  if (!ok && timeout_s > 0) ok = a->cv.wait_until(lk, deadline, ready);
  if (!ok) {
    // give up WITHOUT waiting for our turn (a fail-fast caller with timeout 0 must return at once even when another
    // reservation is queued ahead): the ticket retires out of order and is skipped when the queue reaches it
    if (a->serving == ticket) a->advance();
    else a->abandoned.insert(ticket);
    return -1;
  }
  a->reserved += pages;
  a->advance();
  return 0;
}
extern "C" void pb_kv_unreserve(void* h, long pages) {
  auto* a = static_cast<KvAllocator*>(h);
  {
    std::lock_guard<std::mutex> g(a->mu);
    if (pages > 0) a->reserved -= pages;
    if (a->reserved < 0) a->reserved = 0;
  }
  a->cv.notify_all();
}
extern "C" long pb_kv_reserved(void* h) {
  auto* a = static_cast<KvAllocator*>(h);
  std::lock_guard<std::mutex> g(a->mu);
  return a->reserved;
}
