// ThreadSanitizer stress of the host runtime (KV page allocator + prioritised task queue): the two pieces of native code that
// handler threads and the stage runtime share. Built and run by tests/test_native_runtime_tsan.py:
//   g++ -std=c++17 -O1 -g -fsanitize=thread kv_allocator.cpp task_queue.cpp tests/stress_tsan.cpp -lpthread
// Exit code 0 and no "WARNING: ThreadSanitizer" in the output = pass.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "../runtime.h"

static std::atomic<long> g_errors{0};
#define CHECK(cond)                                                      \
  do {                                                                   \
    if (!(cond)) { std::fprintf(stderr, "CHECK failed: %s (line %d)\n", #cond, __LINE__); g_errors++; } \
  } while (0)

int main() {
  // ---- KV allocator: sessions reserve, allocate, share (incref), free — from 8 threads ------------------------------------
  const int kPages = 256, kThreads = 8, kIters = 2000;
  void* kv = pb_kv_create(kPages);
  std::vector<std::thread> th;
  for (int t = 0; t < kThreads; ++t) {
    th.emplace_back([&, t] {
      unsigned seed = 1234u + t;
      for (int i = 0; i < kIters; ++i) {
        const int n = 1 + static_cast<int>(rand_r(&seed) % 6);
        if (pb_kv_reserve(kv, n, 0.05) != 0) continue;  // admission control with a short timeout
        int pages[8];
        if (pb_kv_alloc(kv, n, pages) == 0) {
          for (int j = 0; j < n; ++j) CHECK(pages[j] >= 0 && pages[j] < kPages);
          if (rand_r(&seed) & 1) {  // a beam fork shares the pages for a while
            pb_kv_incref(kv, pages, n);
            for (int j = 0; j < n; ++j) CHECK(pb_kv_refcount(kv, pages[j]) >= 2);
            pb_kv_free(kv, pages, n);
          }
          pb_kv_free(kv, pages, n);
        }
        pb_kv_unreserve(kv, n);
      }
    });
  }
  for (auto& x : th) x.join();
  th.clear();
  CHECK(pb_kv_num_free(kv) == kPages);
  CHECK(pb_kv_reserved(kv) == 0);
  pb_kv_destroy(kv);

  // ---- task queue: 4 producers (two priority classes), 3 consumers, then close -----------------------------------------------
  void* tq = pb_tq_create();
  const int kProducers = 4, kPerProducer = 3000;
  std::atomic<long> popped{0}, sum{0};
  std::vector<std::thread> cons;
  for (int c = 0; c < 3; ++c) {
    cons.emplace_back([&] {
      for (;;) {
        int64_t id = -1;
        double prio = 0;
        const int rc = pb_tq_pop(tq, 0.2, &id, &prio);
        if (rc == -2) return;   // closed and drained
        if (rc == -1) continue; // timeout
        CHECK(prio == 1.0 || prio == 2.0);
        popped++;
        sum += id;
      }
    });
  }
  for (int p = 0; p < kProducers; ++p) {
    th.emplace_back([&, p] {
      for (int i = 0; i < kPerProducer; ++i) pb_tq_push(tq, (i & 1) ? 1.0 : 2.0, static_cast<int64_t>(p) * kPerProducer + i);
    });
  }
  for (auto& x : th) x.join();
  while (pb_tq_size(tq) > 0) std::this_thread::yield();
  pb_tq_close(tq);
  for (auto& x : cons) x.join();
  const long n = static_cast<long>(kProducers) * kPerProducer;
  CHECK(popped.load() == n);
  CHECK(sum.load() == n * (n - 1) / 2);
  pb_tq_destroy(tq);

  if (g_errors.load()) { std::fprintf(stderr, "%ld checks failed\n", g_errors.load()); return 1; }
  std::puts("runtime stress ok");
  return 0;
}
