// Prioritised blocking task queue for the per-GPU stage scheduler thread.
// Semantics of the reference's PrioritizedTaskPool + Runtime selection
// (src/petals/server/task_pool.py:78-86,158-167, task_prioritizer.py:15-20): the runnable task with the
// smallest (priority, submission order) goes first; inference (1.0) beats forward/backward (2.0);
// no batching across requests.
#include "runtime.h"

#include <chrono>
#include <condition_variable>
#include <limits>
#include <mutex>
#include <queue>
#include <vector>

namespace {
struct Item {
  double priority;
  uint64_t seq;
  int64_t id;
};
struct Cmp {
  bool operator()(const Item& a, const Item& b) const {
    if (a.priority != b.priority) return a.priority > b.priority;
    return a.seq > b.seq;
  }
};
struct TaskQueue {
  std::mutex mu;
  std::condition_variable cv;
  std::priority_queue<Item, std::vector<Item>, Cmp> q;
  uint64_t seq = 0;
  bool closed = false;
};
}  // namespace

extern "C" void* pb_tq_create(void) { return new TaskQueue(); }
extern "C" void pb_tq_destroy(void* h) { delete static_cast<TaskQueue*>(h); }
extern "C" void pb_tq_push(void* h, double priority, int64_t id) {
  auto* t = static_cast<TaskQueue*>(h);
  {
    std::lock_guard<std::mutex> g(t->mu);
    if (priority != priority) priority = std::numeric_limits<double>::infinity();  // NaN would break the heap's ordering: run it last
    t->q.push(Item{priority, t->seq++, id});
  }
  t->cv.notify_one();
}
extern "C" int pb_tq_pop(void* h, double timeout_s, int64_t* id, double* priority) {
  auto* t = static_cast<TaskQueue*>(h);
  std::unique_lock<std::mutex> lk(t->mu);
  auto ready = [&] { return !t->q.empty() || t->closed; };
  if (!ready()) {
    if (timeout_s <= 0) return -1;
    if (!t->cv.wait_for(lk, std::chrono::duration<double>(timeout_s), ready)) return -1;
  }
  if (t->q.empty()) return -2;
  Item it = t->q.top();
  t->q.pop();
  *id = it.id;
  if (priority) *priority = it.priority;
  return 0;
}
extern "C" int pb_tq_size(void* h) {
  auto* t = static_cast<TaskQueue*>(h);
  std::lock_guard<std::mutex> g(t->mu);
  return static_cast<int>(t->q.size());
}
extern "C" void pb_tq_close(void* h) {
  auto* t = static_cast<TaskQueue*>(h);
  {
    std::lock_guard<std::mutex> g(t->mu);
    t->closed = true;
  }
  t->cv.notify_all();
}
