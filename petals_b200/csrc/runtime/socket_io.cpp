// Framed socket I/O for the control-plane transport (parallel/transport.py).
//
// The reference moves every tensor through a Go libp2p daemon: Python serialises into protobuf bytes, copies them to
// the daemon over a unix socket, and the daemon frames them again.  Here a message is
// [u32 header length][msgpack header][tensor bytes ...] on a unix stream socket, and the two hot loops are native:
//
//   * pb_sock_send_frames: one scatter-gather sendmsg() loop over the header and every tensor's storage — the payload
//     is read straight from the tensors' memory, there is no Python-level bytes object, join or copy;
//   * pb_sock_recv_exact: receives straight into the destination tensor's storage (torch.empty / pinned staging).
//
// Both run without the GIL (ctypes releases it), so a 64 MiB prefill hop does not stall the other handler threads.
// Python sockets with a timeout are non-blocking underneath; EAGAIN is therefore handled with poll() against the
// caller's deadline instead of being an error.
#include <errno.h>
#include <poll.h>
#include <stdint.h>
#include <sys/socket.h>
#include <sys/uio.h>
#include <time.h>

#include <algorithm>
#include <vector>

#include "runtime.h"

namespace {

constexpr int kMaxIov = 64;  // well under IOV_MAX; longer lists are sent in batches

double now_s() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return double(ts.tv_sec) + 1e-9 * double(ts.tv_nsec);
}

// Waits until fd is ready for `events` or the deadline passes. 0 ready, -ETIMEDOUT, -errno.
int wait_ready(int fd, short events, double deadline) {
  for (;;) {
    int timeout_ms = -1;
    if (deadline >= 0) {
      double left = deadline - now_s();
      if (left <= 0) return -ETIMEDOUT;
      timeout_ms = int(std::min(left * 1e3 + 1.0, 2e9));
    }
    pollfd p{fd, events, 0};
    int r = poll(&p, 1, timeout_ms);
    if (r > 0) return 0;
    if (r == 0) return -ETIMEDOUT;
    if (errno != EINTR) return -errno;
  }
}

}  // namespace

extern "C" {

int pb_sock_send_frames(int fd, const void* const* bufs, const int64_t* lens, int n, double timeout_s) {
  const double deadline = timeout_s >= 0 ? now_s() + timeout_s : -1.0;
  std::vector<iovec> iov;
  iov.reserve(n);
  for (int i = 0; i < n; ++i)
    if (lens[i] > 0) iov.push_back(iovec{const_cast<void*>(bufs[i]), size_t(lens[i])});
  size_t first = 0;
  while (first < iov.size()) {
    msghdr msg{};
    msg.msg_iov = &iov[first];
    msg.msg_iovlen = std::min<size_t>(iov.size() - first, kMaxIov);
    ssize_t sent = sendmsg(fd, &msg, MSG_NOSIGNAL);
    if (sent < 0) {
      if (errno == EINTR) continue;
      if (errno == EAGAIN || errno == EWOULDBLOCK) {
        int w = wait_ready(fd, POLLOUT, deadline);
        if (w != 0) return w;
        continue;
      }
      return -errno;
    }
    size_t left = size_t(sent);
    while (left > 0 && first < iov.size()) {  // advance over fully and partially sent segments
      if (left >= iov[first].iov_len) {
        left -= iov[first].iov_len;
        ++first;
      } else {
        iov[first].iov_base = static_cast<char*>(iov[first].iov_base) + left;
        iov[first].iov_len -= left;
        left = 0;
      }
    }
  }
  return 0;
}

// 0 ok; -1 the peer closed the connection before `n` bytes arrived (-1 also when nothing arrived: a clean EOF);
// -ETIMEDOUT; other -errno.
int pb_sock_recv_exact(int fd, void* dst, int64_t n, double timeout_s) {
  const double deadline = timeout_s >= 0 ? now_s() + timeout_s : -1.0;
  char* p = static_cast<char*>(dst);
  int64_t got = 0;
  while (got < n) {
    ssize_t r = recv(fd, p + got, size_t(n - got), 0);
    if (r > 0) {
      got += r;
      continue;
    }
    if (r == 0) return -1;
    if (errno == EINTR) continue;
    if (errno == EAGAIN || errno == EWOULDBLOCK) {
      int w = wait_ready(fd, POLLIN, deadline);
      if (w != 0) return w;
      continue;
    }
    return -errno;
  }
  return 0;
}

}  // extern "C"
