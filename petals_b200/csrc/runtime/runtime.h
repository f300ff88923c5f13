// petals_b200 host runtime — C ABI (libpetals_b200_rt.so). No CUDA dependency: loads on CPU-only hosts.
#pragma once
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

// ---- KV page allocator (refcounted pages; copy-on-write support for beam search) -------------------
void* pb_kv_create(int num_pages);
void pb_kv_destroy(void* h);
int pb_kv_alloc(void* h, int n, int* out_pages);          // 0 ok, -1 not enough free pages (nothing taken)
void pb_kv_incref(void* h, const int* pages, int n);
void pb_kv_free(void* h, const int* pages, int n);         // decref; page returns to the pool at 0
int pb_kv_num_free(void* h);
int pb_kv_refcount(void* h, int page);
// Reserve / release a token budget without binding pages (session admission control).
int pb_kv_reserve(void* h, long pages, double timeout_s);  // 0 ok, -1 timeout
void pb_kv_unreserve(void* h, long pages);
long pb_kv_reserved(void* h);

// ---- prioritised task queue (inference < forward/backward; FIFO among equals) --------------------------
void* pb_tq_create(void);
void pb_tq_destroy(void* h);
void pb_tq_push(void* h, double priority, int64_t task_id);
int pb_tq_pop(void* h, double timeout_s, int64_t* task_id, double* priority);  // 0 ok, -1 timeout, -2 closed
int pb_tq_size(void* h);
void pb_tq_close(void* h);

// ---- safetensors reader (mmap + JSON header) --------------------------------------------------------------
void* pb_st_open(const char* path);
void pb_st_close(void* h);
int pb_st_num_tensors(void* h);
// Fills name/dtype (NUL-terminated, truncated to cap), shape (up to 8 dims). Returns ndim or -1.
int pb_st_tensor_info(void* h, int idx, char* name, int name_cap, char* dtype, int dtype_cap, int64_t* shape,
                      int64_t* data_offset, int64_t* nbytes);
int pb_st_find(void* h, const char* name);                   // index or -1
const void* pb_st_data(void* h);                             // base of the tensor data section (mmap)
// Multi-threaded copy of one tensor's bytes into dst (e.g. a pinned staging buffer).
int pb_st_read(void* h, int idx, void* dst, int64_t cap, int threads);
const char* pb_st_error(void);

// ---- framed socket I/O of the control transport (scatter-gather send, receive into place; no GIL) ---------------
// timeout_s < 0 blocks forever. Return 0 ok, -ETIMEDOUT, other -errno; recv: -1 = peer closed the connection.
int pb_sock_send_frames(int fd, const void* const* bufs, const int64_t* lens, int n, double timeout_s);
int pb_sock_recv_exact(int fd, void* dst, int64_t n, double timeout_s);

#ifdef __cplusplus
}
#endif
