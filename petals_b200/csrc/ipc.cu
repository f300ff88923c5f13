// petals_b200 — NVLink symmetric memory plumbing and point-to-point primitives.
//
// One process per GPU. Every rank cudaMalloc's one "symmetric heap" and exports it with CUDA IPC; every
// rank maps all peers' heaps, so a (rank, offset) pair is a device pointer usable from any kernel.
// Data moves with plain st.global to those peer pointers issued from GEMV/GEMM epilogues (see
// linear_decode.cu / gemm_tcgen05.cu) and is published with a release-increment of a 64-bit flag in the
// peer's heap; consumers spin with ld.acquire.sys under a wall-clock watchdog (common.cuh).
// This replaces the reference's whole wire stack for activations: protobuf + libp2p daemon + TCP +
// host staging (src/petals/server/task_pool.py:119-142).
//
// Kernels here: row broadcast (+flag), flag wait, arg-max candidate exchange for vocab-parallel sampling,
// and the probes that fill the roofline denominators / hop-latency numbers (peer copy bandwidth, ping-pong).
#include "common.cuh"
#include "petals_b200.h"

#include <string.h>

namespace pb {

// copy `n_vec` 16-byte vectors from src to up to 8 destinations (peer pointers), then one release per peer.
__global__ void __launch_bounds__(512) push_rows_kernel(const uint4* __restrict__ src, uint4* d0, uint4* d1, uint4* d2, uint4* d3,
                                                        uint4* d4, uint4* d5, uint4* d6, uint4* d7, int n_dst, long n_vec,
                                                        uint64_t* f0, uint64_t* f1, uint64_t* f2, uint64_t* f3, uint64_t* f4,
                                                        uint64_t* f5, uint64_t* f6, uint64_t* f7) {
  uint4* dst[8] = {d0, d1, d2, d3, d4, d5, d6, d7};
  uint64_t* flg[8] = {f0, f1, f2, f3, f4, f5, f6, f7};
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n_vec; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const uint4 v = src[i];
    for (int r = 0; r < n_dst; ++r) dst[r][i] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    for (int r = 0; r < n_dst; ++r)
      if (flg[r] != nullptr) red_release_sys_add(flg[r], 1ull);
  }
}

__global__ void wait_flag_kernel(const uint64_t* flag, const uint64_t* epoch, uint64_t per_epoch, uint64_t absolute, int* error_flag) {
  const uint64_t target = epoch != nullptr ? *epoch * per_epoch : absolute;
  if (!spin_wait_ge(flag, target)) atomicExch(error_flag, 1);
}

// vocab-parallel greedy sampling: each rank found (value, index) over its vocabulary slice; every rank writes
// its candidate into slot [rank] of every peer; after the flag wait each rank reduces the R candidates.
__global__ void argmax_publish_kernel(const float* val, const long long* idx, long long idx_offset, int rows, int rank, int n_peers,
                                      float* c0, float* c1, float* c2, float* c3, float* c4, float* c5, float* c6, float* c7,
                                      uint64_t* f0, uint64_t* f1, uint64_t* f2, uint64_t* f3, uint64_t* f4, uint64_t* f5,
                                      uint64_t* f6, uint64_t* f7, int max_rows) {
  float* cand[8] = {c0, c1, c2, c3, c4, c5, c6, c7};
  uint64_t* flg[8] = {f0, f1, f2, f3, f4, f5, f6, f7};
  const int t = threadIdx.x;
  if (t < rows) {
    const float v = val[t];
    for (int r = 0; r < n_peers; ++r) {
      // layout per peer: [R][max_rows][2] floats: value, index-as-int bits
      float* slot = cand[r] + (static_cast<size_t>(rank) * max_rows + t) * 2;
      slot[0] = v;
      slot[1] = __int_as_float(static_cast<int>(idx[t] + idx_offset));
    }
  }
  __syncthreads();
  if (t == 0) {
    __threadfence_system();
    for (int r = 0; r < n_peers; ++r) red_release_sys_add(flg[r], 1ull);
  }
}

__global__ void argmax_reduce_kernel(const float* cand, int n_peers, int rows, int max_rows, const uint64_t* flag, const uint64_t* epoch,
                                     long long* out_ids, int* error_flag) {
  if (threadIdx.x == 0) {
    if (!spin_wait_ge(flag, *epoch * static_cast<uint64_t>(n_peers))) atomicExch(error_flag, 1);
  }
  __syncthreads();
  const int t = threadIdx.x;
  if (t < rows) {
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int r = 0; r < n_peers; ++r) {
      const float* slot = cand + (static_cast<size_t>(r) * max_rows + t) * 2;
      const float v = __ldcg(slot);
      const int i = __float_as_int(__ldcg(slot + 1));
      if (v > best || (v == best && i < bi)) { best = v; bi = i; }
    }
    out_ids[t] = bi == 0x7fffffff ? 0 : bi;
  }
}

// per-row (value, index) arg-max over a vocabulary slice (bf16 logits) — the local half of the exchange above
__global__ void __launch_bounds__(1024) argmax_val_kernel(const __nv_bfloat16* __restrict__ logits, float* out_val, long long* out_idx, int vocab) {
  __shared__ float sval[32];
  __shared__ int sidx[32];
  const __nv_bfloat16* row = logits + static_cast<size_t>(blockIdx.x) * vocab;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < vocab; i += blockDim.x) {
    const float v = __bfloat162float(row[i]);
    if (v > best || (v == best && i < bi)) { best = v; bi = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (lane == 0) { sval[warp] = best; sidx[warp] = bi; }
  __syncthreads();
  if (warp == 0) {
    best = lane < nw ? sval[lane] : -INFINITY;
    bi = lane < nw ? sidx[lane] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (lane == 0) { out_val[blockIdx.x] = best; out_idx[blockIdx.x] = bi == 0x7fffffff ? 0 : bi; }
  }
}

// out = residual + sum_r parts[r]  (the tail of a one-shot all-reduce whose producers were GEMV epilogues)
__global__ void __launch_bounds__(256) reduce_parts_kernel(const uint4* __restrict__ res, const uint4* p0, const uint4* p1, const uint4* p2,
                                                           const uint4* p3, const uint4* p4, const uint4* p5, const uint4* p6, const uint4* p7,
                                                           int n_parts, const uint64_t* wait_flag, uint64_t per_epoch, const uint64_t* epoch,
                                                           uint4* __restrict__ out, long n_vec, int* error_flag) {
  const uint4* parts[8] = {p0, p1, p2, p3, p4, p5, p6, p7};
  if (wait_flag != nullptr) {
    if (threadIdx.x == 0 && !spin_wait_ge(wait_flag, *epoch * per_epoch)) atomicExch(error_flag, 1);
    __syncthreads();
  }
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n_vec; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const uint4 x = __ldcg(res + i);
    float f[8] = {bf16_lo(x.x), bf16_hi(x.x), bf16_lo(x.y), bf16_hi(x.y), bf16_lo(x.z), bf16_hi(x.z), bf16_lo(x.w), bf16_hi(x.w)};
    for (int r = 0; r < n_parts; ++r) {
      const uint4 v = __ldcg(parts[r] + i);
      f[0] += bf16_lo(v.x); f[1] += bf16_hi(v.x); f[2] += bf16_lo(v.y); f[3] += bf16_hi(v.y);
      f[4] += bf16_lo(v.z); f[5] += bf16_hi(v.z); f[6] += bf16_lo(v.w); f[7] += bf16_hi(v.w);
    }
    uint4 o;
    o.x = pack_bf16(f[0], f[1]); o.y = pack_bf16(f[2], f[3]); o.z = pack_bf16(f[4], f[5]); o.w = pack_bf16(f[6], f[7]);
    out[i] = o;
  }
}

// ---- probes ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) peer_copy_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, long n_vec) {
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n_vec; i += static_cast<long>(gridDim.x) * blockDim.x) dst[i] = src[i];
}

// Ping-pong between two ranks on one flag pair; the initiator records round-trip times with %globaltimer.
__global__ void pingpong_kernel(uint64_t* my_flag, uint64_t* peer_flag, int iters, int initiator, unsigned long long* rtt_ns, int* error_flag) {
  for (int i = 1; i <= iters; ++i) {
    if (initiator) {
      const unsigned long long t0 = globaltimer_ns();
      st_release_sys(peer_flag, static_cast<uint64_t>(i));
      if (!spin_wait_ge(my_flag, static_cast<uint64_t>(i))) { atomicExch(error_flag, 1); return; }
      rtt_ns[i - 1] = globaltimer_ns() - t0;
    } else {
      if (!spin_wait_ge(my_flag, static_cast<uint64_t>(i))) { atomicExch(error_flag, 1); return; }
      st_release_sys(peer_flag, static_cast<uint64_t>(i));
    }
  }
}

}  // namespace pb

using namespace pb;

extern "C" int pb_ipc_malloc(void** out, long bytes) {
  if (cudaMalloc(out, static_cast<size_t>(bytes)) != cudaSuccess) return PB_ERR_CUDA;
  if (cudaMemset(*out, 0, static_cast<size_t>(bytes)) != cudaSuccess) return PB_ERR_CUDA;
  return cudaDeviceSynchronize() == cudaSuccess ? PB_OK : PB_ERR_CUDA;
}
extern "C" int pb_ipc_free(void* p) { return cudaFree(p) == cudaSuccess ? PB_OK : PB_ERR_CUDA; }
extern "C" int pb_ipc_get_handle(void* p, void* out64) {
  cudaIpcMemHandle_t h;
  if (cudaIpcGetMemHandle(&h, p) != cudaSuccess) return PB_ERR_CUDA;
  memcpy(out64, &h, sizeof(h));
  return PB_OK;
}
extern "C" int pb_ipc_open_handle(const void* in64, void** out) {
  cudaIpcMemHandle_t h;
  memcpy(&h, in64, sizeof(h));
  cudaError_t e = cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) { cudaGetLastError(); return PB_ERR_CUDA; }
  return PB_OK;
}
extern "C" int pb_ipc_close_handle(void* p) { return cudaIpcCloseMemHandle(p) == cudaSuccess ? PB_OK : PB_ERR_CUDA; }
extern "C" int pb_ipc_handle_size(void) { return static_cast<int>(sizeof(cudaIpcMemHandle_t)); }

extern "C" int pb_push_rows(const void* src, void* const* dsts, void* const* flags, int n_dst, long bytes, void* stream) {
  if (n_dst < 1 || n_dst > 8 || (bytes & 15)) return PB_ERR_SHAPE;
  uint4* d[8] = {nullptr};
  uint64_t* f[8] = {nullptr};
  for (int i = 0; i < n_dst; ++i) { d[i] = static_cast<uint4*>(dsts[i]); f[i] = static_cast<uint64_t*>(flags[i]); }
  // ONE CTA: the release must follow every store of the payload; payloads here are activations of a few rows
  // (decode) — large tensors go through the GEMM epilogue push instead.
  push_rows_kernel<<<1, 512, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint4*>(src), d[0], d[1], d[2], d[3], d[4], d[5], d[6],
                                                                    d[7], n_dst, bytes >> 4, f[0], f[1], f[2], f[3], f[4], f[5], f[6], f[7]);
  return pb_check_launch("ipc");
}

extern "C" int pb_wait_flag(const void* flag, const void* epoch, uint64_t per_epoch, uint64_t absolute, void* error_flag, void* stream) {
  wait_flag_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint64_t*>(flag), static_cast<const uint64_t*>(epoch),
                                                                  per_epoch, absolute, static_cast<int*>(error_flag));
  return pb_check_launch("ipc");
}

extern "C" int pb_argmax_val(const void* logits, void* out_val, void* out_idx, int rows, int vocab, void* stream) {
  if (rows == 0) return PB_OK;
  argmax_val_kernel<<<rows, 1024, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(logits), static_cast<float*>(out_val),
                                                                         static_cast<long long*>(out_idx), vocab);
  return pb_check_launch("ipc");
}

extern "C" int pb_argmax_exchange(const void* val, const void* idx, long idx_offset, int rows, int rank, int n_peers, void* const* cands,
                                  void* const* flags, int max_rows, const void* my_cand, const void* my_flag, const void* epoch, void* out_ids,
                                  void* error_flag, void* stream) {
  if (n_peers < 1 || n_peers > 8 || rows > max_rows || rows > 32) return PB_ERR_SHAPE;
  float* c[8] = {nullptr};
  uint64_t* f[8] = {nullptr};
  for (int i = 0; i < n_peers; ++i) { c[i] = static_cast<float*>(cands[i]); f[i] = static_cast<uint64_t*>(flags[i]); }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  argmax_publish_kernel<<<1, 32, 0, s>>>(static_cast<const float*>(val), static_cast<const long long*>(idx), idx_offset, rows, rank, n_peers, c[0],
                                         c[1], c[2], c[3], c[4], c[5], c[6], c[7], f[0], f[1], f[2], f[3], f[4], f[5], f[6], f[7], max_rows);
  argmax_reduce_kernel<<<1, 32, 0, s>>>(static_cast<const float*>(my_cand), n_peers, rows, max_rows, static_cast<const uint64_t*>(my_flag),
                                        static_cast<const uint64_t*>(epoch), static_cast<long long*>(out_ids), static_cast<int*>(error_flag));
  return pb_check_launch("ipc");
}

extern "C" int pb_reduce_parts(const void* res, void* const* parts, int n_parts, const void* wait_flag, uint64_t per_epoch, const void* epoch,
                               void* out, long bytes, void* error_flag, void* stream) {
  if (n_parts < 0 || n_parts > 8 || (bytes & 15)) return PB_ERR_SHAPE;
  const uint4* p[8] = {nullptr};
  for (int i = 0; i < n_parts; ++i) p[i] = static_cast<const uint4*>(parts[i]);
  const long n_vec = bytes >> 4;
  int grid = static_cast<int>((n_vec + 255) / 256);
  if (grid > 64) grid = 64;
  reduce_parts_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint4*>(res), p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7],
                                                                          n_parts, static_cast<const uint64_t*>(wait_flag), per_epoch,
                                                                          static_cast<const uint64_t*>(epoch), static_cast<uint4*>(out), n_vec,
                                                                          static_cast<int*>(error_flag));
  return pb_check_launch("reduce_parts");
}

extern "C" int pb_peer_copy(const void* src, void* dst, long bytes, int ctas, void* stream) {
  if (bytes & 15) return PB_ERR_SHAPE;
  peer_copy_kernel<<<ctas > 0 ? ctas : 148 * 4, 512, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint4*>(src), static_cast<uint4*>(dst), bytes >> 4);
  return pb_check_launch("ipc");
}

extern "C" int pb_pingpong(void* my_flag, void* peer_flag, int iters, int initiator, void* rtt_ns, void* error_flag, void* stream) {
  pingpong_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<uint64_t*>(my_flag), static_cast<uint64_t*>(peer_flag), iters, initiator,
                                                                 static_cast<unsigned long long*>(rtt_ns), static_cast<int*>(error_flag));
  return pb_check_launch("ipc");
}
