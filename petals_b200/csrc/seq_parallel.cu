// Sequence-parallel glue for tensor-parallel PREFILL: the owner-side half of a fused reduce-scatter -> norm -> all-gather.
//
// A tensor-parallel layer over R ranks with M prompt rows runs, per row-parallel projection (attention out / MLP down):
//
//   GEMM (tcgen05, gemm_tcgen05.cu)   epilogue stores each output row ONLY into its owner's `parts[src]` buffer over NVLink
//                                     (push_rows_per_owner routing) -> that is the reduce-scatter, overlapped tile by tile
//                                     with the MMA main loop; the last CTA publishes one flag increment per peer.
//   norm_reduce_gather (this file)    the owner waits for R increments, sums the R partial rows + its residual slice in
//                                     fp32, writes the new residual slice locally, applies RMSNorm/LayerNorm and stores
//                                     the normalised rows into EVERY rank's activation buffer over NVLink -> the all-gather;
//                                     the last CTA publishes one flag increment per peer.
//   next column-parallel GEMM         waits for R increments on its activation buffer's flag in its prologue.
//
// No NCCL call, no host involvement and no separate collective kernel: the residual stream stays row-sharded
// (1/R of the rows per rank), norms are computed once per row instead of R times, and every byte crosses NVLink once.
// Reference behaviour being replaced: tensor_parallel's all-reduce after each row-parallel linear
// (src/petals/utils/convert_block.py:118-135).
#include "common.cuh"
#include "petals_b200.h"

namespace pb {

constexpr int kMaxPeers = PB_MAX_PEERS;
constexpr int kNrgThreads = 256;
constexpr int kNrgMaxVec = 8;  // per-thread uint4 slots: H <= 256 * 8 * 8 = 16384

struct NrgParams {
  const uint4* x_res_in;        // [rows, H] bf16 residual rows owned by this rank (nullptr: zero)
  uint4* x_res_out;             // [rows, H] updated residual (may alias x_res_in; nullptr: not stored)
  const uint4* parts[kMaxPeers];
  int n_parts;
  const __nv_bfloat16* norm_w;  // nullptr with norm_kind 0
  const __nv_bfloat16* norm_b;
  float eps;
  int norm_kind;                // 0 none (gather the raw sum), 1 RMS, 2 LayerNorm
  uint4* gather_out[kMaxPeers]; // destination of THIS rank's row 0 on every peer
  uint64_t* gather_flag[kMaxPeers];
  int n_gather;
  const uint64_t* wait_flag;
  uint64_t wait_per_epoch;
  const uint64_t* epoch;
  unsigned int* done_counter;
  int* error_flag;
  int rows, H;
};

__global__ void __launch_bounds__(kNrgThreads) norm_reduce_gather_kernel(const NrgParams p) {
  __shared__ float red[2][kNrgThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (p.wait_flag != nullptr) {
    if (tid == 0 && !spin_wait_ge(p.wait_flag, *p.epoch * p.wait_per_epoch)) atomicExch(p.error_flag, 1);
    __syncthreads();
  }
  const int nvec = p.H >> 3;  // uint4 per row
  for (int row = blockIdx.x; row < p.rows; row += gridDim.x) {
    const size_t base = static_cast<size_t>(row) * nvec;
    float f[kNrgMaxVec][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < kNrgMaxVec; ++j) {
      const int v = tid + j * kNrgThreads;
      if (v < nvec) {
        // issue every load of this slot first (residual + all partials), then add: an in-order warp that consumed each load right
        // after issuing it paid (1 + n_parts) serial memory latencies per slot
        uint4 ld[kMaxPeers + 1];
        ld[0] = p.x_res_in != nullptr ? __ldcg(p.x_res_in + base + v) : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int r = 0; r < kMaxPeers; ++r)
          if (r < p.n_parts) ld[r + 1] = __ldcg(p.parts[r] + base + v);
        f[j][0] = bf16_lo(ld[0].x); f[j][1] = bf16_hi(ld[0].x); f[j][2] = bf16_lo(ld[0].y); f[j][3] = bf16_hi(ld[0].y);
        f[j][4] = bf16_lo(ld[0].z); f[j][5] = bf16_hi(ld[0].z); f[j][6] = bf16_lo(ld[0].w); f[j][7] = bf16_hi(ld[0].w);
#pragma unroll
        for (int r = 0; r < kMaxPeers; ++r) {
          if (r < p.n_parts) {
            const uint4 y = ld[r + 1];
            f[j][0] += bf16_lo(y.x); f[j][1] += bf16_hi(y.x); f[j][2] += bf16_lo(y.y); f[j][3] += bf16_hi(y.y);
            f[j][4] += bf16_lo(y.z); f[j][5] += bf16_hi(y.z); f[j][6] += bf16_lo(y.w); f[j][7] += bf16_hi(y.w);
          }
        }
        // the residual stream is carried in bf16 (as the dense model does): round first, normalise the rounded value
        uint4 o;
        o.x = pack_bf16(f[j][0], f[j][1]); o.y = pack_bf16(f[j][2], f[j][3]); o.z = pack_bf16(f[j][4], f[j][5]); o.w = pack_bf16(f[j][6], f[j][7]);
        if (p.x_res_out != nullptr) p.x_res_out[base + v] = o;
        f[j][0] = bf16_lo(o.x); f[j][1] = bf16_hi(o.x); f[j][2] = bf16_lo(o.y); f[j][3] = bf16_hi(o.y);
        f[j][4] = bf16_lo(o.z); f[j][5] = bf16_hi(o.z); f[j][6] = bf16_lo(o.w); f[j][7] = bf16_hi(o.w);
#pragma unroll
        for (int i = 0; i < 8; ++i) { s1 += f[j][i]; s2 += f[j][i] * f[j][i]; }
      }
    }
    float mean = 0.f, rstd = 1.f;
    if (p.norm_kind != 0) {
      s1 = warp_sum(s1); s2 = warp_sum(s2);
      if (lane == 0) { red[0][warp] = s1; red[1][warp] = s2; }
      __syncthreads();
      float t1 = 0.f, t2 = 0.f;
#pragma unroll
      for (int w = 0; w < kNrgThreads / 32; ++w) { t1 += red[0][w]; t2 += red[1][w]; }
      __syncthreads();
      const float inv_h = 1.f / static_cast<float>(p.H);
      if (p.norm_kind == 1) {
        rstd = rsqrtf(t2 * inv_h + p.eps);
      } else {
        mean = t1 * inv_h;
        rstd = rsqrtf(fmaxf(t2 * inv_h - mean * mean, 0.f) + p.eps);
      }
    }
#pragma unroll
    for (int j = 0; j < kNrgMaxVec; ++j) {
      const int v = tid + j * kNrgThreads;
      if (v < nvec) {
        uint4 o;
        if (p.norm_kind != 0) {
          const uint4 wv = __ldg(reinterpret_cast<const uint4*>(p.norm_w) + v);
          const float w[8] = {bf16_lo(wv.x), bf16_hi(wv.x), bf16_lo(wv.y), bf16_hi(wv.y), bf16_lo(wv.z), bf16_hi(wv.z), bf16_lo(wv.w), bf16_hi(wv.w)};
          float b[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          if (p.norm_b != nullptr) {
            const uint4 bv = __ldg(reinterpret_cast<const uint4*>(p.norm_b) + v);
            b[0] = bf16_lo(bv.x); b[1] = bf16_hi(bv.x); b[2] = bf16_lo(bv.y); b[3] = bf16_hi(bv.y);
            b[4] = bf16_lo(bv.z); b[5] = bf16_hi(bv.z); b[6] = bf16_lo(bv.w); b[7] = bf16_hi(bv.w);
          }
          float g[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) g[i] = (f[j][i] - mean) * rstd * w[i] + b[i];
          o.x = pack_bf16(g[0], g[1]); o.y = pack_bf16(g[2], g[3]); o.z = pack_bf16(g[4], g[5]); o.w = pack_bf16(g[6], g[7]);
        } else {
          o.x = pack_bf16(f[j][0], f[j][1]); o.y = pack_bf16(f[j][2], f[j][3]); o.z = pack_bf16(f[j][4], f[j][5]); o.w = pack_bf16(f[j][6], f[j][7]);
        }
        for (int r = 0; r < p.n_gather; ++r) p.gather_out[r][base + v] = o;
      }
    }
  }
  // publish: the last CTA to finish increments every peer's flag once
  if (p.n_gather > 0) {
    __syncthreads();
    if (tid == 0) {
      __threadfence_system();
      const unsigned int prev = atomicAdd(p.done_counter, 1u);
      if (prev == gridDim.x - 1) {
        __threadfence_system();
        *p.done_counter = 0u;
        for (int r = 0; r < p.n_gather; ++r)
          if (p.gather_flag[r] != nullptr) red_release_sys_add(p.gather_flag[r], 1ull);
      }
    }
  }
}

}  // namespace pb

using namespace pb;

extern "C" int pb_norm_reduce_gather(const PbNormReduceGatherArgs* a, void* stream) {
  if (a->H <= 0 || (a->H & 7) || a->H > kNrgThreads * kNrgMaxVec * 8 || a->rows < 0) return PB_ERR_SHAPE;
  if (a->n_parts < 0 || a->n_parts > kMaxPeers || a->n_gather < 0 || a->n_gather > kMaxPeers) return PB_ERR_SHAPE;
  if (a->norm_kind != 0 && a->norm_w == nullptr) return PB_ERR_SHAPE;
  if (a->n_gather > 0 && a->done_counter == nullptr) return PB_ERR_SHAPE;
  NrgParams p;
  p.x_res_in = static_cast<const uint4*>(a->x_res_in);
  p.x_res_out = static_cast<uint4*>(a->x_res_out);
  for (int i = 0; i < kMaxPeers; ++i) {
    p.parts[i] = i < a->n_parts ? static_cast<const uint4*>(a->parts[i]) : nullptr;
    p.gather_out[i] = i < a->n_gather ? static_cast<uint4*>(a->gather_out[i]) : nullptr;
    p.gather_flag[i] = i < a->n_gather ? static_cast<uint64_t*>(a->gather_flag[i]) : nullptr;
  }
  p.n_parts = a->n_parts;
  p.n_gather = a->n_gather;
  p.norm_w = static_cast<const __nv_bfloat16*>(a->norm_w);
  p.norm_b = static_cast<const __nv_bfloat16*>(a->norm_b);
  p.eps = a->eps;
  p.norm_kind = a->norm_kind;
  p.wait_flag = static_cast<const uint64_t*>(a->wait_flag);
  p.wait_per_epoch = a->wait_per_epoch;
  p.epoch = static_cast<const uint64_t*>(a->epoch);
  p.done_counter = static_cast<unsigned int*>(a->done_counter);
  p.error_flag = static_cast<int*>(a->error_flag);
  p.rows = a->rows;
  p.H = a->H;
  if (p.wait_flag != nullptr && (p.epoch == nullptr || p.error_flag == nullptr)) return PB_ERR_SHAPE;
  const int sms = a->num_sms > 0 ? a->num_sms : 148;
  int grid = a->rows < sms * 4 ? a->rows : sms * 4;  // 4 x 256-thread CTAs per SM are co-resident: the flag spin cannot deadlock
  if (grid < 1) grid = 1;
  norm_reduce_gather_kernel<<<grid, kNrgThreads, 0, static_cast<cudaStream_t>(stream)>>>(p);
  return pb_check_launch("norm_reduce_gather");
}
