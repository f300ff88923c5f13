// petals_b200 — sparse mixture-of-experts decode path (Mixtral: softmax router, top-k of E experts, SwiGLU experts).
//
// The reference wraps Hugging Face's MixtralSparseMoeBlock: a Python loop over experts with boolean masks, torch.where
// and index_add (src/petals/models/mixtral/block.py:13-19) — host-synchronising, so it cannot live
// in a CUDA graph and reads every expert's routing mask per token. Here the decode step is three sync-free kernels:
//
//   moe_router   : RMSNorm(h) -> router logits -> fp32 softmax -> top-k -> renormalise; writes the normed activations
//                  once, the chosen expert ids and their weights (all on device: graph replayable)
//   moe_gemv     : for every (token, choice) pair, a weight-streaming GEMV whose weight base pointer is selected ON THE
//                  DEVICE from the expert id (gate/up fused with SwiGLU, or down) — only the k chosen experts are read
//   moe_combine  : out = residual + sum_j w_j * y_j
//
// Prefill uses per-expert tcgen05 GEMMs over gathered tokens (server/stage_engine.py::_moe_prefill).
#include "common.cuh"
#include "petals_b200.h"

namespace pb {

PB_DEVICE float rbf16(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// grid: tokens; block: 256. E <= 64, topk <= 8.
__global__ void __launch_bounds__(256) moe_router_kernel(const __nv_bfloat16* __restrict__ h, const __nv_bfloat16* __restrict__ norm_w,
                                                         const __nv_bfloat16* __restrict__ router, __nv_bfloat16* __restrict__ xn_out,
                                                         int* __restrict__ topi, float* __restrict__ topw, int H, int E, int topk, float eps) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  __shared__ float red[32];
  __shared__ float logits[64];
  const int m = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  const __nv_bfloat16* x = h + static_cast<size_t>(m) * H;
  if (norm_w != nullptr) {
    float ss = 0.f;
    for (int k = tid; k < H; k += blockDim.x) { const float f = __bfloat162float(x[k]); ss += f * f; }
    ss = warp_sum(ss);
    if (lane == 0) red[warp] = ss;
    __syncthreads();
    float tot = lane < nw ? red[lane] : 0.f;
    tot = warp_sum(tot);
    const float rstd = rsqrtf(tot / H + eps);
    for (int k = tid; k < H; k += blockDim.x) {
      const __nv_bfloat16 v = __float2bfloat16_rn(rbf16(__bfloat162float(x[k]) * rstd) * __bfloat162float(norm_w[k]));
      xs[k] = v;
      xn_out[static_cast<size_t>(m) * H + k] = v;
    }
  } else {  // the rows are normalised already (tensor-parallel prefill: the all-gather delivers norm outputs)
    for (int k = tid; k < H; k += blockDim.x) xs[k] = x[k];
  }
  __syncthreads();
  for (int e = warp; e < E; e += nw) {
    const __nv_bfloat16* w = router + static_cast<size_t>(e) * H;
    float acc = 0.f;
    for (int k = lane; k < H; k += 32) acc = fmaf(__bfloat162float(xs[k]), __bfloat162float(w[k]), acc);
    acc = warp_sum(acc);
    if (lane == 0) logits[e] = rbf16(acc);  // HF computes the router in bf16, the softmax in fp32
  }
  __syncthreads();
  if (tid == 0) {
    float mx = -INFINITY;
    for (int e = 0; e < E; ++e) mx = fmaxf(mx, logits[e]);
    float p[64], sum = 0.f;
    for (int e = 0; e < E; ++e) { p[e] = __expf(logits[e] - mx); sum += p[e]; }
    float chosen_sum = 0.f;
    int idx[8];
    float val[8];
    for (int j = 0; j < topk; ++j) {
      int best = 0;
      float bv = -1.f;
      for (int e = 0; e < E; ++e)
        if (p[e] > bv) { bv = p[e]; best = e; }
      idx[j] = best; val[j] = bv / sum; chosen_sum += val[j];
      p[best] = -2.f;
    }
    for (int j = 0; j < topk; ++j) {
      topi[m * topk + j] = idx[j];
      topw[m * topk + j] = rbf16(val[j] / chosen_sum);
    }
  }
}

PB_DEVICE void fma8f(float& acc, const uint4& w, const uint4& xv) {
  acc = fmaf(bf16_lo(w.x), bf16_lo(xv.x), acc); acc = fmaf(bf16_hi(w.x), bf16_hi(xv.x), acc);
  acc = fmaf(bf16_lo(w.y), bf16_lo(xv.y), acc); acc = fmaf(bf16_hi(w.y), bf16_hi(xv.y), acc);
  acc = fmaf(bf16_lo(w.z), bf16_lo(xv.z), acc); acc = fmaf(bf16_hi(w.z), bf16_hi(xv.z), acc);
  acc = fmaf(bf16_lo(w.w), bf16_lo(xv.w), acc); acc = fmaf(bf16_hi(w.w), bf16_hi(xv.w), acc);
}

// grid: (ctas, pairs). Pair p uses expert topi[p]; its input row is x[p / x_row_div].
template <bool DUAL>
__global__ void __launch_bounds__(512, 2) moe_gemv_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w_all,
                                                          const __nv_bfloat16* __restrict__ w2_all, const int* __restrict__ topi,
                                                          __nv_bfloat16* __restrict__ out, int N, int K, size_t expert_stride, int x_row_div) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  const int p = blockIdx.y;
  const int e = topi[p];
  const __nv_bfloat16* xrow = x + static_cast<size_t>(p / x_row_div) * K;
  for (int v = threadIdx.x; v < (K >> 3); v += blockDim.x) reinterpret_cast<uint4*>(xs)[v] = __ldcg(reinterpret_cast<const uint4*>(xrow) + v);
  __syncthreads();
  const __nv_bfloat16* w = w_all + static_cast<size_t>(e) * expert_stride;
  const __nv_bfloat16* w2 = DUAL ? w2_all + static_cast<size_t>(e) * expert_stride : nullptr;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  constexpr int U = DUAL ? 2 : 4;
  const int ntasks = N >> 1, total_warps = gridDim.x * nwarps, kstep = 256 * U;
  for (int task = warp * gridDim.x + blockIdx.x; task < ntasks; task += total_warps) {
    const int n0 = task << 1;
    const __nv_bfloat16* r0 = w + static_cast<size_t>(n0) * K;
    const __nv_bfloat16* r1 = r0 + K;
    const __nv_bfloat16* q0 = DUAL ? w2 + static_cast<size_t>(n0) * K : nullptr;
    const __nv_bfloat16* q1 = DUAL ? q0 + K : nullptr;
    float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
    for (int kb = 0; kb < K; kb += kstep) {
      uint4 wa[U], wb[U], ua[U], ub[U];
      bool ok[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int k = kb + u * 256 + lane * 8;
        ok[u] = k < K;
        if (ok[u]) {
          wa[u] = ld_stream(r0 + k); wb[u] = ld_stream(r1 + k);
          if (DUAL) { ua[u] = ld_stream(q0 + k); ub[u] = ld_stream(q1 + k); }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (ok[u]) {
          const uint4 xv = *reinterpret_cast<const uint4*>(xs + kb + u * 256 + lane * 8);
          fma8f(a0, wa[u], xv); fma8f(a1, wb[u], xv);
          if (DUAL) { fma8f(b0, ua[u], xv); fma8f(b1, ub[u], xv); }
        }
      }
    }
    a0 = warp_sum(a0); a1 = warp_sum(a1);
    if (DUAL) { b0 = warp_sum(b0); b1 = warp_sum(b1); }
    if (lane == 0) {
      float v0 = a0, v1 = a1;
      if (DUAL) {
        v0 = rbf16(rbf16(v0) / (1.f + __expf(-rbf16(v0)))) * rbf16(b0);
        v1 = rbf16(rbf16(v1) / (1.f + __expf(-rbf16(v1)))) * rbf16(b1);
      }
      *reinterpret_cast<uint32_t*>(out + static_cast<size_t>(p) * N + n0) = pack_bf16(v0, v1);
    }
  }
}

// out[m, :] = residual[m, :] + sum_j bf16(y[m*topk + j, :] * w[m*topk + j])
__global__ void moe_combine_kernel(const __nv_bfloat16* __restrict__ y, const float* __restrict__ topw, const __nv_bfloat16* __restrict__ residual,
                                   __nv_bfloat16* __restrict__ out, int H, int topk) {
  const int m = blockIdx.x;
  for (int k = threadIdx.x; k < H; k += blockDim.x) {
    float acc = 0.f;
    for (int j = 0; j < topk; ++j) {
      const int p = m * topk + j;
      acc = rbf16(acc + rbf16(__bfloat162float(y[static_cast<size_t>(p) * H + k]) * topw[p]));
    }
    out[static_cast<size_t>(m) * H + k] = __float2bfloat16_rn((residual != nullptr ? __bfloat162float(residual[static_cast<size_t>(m) * H + k]) : 0.f) + acc);
  }
}

}  // namespace pb

using namespace pb;

// ---- prefill-sized MoE without a host round trip --------------------------------------------------------------------------
// The reference loops over the experts in Python with boolean masks and index_add (HF MixtralSparseMoeBlock, wrapped at
// src/petals/models/mixtral/block.py:13-19): one host synchronisation per layer to learn the group sizes. Here the routing stays
// on the device: moe_plan_kernel turns topi[M*k] into (a) the destination row of every (token, expert) pair in expert-major order
// and (b) the tile table of the grouped tcgen05 GEMM (gemm_tcgen05.cu, `grp`): for every expert its row range cut into 128-row tiles.
// One CTA; E <= 64 experts.
__global__ void __launch_bounds__(1024) moe_plan_kernel(const int* __restrict__ topi, int pairs, int E, int* __restrict__ pos, int* __restrict__ table,
                                                        int cap) {
  __shared__ int cnt[64], off[64], cur[64];
  const int tid = threadIdx.x;
  if (tid < 64) { cnt[tid] = 0; cur[tid] = 0; }
  __syncthreads();
  for (int p = tid; p < pairs; p += blockDim.x) {
    const int e = topi[p];
    if (e >= 0 && e < E) atomicAdd(&cnt[e], 1);
  }
  __syncthreads();
  if (tid == 0) {
    int acc = 0, nt = 0;
    int* t_e = table + 1; int* t_row0 = t_e + cap; int* t_rows = t_row0 + cap;
    for (int e = 0; e < E; ++e) {
      off[e] = acc;
      for (int r = 0; r < cnt[e] && nt < cap; r += 128, ++nt) { t_e[nt] = e; t_row0[nt] = acc + r; t_rows[nt] = min(128, cnt[e] - r); }
      acc += cnt[e];
    }
    table[0] = nt;
  }
  __syncthreads();
  for (int p = tid; p < pairs; p += blockDim.x) {
    const int e = topi[p];
    pos[p] = (e >= 0 && e < E) ? off[e] + atomicAdd(&cur[e], 1) : 0;
  }
}

// gathered[pos[p]] = x[p / topk]   (one CTA per pair, 16-byte vectors)
__global__ void moe_gather_kernel(const uint4* __restrict__ x, const int* __restrict__ pos, uint4* __restrict__ out, int vec_per_row, int topk) {
  const int p = blockIdx.x;
  const uint4* src = x + static_cast<size_t>(p / topk) * vec_per_row;
  uint4* dst = out + static_cast<size_t>(pos[p]) * vec_per_row;
  for (int i = threadIdx.x; i < vec_per_row; i += blockDim.x) dst[i] = src[i];
}

// out[m] = residual[m] + sum_j topw[m, j] * y[pos[m * topk + j]]   (HF rounding as in moe_combine_kernel)
__global__ void moe_combine_pos_kernel(const __nv_bfloat16* __restrict__ y, const float* __restrict__ topw, const int* __restrict__ pos,
                                       const __nv_bfloat16* __restrict__ residual, __nv_bfloat16* __restrict__ out, int H, int topk) {
  const int m = blockIdx.x;
  for (int k = threadIdx.x; k < H; k += blockDim.x) {
    float acc = 0.f;
    for (int j = 0; j < topk; ++j) {
      const int p = m * topk + j;
      acc = rbf16(acc + rbf16(__bfloat162float(y[static_cast<size_t>(pos[p]) * H + k]) * topw[p]));
    }
    out[static_cast<size_t>(m) * H + k] = __float2bfloat16_rn((residual != nullptr ? __bfloat162float(residual[static_cast<size_t>(m) * H + k]) : 0.f) + acc);
  }
}

extern "C" int pb_moe_router(const void* h, const void* norm_w, const void* router, void* xn_out, void* topi, void* topw, int M, int H, int E,
                             int topk, float eps, void* stream) {
  if (M == 0) return PB_OK;
  if (E > 64 || topk > 8 || topk > E || H * 2 > 200 * 1024) return PB_ERR_SHAPE;
  auto k = moe_router_kernel;
  const size_t smem = static_cast<size_t>(H) * 2;
  if (smem > 32 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  k<<<M, 256, smem, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(h), static_cast<const __nv_bfloat16*>(norm_w),
                                                         static_cast<const __nv_bfloat16*>(router), static_cast<__nv_bfloat16*>(xn_out),
                                                         static_cast<int*>(topi), static_cast<float*>(topw), H, E, topk, eps);
  return pb_check_launch("moe_router");
}

extern "C" int pb_moe_gemv(const void* x, const void* w_all, const void* w2_all, const void* topi, void* out, int pairs, int N, int K,
                           long expert_stride, int x_row_div, int num_sms, void* stream) {
  if (pairs == 0) return PB_OK;
  if ((N & 1) || (K & 7) || K * 2 > 100 * 1024 || x_row_div < 1) return PB_ERR_SHAPE;
  const size_t smem = static_cast<size_t>(K) * 2;
  const int sms = num_sms > 0 ? num_sms : 148;
  dim3 grid(sms, pairs);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (w2_all != nullptr) {
    auto k = moe_gemv_kernel<true>;
    if (smem > 32 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    k<<<grid, 512, smem, s>>>(static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(w_all), static_cast<const __nv_bfloat16*>(w2_all),
                              static_cast<const int*>(topi), static_cast<__nv_bfloat16*>(out), N, K, static_cast<size_t>(expert_stride), x_row_div);
  } else {
    auto k = moe_gemv_kernel<false>;
    if (smem > 32 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    k<<<grid, 512, smem, s>>>(static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(w_all), nullptr, static_cast<const int*>(topi),
                              static_cast<__nv_bfloat16*>(out), N, K, static_cast<size_t>(expert_stride), x_row_div);
  }
  return pb_check_launch("moe_gemv");
}

extern "C" int pb_moe_combine(const void* y, const void* topw, const void* residual, void* out, int M, int H, int topk, void* stream) {
  if (M == 0) return PB_OK;
  moe_combine_kernel<<<M, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(y), static_cast<const float*>(topw),
                                                                      static_cast<const __nv_bfloat16*>(residual), static_cast<__nv_bfloat16*>(out), H, topk);
  return pb_check_launch("moe_combine");
}

extern "C" int pb_moe_plan(const void* topi, int pairs, int E, void* pos, void* table, int cap, void* stream) {
  if (E > 64 || E <= 0 || cap <= 0) return PB_ERR_UNSUPPORTED;
  if (pairs == 0) return PB_OK;
  moe_plan_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const int*>(topi), pairs, E, static_cast<int*>(pos), static_cast<int*>(table), cap);
  return pb_check_launch("moe_plan");
}
extern "C" int pb_moe_gather(const void* x, const void* pos, void* out, int pairs, int H, int topk, void* stream) {
  if (H % 8 != 0) return PB_ERR_SHAPE;
  if (pairs == 0) return PB_OK;
  moe_gather_kernel<<<pairs, 128, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint4*>(x), static_cast<const int*>(pos), static_cast<uint4*>(out), H / 8, topk);
  return pb_check_launch("moe_gather");
}
extern "C" int pb_moe_combine_pos(const void* y, const void* topw, const void* pos, const void* residual, void* out, int M, int H, int topk, void* stream) {
  if (M == 0) return PB_OK;
  moe_combine_pos_kernel<<<M, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(y), static_cast<const float*>(topw),
                                                                           static_cast<const int*>(pos), static_cast<const __nv_bfloat16*>(residual),
                                                                           static_cast<__nv_bfloat16*>(out), H, topk);
  return pb_check_launch("moe_combine_pos");
}
