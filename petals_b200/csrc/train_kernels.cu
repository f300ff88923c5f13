// petals_b200 — row-wise / element-wise pieces of a transformer block's backward pass (activation gradients only).
//
// `rpc_backward` differentiates frozen blocks with respect to their inputs and deep prompts (reference:
// src/petals/server/block_functions.py:84-141; the reference runs torch.autograd over eager ATen ops).
// The GEMM-shaped part is the tcgen05 GEMM with the weight consumed untransposed (gemm_tcgen05.cu, MN-major B) and the
// attention part is attention_bwd.cu; this file holds what sits between them. All tensors are bf16, all math is fp32.
#include "common.cuh"
#include "petals_b200.h"

extern "C" int pb_set_error(const char* msg);

namespace pb {

// y = (x * rstd) * w  with rstd = rsqrt(mean(x^2) + eps)
// dx = rstd * (dy*w) - x * rstd^3 * mean(dy*w*x)        (+ d_res: the gradient arriving over the residual connection)
// One CTA per row; the row lives in registers between the two reductions (H <= 8 * 256 * VEC).
template <int VEC>  // 16-byte vectors per thread
__global__ void __launch_bounds__(256) rmsnorm_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                                                         const __nv_bfloat16* __restrict__ w, const __nv_bfloat16* d_res,
                                                         __nv_bfloat16* dx_out, int H, float eps) {
  __shared__ float red[2][8];
  const size_t row = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nvec = H >> 3;
  float xs[VEC][8], gs[VEC][8];
  float ss = 0.f, sg = 0.f;
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    const int i = tid + v * 256;
    if (i < nvec) {
      const uint4 xv = *reinterpret_cast<const uint4*>(x + row * H + i * 8);
      const uint4 dv = *reinterpret_cast<const uint4*>(dy + row * H + i * 8);
      const uint4 wv = __ldg(reinterpret_cast<const uint4*>(w + i * 8));
      const uint32_t xa[4] = {xv.x, xv.y, xv.z, xv.w}, da[4] = {dv.x, dv.y, dv.z, dv.w}, wa[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        xs[v][2 * k] = bf16_lo(xa[k]); xs[v][2 * k + 1] = bf16_hi(xa[k]);
        gs[v][2 * k] = bf16_lo(da[k]) * bf16_lo(wa[k]); gs[v][2 * k + 1] = bf16_hi(da[k]) * bf16_hi(wa[k]);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) { ss += xs[v][k] * xs[v][k]; sg += gs[v][k] * xs[v][k]; }
    }
  }
  ss = warp_sum(ss); sg = warp_sum(sg);
  if (lane == 0) { red[0][warp] = ss; red[1][warp] = sg; }
  __syncthreads();
  float tss = 0.f, tsg = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { tss += red[0][i]; tsg += red[1][i]; }
  const float rstd = rsqrtf(tss / H + eps);
  const float coef = rstd * rstd * rstd * tsg / H;
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    const int i = tid + v * 256;
    if (i < nvec) {
      float o[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = rstd * gs[v][k] - coef * xs[v][k];
      if (d_res != nullptr) {
        const uint4 rv = *reinterpret_cast<const uint4*>(d_res + row * H + i * 8);
        o[0] += bf16_lo(rv.x); o[1] += bf16_hi(rv.x); o[2] += bf16_lo(rv.y); o[3] += bf16_hi(rv.y);
        o[4] += bf16_lo(rv.z); o[5] += bf16_hi(rv.z); o[6] += bf16_lo(rv.w); o[7] += bf16_hi(rv.w);
      }
      uint4 ov;
      ov.x = pack_bf16(o[0], o[1]); ov.y = pack_bf16(o[2], o[3]); ov.z = pack_bf16(o[4], o[5]); ov.w = pack_bf16(o[6], o[7]);
      *reinterpret_cast<uint4*>(dx_out + row * H + i * 8) = ov;
    }
  }
}

// act = silu(g) * u  =>  dg = d_act * u * sigma(g) * (1 + g * (1 - sigma(g))),  du = d_act * silu(g); written over g and u.
__global__ void __launch_bounds__(256) swiglu_bwd_kernel(const __nv_bfloat16* __restrict__ d_act, __nv_bfloat16* g, __nv_bfloat16* u, long nvec) {
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const uint4 dv = *reinterpret_cast<const uint4*>(d_act + i * 8);
    uint4 gv = *reinterpret_cast<const uint4*>(g + i * 8), uv = *reinterpret_cast<const uint4*>(u + i * 8);
    const uint32_t da[4] = {dv.x, dv.y, dv.z, dv.w};
    uint32_t ga[4] = {gv.x, gv.y, gv.z, gv.w}, ua[4] = {uv.x, uv.y, uv.z, uv.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float r[2][2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float gg = h ? bf16_hi(ga[k]) : bf16_lo(ga[k]), uu = h ? bf16_hi(ua[k]) : bf16_lo(ua[k]), dd = h ? bf16_hi(da[k]) : bf16_lo(da[k]);
        const float sg = 1.f / (1.f + __expf(-gg));
        r[0][h] = dd * uu * sg * (1.f + gg * (1.f - sg));
        r[1][h] = dd * gg * sg;
      }
      ga[k] = pack_bf16(r[0][0], r[0][1]);
      ua[k] = pack_bf16(r[1][0], r[1][1]);
    }
    *reinterpret_cast<uint4*>(g + i * 8) = make_uint4(ga[0], ga[1], ga[2], ga[3]);
    *reinterpret_cast<uint4*>(u + i * 8) = make_uint4(ua[0], ua[1], ua[2], ua[3]);
  }
}

// The forward rotates a pair (x0, x1) = (x[i], x[i + D/2]) into (x0 c - x1 s, x1 c + x0 s); its transpose maps the gradients
// (g0, g1) to (g0 c + g1 s, g1 c - g0 s). grid = (rows, Hq + 2 Hkv), block = D/2 threads.
__global__ void qkv_grad_merge_kernel(const __nv_bfloat16* __restrict__ dq, const __nv_bfloat16* __restrict__ dk, const __nv_bfloat16* __restrict__ dv,
                                      const float* __restrict__ cos, const float* __restrict__ sin, __nv_bfloat16* __restrict__ dqkv,
                                      int T, int Hq, int Hkv, int D, int max_pos) {
  const int row = blockIdx.x, h = blockIdx.y, i = threadIdx.x, half = D >> 1;
  const __nv_bfloat16* src;
  if (h < Hq) src = dq + (static_cast<size_t>(row) * Hq + h) * D;
  else if (h < Hq + Hkv) src = dk + (static_cast<size_t>(row) * Hkv + (h - Hq)) * D;
  else src = dv + (static_cast<size_t>(row) * Hkv + (h - Hq - Hkv)) * D;
  float g0 = __bfloat162float(src[i]), g1 = __bfloat162float(src[i + half]);
  if (h < Hq + Hkv && cos != nullptr) {
    int pos = row % T;
    pos = pos < max_pos ? pos : max_pos - 1;
    const float c = cos[static_cast<size_t>(pos) * half + i], s = sin[static_cast<size_t>(pos) * half + i];
    const float a = g0 * c + g1 * s, b = g1 * c - g0 * s;
    g0 = a; g1 = b;
  }
  __nv_bfloat16* dst = dqkv + (static_cast<size_t>(row) * (Hq + 2 * Hkv) + h) * D;
  dst[i] = __float2bfloat16_rn(g0);
  dst[i + half] = __float2bfloat16_rn(g1);
}

}  // namespace pb

using namespace pb;

extern "C" int pb_rmsnorm_bwd(const void* dy, const void* x, const void* w, const void* d_res, void* dx_out, int rows, int H, float eps, void* stream) {
  if (rows == 0) return PB_OK;
  if (H % 8 != 0 || H > 8 * 256 * 8) { pb_set_error("rmsnorm_bwd: H must be a multiple of 8 and at most 16384"); return PB_ERR_SHAPE; }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int nvec = H / 8, vec = (nvec + 255) / 256;
  auto a = static_cast<const __nv_bfloat16*>(dy); auto b = static_cast<const __nv_bfloat16*>(x); auto c = static_cast<const __nv_bfloat16*>(w);
  auto d = static_cast<const __nv_bfloat16*>(d_res); auto o = static_cast<__nv_bfloat16*>(dx_out);
  if (vec <= 1) rmsnorm_bwd_kernel<1><<<rows, 256, 0, s>>>(a, b, c, d, o, H, eps);
  else if (vec <= 2) rmsnorm_bwd_kernel<2><<<rows, 256, 0, s>>>(a, b, c, d, o, H, eps);
  else if (vec <= 4) rmsnorm_bwd_kernel<4><<<rows, 256, 0, s>>>(a, b, c, d, o, H, eps);
  else rmsnorm_bwd_kernel<8><<<rows, 256, 0, s>>>(a, b, c, d, o, H, eps);
  return pb_check_launch("rmsnorm_bwd");
}

extern "C" int pb_swiglu_bwd(const void* d_act, void* g, void* u, long n, void* stream) {
  if (n == 0) return PB_OK;
  if (n % 8 != 0) { pb_set_error("swiglu_bwd: element count must be a multiple of 8"); return PB_ERR_SHAPE; }
  const long nvec = n / 8;
  const int blocks = static_cast<int>(nvec / 256 + 1 < 148L * 16 ? nvec / 256 + 1 : 148L * 16);
  swiglu_bwd_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(d_act), static_cast<__nv_bfloat16*>(g),
                                                                           static_cast<__nv_bfloat16*>(u), nvec);
  return pb_check_launch("swiglu_bwd");
}

extern "C" int pb_qkv_grad_merge(const void* dq, const void* dk, const void* dv, const void* cos, const void* sin, void* dqkv, int M, int T, int Hq,
                                 int Hkv, int D, int max_pos, void* stream) {
  if (M == 0) return PB_OK;
  if (D % 2 != 0 || D > 2048 || T <= 0) { pb_set_error("qkv_grad_merge: bad head_dim / T"); return PB_ERR_SHAPE; }
  qkv_grad_merge_kernel<<<dim3(M, Hq + 2 * Hkv), D / 2, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(dq), static_cast<const __nv_bfloat16*>(dk), static_cast<const __nv_bfloat16*>(dv),
      static_cast<const float*>(cos), static_cast<const float*>(sin), static_cast<__nv_bfloat16*>(dqkv), T, Hq, Hkv, D, max_pos);
  return pb_check_launch("qkv_grad_merge");
}
