// petals_b200 — memory-bound helpers: RMSNorm / LayerNorm (+residual), SwiGLU, add, embedding gather,
// arg-max sampling, deep-prompt injection, device-side step counters.
// All are single-pass, 128-bit vectorised, bf16 in / bf16 out with fp32 math.
// Reference behaviour: HF LlamaRMSNorm as used at src/petals/models/llama/block.py:143-155 (G2/G3 in
// SURVEY.md §2.5(a)), deep prompts at src/petals/server/backend.py:231-233 (L14).
#include "common.cuh"
#include "petals_b200.h"

namespace pb {

// one CTA per row; cols % 8 == 0
template <int KIND>
__global__ void __launch_bounds__(256) norm_kernel(const __nv_bfloat16* __restrict__ x,
                                                   const __nv_bfloat16* __restrict__ residual,
                                                   const __nv_bfloat16* __restrict__ weight,
                                                   const __nv_bfloat16* __restrict__ bias,
                                                   __nv_bfloat16* __restrict__ out,
                                                   __nv_bfloat16* __restrict__ sum_out, int cols, float eps) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint4* row = reinterpret_cast<uint4*>(smem_raw);  // staged (x + residual) as bf16
  __shared__ float red[32];
  __shared__ float stats[2];
  const size_t base = static_cast<size_t>(blockIdx.x) * cols;
  const int nvec = cols >> 3;
  float s1 = 0.f, s2 = 0.f;
  for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
    uint4 xv = __ldg(reinterpret_cast<const uint4*>(x + base) + v);
    if (residual != nullptr) {
      const uint4 rv = __ldg(reinterpret_cast<const uint4*>(residual + base) + v);
      xv.x = pack_bf16(bf16_lo(xv.x) + bf16_lo(rv.x), bf16_hi(xv.x) + bf16_hi(rv.x));
      xv.y = pack_bf16(bf16_lo(xv.y) + bf16_lo(rv.y), bf16_hi(xv.y) + bf16_hi(rv.y));
      xv.z = pack_bf16(bf16_lo(xv.z) + bf16_lo(rv.z), bf16_hi(xv.z) + bf16_hi(rv.z));
      xv.w = pack_bf16(bf16_lo(xv.w) + bf16_lo(rv.w), bf16_hi(xv.w) + bf16_hi(rv.w));
      if (sum_out != nullptr) reinterpret_cast<uint4*>(sum_out + base)[v] = xv;
    }
    row[v] = xv;
    const float f[8] = {bf16_lo(xv.x), bf16_hi(xv.x), bf16_lo(xv.y), bf16_hi(xv.y),
                        bf16_lo(xv.z), bf16_hi(xv.z), bf16_lo(xv.w), bf16_hi(xv.w)};
#pragma unroll
    for (int i = 0; i < 8; ++i) { s1 += f[i]; s2 += f[i] * f[i]; }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  s1 = warp_sum(s1); s2 = warp_sum(s2);
  if (lane == 0) { red[warp] = s1; red[16 + warp] = s2; }
  __syncthreads();
  if (warp == 0) {
    float a = lane < nw ? red[lane] : 0.f, b = lane < nw ? red[16 + lane] : 0.f;
    a = warp_sum(a); b = warp_sum(b);
    if (lane == 0) {
      if (KIND == 1) { stats[0] = 0.f; stats[1] = rsqrtf(b / cols + eps); }
      else {
        const float mean = a / cols;
        stats[0] = mean;
        stats[1] = rsqrtf(fmaxf(b / cols - mean * mean, 0.f) + eps);
      }
    }
  }
  __syncthreads();
  const float mean = stats[0], rstd = stats[1];
  for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
    const uint4 xv = row[v];
    const uint4 gv = __ldg(reinterpret_cast<const uint4*>(weight) + v);
    uint4 bv = make_uint4(0, 0, 0, 0);
    if (KIND == 2 && bias != nullptr) bv = __ldg(reinterpret_cast<const uint4*>(bias) + v);
    const uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w}, gw[4] = {gv.x, gv.y, gv.z, gv.w}, bw[4] = {bv.x, bv.y, bv.z, bv.w};
    uint32_t ow[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float lo, hi;
      if (KIND == 1) {
        lo = __bfloat162float(__float2bfloat16_rn(bf16_lo(xw[j]) * rstd)) * bf16_lo(gw[j]);
        hi = __bfloat162float(__float2bfloat16_rn(bf16_hi(xw[j]) * rstd)) * bf16_hi(gw[j]);
      } else {
        lo = (bf16_lo(xw[j]) - mean) * rstd * bf16_lo(gw[j]) + bf16_lo(bw[j]);
        hi = (bf16_hi(xw[j]) - mean) * rstd * bf16_hi(gw[j]) + bf16_hi(bw[j]);
      }
      ow[j] = pack_bf16(lo, hi);
    }
    reinterpret_cast<uint4*>(out + base)[v] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
  }
}

__global__ void swiglu_kernel(const uint4* __restrict__ g, const uint4* __restrict__ u, uint4* __restrict__ o, long nvec) {
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < nvec; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const uint4 a = __ldg(g + i), b = __ldg(u + i);
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
    uint32_t ow[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float g0 = bf16_lo(aw[j]), g1 = bf16_hi(aw[j]);
      const float s0 = __bfloat162float(__float2bfloat16_rn(g0 / (1.f + __expf(-g0))));
      const float s1 = __bfloat162float(__float2bfloat16_rn(g1 / (1.f + __expf(-g1))));
      ow[j] = pack_bf16(s0 * bf16_lo(bw[j]), s1 * bf16_hi(bw[j]));
    }
    o[i] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
  }
}

// GELU (tanh or erf form) on its own: what a fused GEMV/GEMM epilogue applies, for paths that add a low-rank term before it
__global__ void gelu_kernel(const uint4* __restrict__ x, uint4* __restrict__ o, long nvec, int erf_form) {
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < nvec; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const uint4 a = __ldg(x + i);
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
    uint32_t ow[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v[2] = {bf16_lo(aw[j]), bf16_hi(aw[j])};
#pragma unroll
      for (int h = 0; h < 2; ++h)
        v[h] = erf_form ? 0.5f * v[h] * (1.f + erff(v[h] * 0.7071067811865475f))
                        : 0.5f * v[h] * (1.f + tanhf(0.7978845608028654f * (v[h] + 0.044715f * v[h] * v[h] * v[h])));
      ow[j] = pack_bf16(v[0], v[1]);
    }
    o[i] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
  }
}

__global__ void add_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, uint4* __restrict__ o, long nvec) {
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < nvec; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const uint4 x = __ldg(a + i), y = __ldg(b + i);
    uint4 r;
    r.x = pack_bf16(bf16_lo(x.x) + bf16_lo(y.x), bf16_hi(x.x) + bf16_hi(y.x));
    r.y = pack_bf16(bf16_lo(x.y) + bf16_lo(y.y), bf16_hi(x.y) + bf16_hi(y.y));
    r.z = pack_bf16(bf16_lo(x.z) + bf16_lo(y.z), bf16_hi(x.z) + bf16_hi(y.z));
    r.w = pack_bf16(bf16_lo(x.w) + bf16_lo(y.w), bf16_hi(x.w) + bf16_hi(y.w));
    o[i] = r;
  }
}

__global__ void embedding_kernel(const __nv_bfloat16* __restrict__ table, const long long* __restrict__ ids,
                                 __nv_bfloat16* __restrict__ out, int hidden) {
  const long long id = ids[blockIdx.x];
  const uint4* src = reinterpret_cast<const uint4*>(table + static_cast<size_t>(id) * hidden);
  uint4* dst = reinterpret_cast<uint4*>(out + static_cast<size_t>(blockIdx.x) * hidden);
  for (int v = threadIdx.x; v < (hidden >> 3); v += blockDim.x) dst[v] = __ldg(src + v);
}

// arg-max over a row; ties resolve to the lowest index (matches torch.argmax on CUDA for greedy decode).
template <typename T>
__global__ void __launch_bounds__(1024) argmax_kernel(const T* __restrict__ logits, long long* __restrict__ out, int vocab) {
  __shared__ float sval[32];
  __shared__ int sidx[32];
  const T* row = logits + static_cast<size_t>(blockIdx.x) * vocab;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < vocab; i += blockDim.x) {
    const float v = static_cast<float>(row[i]);
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
    if (lane == 0) out[blockIdx.x] = bi == 0x7fffffff ? 0 : bi;  // all-NaN row: never emit an out-of-range token id
  }
}

// hidden[b, t, :] += prompts[b or 0, pos + t, :] for absolute positions pos + t < P  (deep prompts
// are added to the first `P` positions of the sequence, so during decode nothing is added).
__global__ void add_prompts_kernel(__nv_bfloat16* hidden, const __nv_bfloat16* __restrict__ prompts, int T, int H,
                                   int Bp, int P, const int* pos_ptr) {
  const int b = blockIdx.y, t = blockIdx.x;
  const int pos = (pos_ptr != nullptr ? *pos_ptr : 0) + t;
  if (pos >= P) return;
  const int pb_ = Bp == 1 ? 0 : b;
  uint4* h = reinterpret_cast<uint4*>(hidden + (static_cast<size_t>(b) * T + t) * H);
  const uint4* pr = reinterpret_cast<const uint4*>(prompts + (static_cast<size_t>(pb_) * P + pos) * H);
  for (int v = threadIdx.x; v < (H >> 3); v += blockDim.x) {
    const uint4 x = h[v], y = __ldg(pr + v);
    uint4 r;
    r.x = pack_bf16(bf16_lo(x.x) + bf16_lo(y.x), bf16_hi(x.x) + bf16_hi(y.x));
    r.y = pack_bf16(bf16_lo(x.y) + bf16_lo(y.y), bf16_hi(x.y) + bf16_hi(y.y));
    r.z = pack_bf16(bf16_lo(x.z) + bf16_lo(y.z), bf16_hi(x.z) + bf16_hi(y.z));
    r.w = pack_bf16(bf16_lo(x.w) + bf16_lo(y.w), bf16_hi(x.w) + bf16_hi(y.w));
    h[v] = r;
  }
}

__global__ void bump_epoch_kernel(unsigned long long* e) { *e += 1ull; }
__global__ void advance_pos_kernel(int* p, int d) { *p += d; }

}  // namespace pb

using namespace pb;

extern "C" int pb_norm(const void* x, const void* residual, const void* weight, const void* bias, void* out,
                       void* sum_out, int rows, int cols, float eps, int kind, void* stream) {
  if (rows <= 0) return PB_OK;
  if ((cols & 7) || cols * 2 > 200 * 1024 || (kind != 1 && kind != 2)) return PB_ERR_SHAPE;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t smem = static_cast<size_t>(cols) * 2;
  auto k1 = norm_kernel<1>;
  auto k2 = norm_kernel<2>;
  if (smem > 32 * 1024) {
    cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  }
  auto X = static_cast<const __nv_bfloat16*>(x);
  auto R = static_cast<const __nv_bfloat16*>(residual);
  auto W = static_cast<const __nv_bfloat16*>(weight);
  auto Bv = static_cast<const __nv_bfloat16*>(bias);
  auto O = static_cast<__nv_bfloat16*>(out);
  auto S = static_cast<__nv_bfloat16*>(sum_out);
  if (kind == 1) k1<<<rows, 256, smem, s>>>(X, R, W, Bv, O, S, cols, eps);
  else k2<<<rows, 256, smem, s>>>(X, R, W, Bv, O, S, cols, eps);
  return pb_check_launch("elementwise");
}

static int grid_for(long nvec) {
  long g = (nvec + 255) / 256;
  if (g > 148 * 8) g = 148 * 8;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

extern "C" int pb_swiglu(const void* gate, const void* up, void* out, long n, void* stream) {
  if (n & 7) return PB_ERR_SHAPE;
  if (n == 0) return PB_OK;
  swiglu_kernel<<<grid_for(n >> 3), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(gate), static_cast<const uint4*>(up), static_cast<uint4*>(out), n >> 3);
  return pb_check_launch("elementwise");
}

extern "C" int pb_gelu(const void* x, void* out, long n, int erf_form, void* stream) {
  if (n & 7) return PB_ERR_SHAPE;
  if (n == 0) return PB_OK;
  gelu_kernel<<<grid_for(n >> 3), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint4*>(x), static_cast<uint4*>(out), n >> 3, erf_form);
  return pb_check_launch("elementwise");
}

extern "C" int pb_add(const void* a, const void* b, void* out, long n, void* stream) {
  if (n & 7) return PB_ERR_SHAPE;
  if (n == 0) return PB_OK;
  add_kernel<<<grid_for(n >> 3), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(a), static_cast<const uint4*>(b), static_cast<uint4*>(out), n >> 3);
  return pb_check_launch("elementwise");
}

extern "C" int pb_embedding(const void* table, const void* ids, void* out, int n_tokens, int hidden, void* stream) {
  if (hidden & 7) return PB_ERR_SHAPE;
  if (n_tokens == 0) return PB_OK;
  embedding_kernel<<<n_tokens, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(table), static_cast<const long long*>(ids),
      static_cast<__nv_bfloat16*>(out), hidden);
  return pb_check_launch("elementwise");
}

extern "C" int pb_argmax(const void* logits, int is_fp32, void* out_ids, int rows, int vocab, void* stream) {
  if (rows == 0) return PB_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (is_fp32) argmax_kernel<float><<<rows, 1024, 0, s>>>(static_cast<const float*>(logits), static_cast<long long*>(out_ids), vocab);
  else argmax_kernel<__nv_bfloat16><<<rows, 1024, 0, s>>>(static_cast<const __nv_bfloat16*>(logits), static_cast<long long*>(out_ids), vocab);
  return pb_check_launch("elementwise");
}

extern "C" int pb_add_prompts(void* hidden, const void* prompts, int B, int T, int H, int Bp, int P,
                              const void* pos_ptr, void* stream) {
  if ((H & 7) || (Bp != 1 && Bp != B)) return PB_ERR_SHAPE;
  if (B == 0 || T == 0 || P == 0) return PB_OK;
  add_prompts_kernel<<<dim3(T, B), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<__nv_bfloat16*>(hidden), static_cast<const __nv_bfloat16*>(prompts), T, H, Bp, P,
      static_cast<const int*>(pos_ptr));
  return pb_check_launch("elementwise");
}

extern "C" int pb_bump_epoch(void* epoch, void* stream) {
  bump_epoch_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<unsigned long long*>(epoch));
  return pb_check_launch("elementwise");
}
extern "C" int pb_advance_pos(void* pos, int delta, void* stream) {
  advance_pos_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<int*>(pos), delta);
  return pb_check_launch("elementwise");
}

extern "C" int pb_device_sm_count(int device) {
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return -1;
  return n;
}
static thread_local char g_last_error[512] = "";
extern "C" const char* pb_last_error(void) { return g_last_error; }
extern "C" int pb_check_launch(const char* what) {
  const cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) return PB_OK;
  snprintf(g_last_error, sizeof(g_last_error), "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
  return PB_ERR_CUDA;
}
extern "C" int pb_set_error(const char* msg) { snprintf(g_last_error, sizeof(g_last_error), "%s", msg); return 0; }
extern "C" int pb_version(void) { return 1; }
