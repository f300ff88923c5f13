// petals_b200 — decode-shape linear layer over block-scaled FP8 weights (MXFP8: E4M3 payload, one UE8M0
// power-of-two scale per 32 values along K; format defined in petals_b200/ops/quant.py).
//
//   out[m, n] = epilogue( sum_k prologue(x)[m, k] * (q[n, k] * 2^(e[n, k/32] - 127)) )
//
// This is the B200 replacement of the reference's bitsandbytes INT8 / NF4 linears on the single-stream path
// (src/petals/utils/convert_block.py:76-115): at decode time the op is bound by weight
// bytes, so 1-byte weights halve the time per token. Same structure as linear_decode.cu (warp per output-column
// pair, 8 x 16-byte non-allocating loads in flight per lane, fused norm prologue and bias / SwiGLU / GELU / residual
// epilogue); the 16 FP8 values of a load share one scale, so the inner product of a load is accumulated
// unscaled and scaled once. x is staged in shared memory as fp16, pre-scaled per token by a power of two so that |x| <= 8:
// the 16 products of a load are then accumulated with packed HFMA2 (e4m3 -> f16x2 is one cvt; |sum| <= 8*448*8 < 65504 cannot
// overflow) and only the per-load partial goes to fp32 - half the instructions and half the shared-memory traffic of an fp32 path.
#include "common.cuh"
#include "petals_b200.h"

#include <cuda_fp16.h>

namespace pb {

struct LinearFp8Params {
  const __nv_bfloat16* x;
  const uint8_t* w;        // [N, K] e4m3
  const uint8_t* ws;       // [N, K/32] ue8m0
  const uint8_t* w2;       // SwiGLU up_proj payload / scales
  const uint8_t* ws2;
  const __nv_bfloat16* bias;
  const __nv_bfloat16* bias2;
  const __nv_bfloat16* residual;
  __nv_bfloat16* out;
  const __nv_bfloat16* norm_w;
  const __nv_bfloat16* norm_b;
  float eps;
  int norm_kind, act, M, N, K;
};

PB_DEVICE float2 e4m3x2_to_float2(uint16_t v) {
  uint32_t h2;
  asm("cvt.rn.f16x2.e4m3x2 %0, %1;" : "=r"(h2) : "h"(v));
  return __half22float2(*reinterpret_cast<__half2*>(&h2));
}
PB_DEVICE float ue8m0_to_float(uint8_t e) { return __uint_as_float(static_cast<uint32_t>(e) << 23); }  // 2^(e-127)
PB_DEVICE float silu8(float x) { return x / (1.f + __expf(-x)); }
PB_DEVICE float gelu_tanh8(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  return 0.5f * x * (1.f + tanhf(k0 * (x + k1 * x * x * x)));
}
PB_DEVICE float rb(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// x lives in shared memory as fp16; a lane reads two consecutive 16-byte chunks per load (32-byte lane stride), which would be a
// 2-way bank conflict. Chunk c of a row is therefore stored at c ^ ((c >> 3) & 1): conflict-free for this access pattern.
PB_DEVICE int swz_chunk(int c) { return c ^ ((c >> 3) & 1); }

PB_DEVICE __half2 e4m3x2_to_half2(uint16_t v) {
  uint32_t h2;
  asm("cvt.rn.f16x2.e4m3x2 %0, %1;" : "=r"(h2) : "h"(v));
  return *reinterpret_cast<__half2*>(&h2);
}
PB_DEVICE __half2 as_half2(uint32_t v) { return *reinterpret_cast<__half2*>(&v); }

// 16 e4m3 values (one uint4) times 16 fp16 activations per token: 8 cvt + 8 HFMA2 per token
template <int M>
PB_DEVICE void dot16(float (&acc)[M], const uint4& w, float scale, const __half* xs, int K, int k) {
  const uint32_t words[4] = {w.x, w.y, w.z, w.w};
  __half2 wh[8];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    wh[2 * i] = e4m3x2_to_half2(static_cast<uint16_t>(words[i] & 0xffffu));
    wh[2 * i + 1] = e4m3x2_to_half2(static_cast<uint16_t>(words[i] >> 16));
  }
  const int c = k >> 3;  // 16-byte chunk (8 halves) index within the row; k is a multiple of 16
#pragma unroll
  for (int m = 0; m < M; ++m) {
    const uint4* xrow = reinterpret_cast<const uint4*>(xs + static_cast<size_t>(m) * K);
    const uint4 xa = xrow[swz_chunk(c)], xb = xrow[swz_chunk(c + 1)];
    __half2 s0 = __hmul2(wh[0], as_half2(xa.x)), s1 = __hmul2(wh[4], as_half2(xb.x));
    s0 = __hfma2(wh[1], as_half2(xa.y), s0); s1 = __hfma2(wh[5], as_half2(xb.y), s1);
    s0 = __hfma2(wh[2], as_half2(xa.z), s0); s1 = __hfma2(wh[6], as_half2(xb.z), s1);
    s0 = __hfma2(wh[3], as_half2(xa.w), s0); s1 = __hfma2(wh[7], as_half2(xb.w), s1);
    const float2 f = __half22float2(__hadd2(s0, s1));
    acc[m] = fmaf(f.x + f.y, scale, acc[m]);
  }
}

PB_DEVICE float block_sum8(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = (lane < nw) ? red[lane] : 0.f;
  return warp_sum(t);
}

PB_DEVICE float block_max8(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = (lane < nw) ? red[lane] : 0.f;
  return warp_max(t);
}

template <int M, bool DUAL>
__global__ void __launch_bounds__(768, 1) linear_decode_fp8_kernel(const LinearFp8Params p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __half* xs = reinterpret_cast<__half*>(smem_raw);  // [M, K] fp16, row m scaled by 2^-xexp[m]
  __shared__ float red[32];
  __shared__ float stat[2 * M];
  __shared__ float xinv[M];  // 2^xexp[m]: undoes the activation scaling in the epilogue
  const int K = p.K, N = p.N;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5, tid = threadIdx.x, nthr = blockDim.x;

  // ---- prologue: x (+ norm) -> scaled fp16 in shared memory ----------------------------------------------------
  {
    if (p.norm_kind != 0) {
      float ssum[M], ssq[M];
#pragma unroll
      for (int m = 0; m < M; ++m) ssum[m] = ssq[m] = 0.f;
      const int kvec = K >> 3;
#pragma unroll
      for (int m = 0; m < M; ++m) {
        for (int v = tid; v < kvec; v += nthr) {
          const uint4 xv = __ldcg(reinterpret_cast<const uint4*>(p.x + static_cast<size_t>(m) * K) + v);
          const float f[8] = {bf16_lo(xv.x), bf16_hi(xv.x), bf16_lo(xv.y), bf16_hi(xv.y), bf16_lo(xv.z), bf16_hi(xv.z), bf16_lo(xv.w), bf16_hi(xv.w)};
#pragma unroll
          for (int i = 0; i < 8; ++i) { ssum[m] += f[i]; ssq[m] += f[i] * f[i]; }
        }
      }
#pragma unroll
      for (int m = 0; m < M; ++m) {
        const float s1 = block_sum8(ssum[m], red), s2 = block_sum8(ssq[m], red);
        if (tid == 0) {
          if (p.norm_kind == 1) { stat[2 * m] = 0.f; stat[2 * m + 1] = rsqrtf(s2 / K + p.eps); }
          else {
            const float mean = s1 / K;
            stat[2 * m] = mean;
            stat[2 * m + 1] = rsqrtf(fmaxf(s2 / K - mean * mean, 0.f) + p.eps);
          }
        }
      }
      __syncthreads();
    }
    // normalised activations (bf16-rounded like the dense path) -> fp16 in smem, and the per-token maximum
    float amax[M];
#pragma unroll
    for (int m = 0; m < M; ++m) amax[m] = 0.f;
    {
      const int kvec = K >> 3;
#pragma unroll
      for (int m = 0; m < M; ++m) {
        for (int v8 = tid; v8 < kvec; v8 += nthr) {
          const uint4 xv = __ldcg(reinterpret_cast<const uint4*>(p.x + static_cast<size_t>(m) * K) + v8);  // second read hits L2
          float f[8] = {bf16_lo(xv.x), bf16_hi(xv.x), bf16_lo(xv.y), bf16_hi(xv.y), bf16_lo(xv.z), bf16_hi(xv.z), bf16_lo(xv.w), bf16_hi(xv.w)};
          if (p.norm_kind != 0) {
            const uint4 gv = __ldg(reinterpret_cast<const uint4*>(p.norm_w) + v8);
            const float g[8] = {bf16_lo(gv.x), bf16_hi(gv.x), bf16_lo(gv.y), bf16_hi(gv.y), bf16_lo(gv.z), bf16_hi(gv.z), bf16_lo(gv.w), bf16_hi(gv.w)};
            float bb[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (p.norm_kind == 2 && p.norm_b != nullptr) {
              const uint4 bv = __ldg(reinterpret_cast<const uint4*>(p.norm_b) + v8);
              bb[0] = bf16_lo(bv.x); bb[1] = bf16_hi(bv.x); bb[2] = bf16_lo(bv.y); bb[3] = bf16_hi(bv.y);
              bb[4] = bf16_lo(bv.z); bb[5] = bf16_hi(bv.z); bb[6] = bf16_lo(bv.w); bb[7] = bf16_hi(bv.w);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
              f[i] = p.norm_kind == 1 ? rb(rb(f[i] * stat[2 * m + 1]) * g[i]) : rb((f[i] - stat[2 * m]) * stat[2 * m + 1] * g[i] + bb[i]);
          }
          __half2 h[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float a = fminf(fmaxf(f[2 * i], -60000.f), 60000.f), b = fminf(fmaxf(f[2 * i + 1], -60000.f), 60000.f);  // fp16 range
            amax[m] = fmaxf(amax[m], fmaxf(fabsf(a), fabsf(b)));
            h[i] = __floats2half2_rn(a, b);
          }
          *reinterpret_cast<uint4*>(xs + static_cast<size_t>(m) * K + swz_chunk(v8) * 8) = *reinterpret_cast<const uint4*>(h);
        }
      }
    }
#pragma unroll
    for (int m = 0; m < M; ++m) {
      const float mx = block_max8(amax[m], red);
      if (tid == 0) {
        // power of two that brings the row maximum into (4, 8]; exact in fp16 for everything that matters
        int e = 0;
        if (mx > 0.f) { frexpf(mx, &e); e -= 3; }   // mx = f * 2^e', f in [0.5, 1): mx * 2^-(e'-3) in [4, 8)
        stat[2 * m] = ldexpf(1.f, -e);              // reuse: multiplier applied to the staged row
        xinv[m] = ldexpf(1.f, e);
      }
    }
    __syncthreads();
    for (int k2 = tid; k2 < (M * K) >> 1; k2 += nthr) {
      const int m = (k2 << 1) / K;
      const float sc = stat[2 * m];
      if (sc != 1.f) {
        __half2* px = reinterpret_cast<__half2*>(xs) + k2;
        *px = __hmul2(*px, __float2half2_rn(sc));
      }
    }
    __syncthreads();
  }

  // ---- main loop ---------------------------------------------------------------------------------------------
  constexpr int U = DUAL ? 2 : 4;
  const int ntasks = N >> 1;
  const int total_warps = gridDim.x * nwarps;
  const int kstep = 512 * U;  // 32 lanes x 16 values x U
  const int KS = K >> 5;      // scales per row
  for (int task = warp * gridDim.x + blockIdx.x; task < ntasks; task += total_warps) {
    const int n0 = task << 1;
    const uint8_t* w0 = p.w + static_cast<size_t>(n0) * K;
    const uint8_t* w1 = w0 + K;
    const uint8_t* s0 = p.ws + static_cast<size_t>(n0) * KS;
    const uint8_t* s1 = s0 + KS;
    const uint8_t* u0 = DUAL ? p.w2 + static_cast<size_t>(n0) * K : nullptr;
    const uint8_t* u1 = DUAL ? u0 + K : nullptr;
    const uint8_t* t0 = DUAL ? p.ws2 + static_cast<size_t>(n0) * KS : nullptr;
    const uint8_t* t1 = DUAL ? t0 + KS : nullptr;
    float a0[M], a1[M], b0[M], b1[M];
#pragma unroll
    for (int m = 0; m < M; ++m) a0[m] = a1[m] = b0[m] = b1[m] = 0.f;
    for (int kb = 0; kb < K; kb += kstep) {
      uint4 wa[U], wb[U], ua[U], ub[U];
      // raw scale bytes: converting them here would make the (in-order) warp wait for every scale load before it can issue the
      // next group of weight loads - four serialised memory latencies per iteration (measured: the fp8 kernel was no faster than
      // the bf16 one on half the bytes). The exponent -> float conversion happens at the point of use instead.
      uint8_t sa[U], sb[U], ta[U], tb[U];
      bool ok[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int k = kb + u * 512 + lane * 16;
        ok[u] = k < K;
        if (ok[u]) {
          wa[u] = ld_stream(w0 + k);
          wb[u] = ld_stream(w1 + k);
          if (DUAL) {
            ua[u] = ld_stream(u0 + k);
            ub[u] = ld_stream(u1 + k);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int k = kb + u * 512 + lane * 16;
        if (ok[u]) {
          sa[u] = __ldg(s0 + (k >> 5));
          sb[u] = __ldg(s1 + (k >> 5));
          if (DUAL) {
            ta[u] = __ldg(t0 + (k >> 5));
            tb[u] = __ldg(t1 + (k >> 5));
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (ok[u]) {
          const int k = kb + u * 512 + lane * 16;
          dot16<M>(a0, wa[u], ue8m0_to_float(sa[u]), xs, K, k);
          dot16<M>(a1, wb[u], ue8m0_to_float(sb[u]), xs, K, k);
          if (DUAL) {
            dot16<M>(b0, ua[u], ue8m0_to_float(ta[u]), xs, K, k);
            dot16<M>(b1, ub[u], ue8m0_to_float(tb[u]), xs, K, k);
          }
        }
      }
    }
#pragma unroll
    for (int m = 0; m < M; ++m) {
      a0[m] = warp_sum(a0[m]); a1[m] = warp_sum(a1[m]);
      if (DUAL) { b0[m] = warp_sum(b0[m]); b1[m] = warp_sum(b1[m]); }
    }
    float v0 = 0.f, v1 = 0.f, c0 = 0.f, c1 = 0.f;
#pragma unroll
    for (int m = 0; m < M; ++m)
      if (lane == m) { v0 = a0[m]; v1 = a1[m]; if (DUAL) { c0 = b0[m]; c1 = b1[m]; } }
    if (lane < M) {
      const float xi = xinv[lane];  // undo the fp16 activation scaling
      v0 *= xi; v1 *= xi; c0 *= xi; c1 *= xi;
      if (p.bias != nullptr) { v0 += __bfloat162float(p.bias[n0]); v1 += __bfloat162float(p.bias[n0 + 1]); }
      if (DUAL) {
        if (p.bias2 != nullptr) { c0 += __bfloat162float(p.bias2[n0]); c1 += __bfloat162float(p.bias2[n0 + 1]); }
        v0 = rb(silu8(rb(v0))) * rb(c0);
        v1 = rb(silu8(rb(v1))) * rb(c1);
      } else if (p.act == 2) {
        v0 = gelu_tanh8(v0); v1 = gelu_tanh8(v1);
      } else if (p.act == 3) {
        v0 = 0.5f * v0 * (1.f + erff(v0 * 0.7071067811865475f));
        v1 = 0.5f * v1 * (1.f + erff(v1 * 0.7071067811865475f));
      }
      const size_t o = static_cast<size_t>(lane) * N + n0;
      if (p.residual != nullptr) {
        const uint32_t rv = *reinterpret_cast<const uint32_t*>(p.residual + o);
        v0 = rb(v0) + bf16_lo(rv);
        v1 = rb(v1) + bf16_hi(rv);
      }
      *reinterpret_cast<uint32_t*>(p.out + o) = pack_bf16(v0, v1);
    }
  }
}

// out[n, k] = q[n, k] * 2^(e[n, k/32] - 127) as bf16 — materialises one projection for the tcgen05 GEMM (prefill): the extra
// pass over the weights is ~1/M of the GEMM's work for M prompt tokens.
__global__ void __launch_bounds__(256) dequant_mxfp8_kernel(const uint4* __restrict__ q, const uint8_t* __restrict__ e, uint4* __restrict__ out, long n_vec) {
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n_vec; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const uint4 w = ld_stream(q + i);                      // 16 values; 32-value scale blocks are 2 vectors wide
    const float scale = ue8m0_to_float(__ldg(e + (i >> 1)));
    const uint32_t words[4] = {w.x, w.y, w.z, w.w};
    uint32_t o[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 a = e4m3x2_to_float2(static_cast<uint16_t>(words[j] & 0xffffu));
      const float2 b = e4m3x2_to_float2(static_cast<uint16_t>(words[j] >> 16));
      o[2 * j] = pack_bf16(a.x * scale, a.y * scale);
      o[2 * j + 1] = pack_bf16(b.x * scale, b.y * scale);
    }
    out[2 * i] = make_uint4(o[0], o[1], o[2], o[3]);
    out[2 * i + 1] = make_uint4(o[4], o[5], o[6], o[7]);
  }
}

template <int M>
static int launch_fp8(const LinearFp8Params& p, bool dual, int grid, int block, size_t smem, cudaStream_t s) {
  if (dual) {
    auto k = linear_decode_fp8_kernel<M, true>;
    if (smem > 32 * 1024 && cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess) return PB_ERR_CUDA;
    k<<<grid, block, smem, s>>>(p);
  } else {
    auto k = linear_decode_fp8_kernel<M, false>;
    if (smem > 32 * 1024 && cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess) return PB_ERR_CUDA;
    k<<<grid, block, smem, s>>>(p);
  }
  return pb_check_launch("linear_decode_fp8");
}

}  // namespace pb

extern "C" int pb_linear_decode_fp8(const PbLinearFp8Args* a, void* stream) {
  using namespace pb;
  if (a->M < 1 || a->M > 4 || (a->N & 1) || (a->K & 31)) return PB_ERR_SHAPE;
  const size_t smem = static_cast<size_t>(a->M) * a->K * 2;  // fp16 staging
  if (smem > 200 * 1024) return PB_ERR_SHAPE;
  LinearFp8Params p{};
  p.x = static_cast<const __nv_bfloat16*>(a->x);
  p.w = static_cast<const uint8_t*>(a->w); p.ws = static_cast<const uint8_t*>(a->w_scale);
  p.w2 = static_cast<const uint8_t*>(a->w2); p.ws2 = static_cast<const uint8_t*>(a->w2_scale);
  p.bias = static_cast<const __nv_bfloat16*>(a->bias); p.bias2 = static_cast<const __nv_bfloat16*>(a->bias2);
  p.residual = static_cast<const __nv_bfloat16*>(a->residual);
  p.out = static_cast<__nv_bfloat16*>(a->out);
  p.norm_w = static_cast<const __nv_bfloat16*>(a->norm_w); p.norm_b = static_cast<const __nv_bfloat16*>(a->norm_b);
  p.eps = a->eps; p.norm_kind = a->norm_kind; p.act = a->act; p.M = a->M; p.N = a->N; p.K = a->K;
  const bool dual = a->act == 1;
  if (dual && (p.w2 == nullptr || p.ws2 == nullptr)) return PB_ERR_SHAPE;
  const int sms = a->num_sms > 0 ? a->num_sms : 148;
  const int ntasks = a->N / 2;
  int best_w = 24;
  double best_eff = -1.0;
  for (int w = 24; w >= 12; --w) {
    const long tw = static_cast<long>(sms) * w;
    const long rounds = (ntasks + tw - 1) / tw;
    const double eff = static_cast<double>(ntasks) / static_cast<double>(rounds * tw);
    if (eff > best_eff + 1e-9) { best_eff = eff; best_w = w; }
  }
  int grid = sms;
  if (ntasks < sms * best_w) grid = max(1, min(sms, (ntasks + best_w - 1) / best_w));
  const int block = best_w * 32;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (a->M) {
    case 1: return launch_fp8<1>(p, dual, grid, block, smem, s);
    case 2: return launch_fp8<2>(p, dual, grid, block, smem, s);
    case 3: return launch_fp8<3>(p, dual, grid, block, smem, s);
    default: return launch_fp8<4>(p, dual, grid, block, smem, s);
  }
}

extern "C" int pb_dequant_mxfp8(const void* q, const void* e, void* out, long n_elems, void* stream) {
  using namespace pb;
  if (n_elems & 31) return PB_ERR_SHAPE;
  if (n_elems == 0) return PB_OK;
  const long n_vec = n_elems >> 4;
  long grid = (n_vec + 255) / 256;
  if (grid > 148 * 16) grid = 148 * 16;
  dequant_mxfp8_kernel<<<static_cast<int>(grid), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint4*>(q), static_cast<const uint8_t*>(e),
                                                                                              static_cast<uint4*>(out), n_vec);
  return pb_check_launch("dequant_mxfp8");
}
