// petals_b200 — decode-shape linear layer for 2..8 rows on the tensor cores ("skinny GEMM").
//
// linear_decode.cu multiplies every streamed weight by M activations with scalar FMAs: fine for one token, but at M = 4 the
// FFMA issue rate caps the kernel at 53-63 % of HBM bandwidth (profiles/r1_kernel_bench_v1.txt). Here the M <= 8 tokens are the
// N = 8 dimension of mma.sync.m16n8k16 (bf16 inputs, fp32 accumulate): one instruction performs 16 weight rows x 8 tokens x
// 16 k, so the math disappears behind the weight stream for any M <= 8 (batched sessions, beam search, speculative verification).
//
//   work unit  = (block of 16 weight rows, slice of K); units are dealt to warps round-robin (persistent grid, 16 warps per SM)
//   A operand  = weights straight from global memory into registers: lane (r, c) = (lane / 4, lane % 4) loads the 32 bytes
//                k0 + 16c .. +15 of rows r and r + 8. The fragment layout of mma wants k = 2c, 2c+1, 2c+8, 2c+9 per lane
//                instead, but k is only summed over: the SAME permutation of k is applied to the B operand (x from shared
//                memory), so no shuffle or shared-memory staging of the weights is needed and every row is read in full 128-byte lines
//   B operand  = x[token = lane / 4][k0 + 16c + 4s .. +3]: one 8-byte shared-memory load per mma (rows padded by 8 bytes:
//                conflict-free for this pattern)
//   reduction  = the K slices of a row block add their 16 x 8 fp32 tiles into a global scratch with red.add; the last slice to
//                arrive (per-block counter) applies bias / SwiGLU / GELU / residual, writes bf16 and re-zeroes the scratch
//
// Prologue (x -> optional RMSNorm / LayerNorm -> bf16 in shared memory) and epilogue rounding follow linear_decode.cu.
#include "common.cuh"
#include "petals_b200.h"

extern "C" int pb_set_error(const char* msg);  // elementwise.cu

namespace pb {

struct SkinnyParams {
  const __nv_bfloat16* x;         // [M, K]
  const __nv_bfloat16* w;         // [N, K]
  const __nv_bfloat16* w2;        // [N, K] (SwiGLU up projection) or null
  const __nv_bfloat16* bias;
  const __nv_bfloat16* bias2;
  const __nv_bfloat16* residual;  // [M, N] or null
  __nv_bfloat16* out;             // [M, N]
  const __nv_bfloat16* norm_w;
  const __nv_bfloat16* norm_b;
  float* scratch;                 // [2, N, 8] fp32, zero between launches (self-cleaning)
  unsigned int* counters;         // [N / 16], zero between launches (self-cleaning)
  float eps;
  int norm_kind, act, M, N, K, ks, slice;  // ks slices of `slice` k each per row block
};

constexpr int kSkThreads = 512;

PB_DEVICE void mma_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
PB_DEVICE float sk_rb(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
PB_DEVICE float sk_silu(float x) { return x / (1.f + __expf(-x)); }
PB_DEVICE float sk_gelu_tanh(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  return 0.5f * x * (1.f + tanhf(k0 * (x + k1 * x * x * x)));
}
PB_DEVICE float sk_block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  return warp_sum(lane < nw ? red[lane] : 0.f);
}

// DUAL: SwiGLU (two weight matrices). XS: x staged (and optionally normalised) in shared memory; otherwise the B fragments are
// read from global memory through L1 (very wide K, e.g. the down projection: 8 rows of x do not fit next to nothing else).
template <bool DUAL, bool XS>
__global__ void __launch_bounds__(kSkThreads, 1) linear_decode_mma_kernel(const SkinnyParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __shared__ float red[32];
  const int K = p.K, N = p.N, M = p.M;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthr = blockDim.x;
  const int row_bytes = K * 2 + 8;  // 8-byte pad per token row: the B-fragment loads below become conflict-free
  pdl_trigger();
  pdl_wait();

  // ---- prologue: x (+ norm) -> bf16 rows in shared memory, rows M..7 zero ------------------------------------------
  for (int m = 0; XS && m < 8; ++m) {
    float mean = 0.f, rstd = 1.f;
    if (m < M && p.norm_kind != 0) {
      float s1 = 0.f, s2 = 0.f;
      for (int k = tid * 2; k < K; k += nthr * 2) {
        const uint32_t v = *reinterpret_cast<const uint32_t*>(p.x + static_cast<size_t>(m) * K + k);
        const float a = bf16_lo(v), b = bf16_hi(v);
        s1 += a + b; s2 += a * a + b * b;
      }
      s1 = sk_block_sum(s1, red);
      s2 = sk_block_sum(s2, red);
      if (p.norm_kind == 1) rstd = rsqrtf(s2 / K + p.eps);
      else { mean = s1 / K; rstd = rsqrtf(fmaxf(s2 / K - mean * mean, 0.f) + p.eps); }
    }
    for (int k = tid * 2; k < K; k += nthr * 2) {
      uint32_t o = 0u;
      if (m < M) {
        const uint32_t v = *reinterpret_cast<const uint32_t*>(p.x + static_cast<size_t>(m) * K + k);
        float a = bf16_lo(v), b = bf16_hi(v);
        if (p.norm_kind != 0) {
          const uint32_t g = *reinterpret_cast<const uint32_t*>(p.norm_w + k);
          if (p.norm_kind == 1) {  // HF RMSNorm rounding: weight * bf16(x * rstd)
            a = sk_rb(a * rstd) * bf16_lo(g);
            b = sk_rb(b * rstd) * bf16_hi(g);
          } else {
            const uint32_t bb = p.norm_b != nullptr ? *reinterpret_cast<const uint32_t*>(p.norm_b + k) : 0u;
            a = (a - mean) * rstd * bf16_lo(g) + bf16_lo(bb);
            b = (b - mean) * rstd * bf16_hi(g) + bf16_hi(bb);
          }
        }
        o = pack_bf16(a, b);
      }
      *reinterpret_cast<uint32_t*>(smem_raw + static_cast<size_t>(m) * row_bytes + k * 2) = o;
    }
  }
  __syncthreads();

  // ---- main loop over (row block, K slice) units -----------------------------------------------------------------------
  const int r = lane >> 2, c = lane & 3;
  const int n_rb = N >> 4;
  const int n_units = n_rb * p.ks;
  const int total_warps = gridDim.x * (nthr >> 5);
  const uint8_t* xrow = XS ? smem_raw + static_cast<size_t>(r) * row_bytes  // token r = lane / 4 is this lane's B column
                           : reinterpret_cast<const uint8_t*>(p.x + static_cast<size_t>(r < M ? r : 0) * K);
  for (int u = warp * gridDim.x + blockIdx.x; u < n_units; u += total_warps) {
    const int rb = u / p.ks, ks = u - rb * p.ks;
    const int k_begin = ks * p.slice, k_end = min(K, k_begin + p.slice);
    const __nv_bfloat16* w_lo = p.w + static_cast<size_t>(rb * 16 + r) * K;
    const __nv_bfloat16* w_hi = w_lo + static_cast<size_t>(8) * K;
    const __nv_bfloat16* u_lo = DUAL ? p.w2 + static_cast<size_t>(rb * 16 + r) * K : nullptr;
    const __nv_bfloat16* u_hi = DUAL ? u_lo + static_cast<size_t>(8) * K : nullptr;
    float acc[4] = {0.f, 0.f, 0.f, 0.f}, acc2[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k0 = k_begin; k0 < k_end; k0 += 128) {  // two 64-k steps in flight: 128 B per lane per matrix
      uint4 a_lo[2][2], a_hi[2][2], b_lo[2][2], b_hi[2][2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int k = k0 + h * 64 + c * 16;
        if (k0 + h * 64 < k_end) {
          a_lo[h][0] = ld_stream(w_lo + k); a_lo[h][1] = ld_stream(w_lo + k + 8);
          a_hi[h][0] = ld_stream(w_hi + k); a_hi[h][1] = ld_stream(w_hi + k + 8);
          if (DUAL) {
            b_lo[h][0] = ld_stream(u_lo + k); b_lo[h][1] = ld_stream(u_lo + k + 8);
            b_hi[h][0] = ld_stream(u_hi + k); b_hi[h][1] = ld_stream(u_hi + k + 8);
          }
        }
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (k0 + h * 64 < k_end) {
          const int k = k0 + h * 64 + c * 16;
          const uint32_t wl[8] = {a_lo[h][0].x, a_lo[h][0].y, a_lo[h][0].z, a_lo[h][0].w, a_lo[h][1].x, a_lo[h][1].y, a_lo[h][1].z, a_lo[h][1].w};
          const uint32_t wh[8] = {a_hi[h][0].x, a_hi[h][0].y, a_hi[h][0].z, a_hi[h][0].w, a_hi[h][1].x, a_hi[h][1].y, a_hi[h][1].z, a_hi[h][1].w};
          uint32_t ul[8] = {0, 0, 0, 0, 0, 0, 0, 0}, uh[8] = {0, 0, 0, 0, 0, 0, 0, 0};
          if (DUAL) {
            const uint32_t t1[8] = {b_lo[h][0].x, b_lo[h][0].y, b_lo[h][0].z, b_lo[h][0].w, b_lo[h][1].x, b_lo[h][1].y, b_lo[h][1].z, b_lo[h][1].w};
            const uint32_t t2[8] = {b_hi[h][0].x, b_hi[h][0].y, b_hi[h][0].z, b_hi[h][0].w, b_hi[h][1].x, b_hi[h][1].y, b_hi[h][1].z, b_hi[h][1].w};
#pragma unroll
            for (int i = 0; i < 8; ++i) { ul[i] = t1[i]; uh[i] = t2[i]; }
          }
#pragma unroll
          for (int s = 0; s < 4; ++s) {
            // virtual k (2c, 2c+1 | 2c+8, 2c+9) of mma step s  <->  physical k0 + 16c + 4s + (0,1 | 2,3), for A and B alike
            uint2 xb = *reinterpret_cast<const uint2*>(xrow + (k + 4 * s) * 2);
            if (!XS && r >= M) xb = make_uint2(0u, 0u);
            mma_16816(acc, wl[2 * s], wh[2 * s], wl[2 * s + 1], wh[2 * s + 1], xb.x, xb.y);
            if (DUAL) mma_16816(acc2, ul[2 * s], uh[2 * s], ul[2 * s + 1], uh[2 * s + 1], xb.x, xb.y);
          }
        }
      }
    }
    // ---- cross-slice reduction: C fragment (row r / r+8, tokens 2c / 2c+1) -> scratch[matrix][row][token] ---------------
    float* sc = p.scratch + static_cast<size_t>(rb) * 128;
    atomicAdd(sc + r * 8 + 2 * c, acc[0]);
    atomicAdd(sc + r * 8 + 2 * c + 1, acc[1]);
    atomicAdd(sc + (r + 8) * 8 + 2 * c, acc[2]);
    atomicAdd(sc + (r + 8) * 8 + 2 * c + 1, acc[3]);
    float* sc2 = sc + static_cast<size_t>(N) * 8;
    if (DUAL) {
      atomicAdd(sc2 + r * 8 + 2 * c, acc2[0]);
      atomicAdd(sc2 + r * 8 + 2 * c + 1, acc2[1]);
      atomicAdd(sc2 + (r + 8) * 8 + 2 * c, acc2[2]);
      atomicAdd(sc2 + (r + 8) * 8 + 2 * c + 1, acc2[3]);
    }
    __threadfence();
    __syncwarp();
    unsigned int prev = 0;
    if (lane == 0) prev = atomicAdd(p.counters + rb, 1u);
    prev = __shfl_sync(0xffffffffu, prev, 0);
    if (prev == static_cast<unsigned int>(p.ks) - 1u) {
      // ---- last slice of this row block: epilogue. lane -> row rb*16 + lane % 16, tokens (lane / 16) * 4 .. +3 --------------
      __threadfence();
      const int row = lane & 15, t0 = (lane >> 4) * 4;
      const int n = rb * 16 + row;
      float4* s4 = reinterpret_cast<float4*>(sc + row * 8 + t0);
      const float4 v4 = __ldcg(s4);
      *s4 = make_float4(0.f, 0.f, 0.f, 0.f);
      float v[4] = {v4.x, v4.y, v4.z, v4.w};
      float g2[4] = {0.f, 0.f, 0.f, 0.f};
      if (DUAL) {
        float4* u4 = reinterpret_cast<float4*>(sc2 + row * 8 + t0);
        const float4 w4 = __ldcg(u4);
        *u4 = make_float4(0.f, 0.f, 0.f, 0.f);
        g2[0] = w4.x; g2[1] = w4.y; g2[2] = w4.z; g2[3] = w4.w;
      }
      const float bv = p.bias != nullptr ? __bfloat162float(p.bias[n]) : 0.f;
      const float bv2 = (DUAL && p.bias2 != nullptr) ? __bfloat162float(p.bias2[n]) : 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int t = t0 + i;
        if (t < M) {
          float y = v[i] + bv;
          if (DUAL) {
            const float up = g2[i] + bv2;
            y = sk_rb(sk_silu(sk_rb(y))) * sk_rb(up);  // HF: act(gate) * up, each projection rounded to bf16 first
          } else if (p.act == 2) {
            y = sk_gelu_tanh(y);
          } else if (p.act == 3) {
            y = 0.5f * y * (1.f + erff(y * 0.7071067811865475f));
          }
          const size_t o = static_cast<size_t>(t) * N + n;
          if (p.residual != nullptr) y = sk_rb(y) + __bfloat162float(p.residual[o]);
          p.out[o] = __float2bfloat16_rn(y);
        }
      }
      __syncwarp();
      if (lane == 0) p.counters[rb] = 0u;
    }
  }
}

}  // namespace pb

// scratch: fp32 [2 * N * 8], counters: uint32 [N / 16]; both zero-initialised once by the caller and left zero by every launch.
extern "C" int pb_linear_decode_mma(const PbLinearDecodeArgs* a, void* scratch, void* counters, void* stream) {
  using namespace pb;
  if (a->M < 1 || a->M > 8 || (a->N & 15) || (a->K & 127) || scratch == nullptr || counters == nullptr) return PB_ERR_SHAPE;
  if (a->n_parts > 0 || a->n_push > 0 || a->n_ll_parts > 0 || a->n_ll_push > 0 || a->wait_flag != nullptr || a->rope_q_out != nullptr ||
      a->x_out != nullptr || a->out == nullptr)
    return PB_ERR_SHAPE;
  const bool dual = a->act == 1;
  if (dual && a->w2 == nullptr) return PB_ERR_SHAPE;
  if (a->norm_kind != 0 && a->norm_w == nullptr) return PB_ERR_SHAPE;
  SkinnyParams p{};
  p.x = static_cast<const __nv_bfloat16*>(a->x);
  p.w = static_cast<const __nv_bfloat16*>(a->w);
  p.w2 = static_cast<const __nv_bfloat16*>(a->w2);
  p.bias = static_cast<const __nv_bfloat16*>(a->bias);
  p.bias2 = static_cast<const __nv_bfloat16*>(a->bias2);
  p.residual = static_cast<const __nv_bfloat16*>(a->residual);
  p.out = static_cast<__nv_bfloat16*>(a->out);
  p.norm_w = static_cast<const __nv_bfloat16*>(a->norm_w);
  p.norm_b = static_cast<const __nv_bfloat16*>(a->norm_b);
  p.scratch = static_cast<float*>(scratch);
  p.counters = static_cast<unsigned int*>(counters);
  p.eps = a->eps; p.norm_kind = a->norm_kind; p.act = a->act; p.M = a->M; p.N = a->N; p.K = a->K;
  const int sms = a->num_sms > 0 ? a->num_sms : 148;
  const int warps = sms * (kSkThreads / 32);
  const int n_rb = a->N / 16;
  // slices: about two units per warp, slice a multiple of 128 k, at least 512 k long
  int ks = (2 * warps + n_rb - 1) / n_rb;
  const int max_ks = a->K / 512 > 0 ? a->K / 512 : 1;
  if (ks > max_ks) ks = max_ks;
  if (ks < 1) ks = 1;
  int slice = (a->K + ks - 1) / ks;
  slice = (slice + 127) / 128 * 128;
  ks = (a->K + slice - 1) / slice;
  p.ks = ks; p.slice = slice;
  size_t smem = static_cast<size_t>(8) * (static_cast<size_t>(a->K) * 2 + 8);
  const bool xs = smem <= 160 * 1024;
  if (!xs) {
    if (a->norm_kind != 0) return PB_ERR_SHAPE;  // a fused norm needs the staged copy
    smem = 0;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  auto launch = [&](auto kern) {
    if (smem > 32 * 1024) {
      const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
      if (e != cudaSuccess) return e;
    }
    return launch_pdl(kPdlGemv, kern, dim3(sms), dim3(kSkThreads), smem, s, p);
  };
  const cudaError_t e = dual ? (xs ? launch(linear_decode_mma_kernel<true, true>) : launch(linear_decode_mma_kernel<true, false>))
                             : (xs ? launch(linear_decode_mma_kernel<false, true>) : launch(linear_decode_mma_kernel<false, false>));
  if (e != cudaSuccess) {
    cudaGetLastError();
    char what[160];
    snprintf(what, sizeof(what), "linear_decode_mma M=%d N=%d K=%d ks=%d smem=%zu: %s", a->M, a->N, a->K, ks, smem, cudaGetErrorString(e));
    pb_set_error(what);
    return PB_ERR_CUDA;
  }
  return PB_OK;
}
