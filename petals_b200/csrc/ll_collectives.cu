// petals_b200 — stand-alone halves of the LL all-reduce for tensor-parallel decode.
//
// The dense blocks fuse the all-reduce of a row-parallel projection into the GEMV that produces the partial sums (epilogue: LL pushes)
// and into the GEMV that consumes the sum (prologue: poll + add), see linear_decode.cu. A sparse-MoE block has no single producing or
// consuming GEMV — the consumer is the router, the producer the weighted combine of the selected experts — so these two small kernels
// do the same two halves on their own:
//   ll_reduce : out = x + sum_r parts[r]   (polls the {payload, tag} units of every source rank; same summation order and rounding as
//               the fused prologue, so every rank computes bit-identical sums and therefore identical routing decisions)
//   ll_push   : stores rows as {2 x bf16, tag} units into this rank's slot on every peer (st.relaxed.sys.v2: payload and validity in
//               one NVLink transaction)
// tag = epoch * mul + add, like linear_decode.cu. Reference behaviour this serves: tensor_parallel's all-reduce around the MoE block
// (src/petals/utils/convert_block.py:128 wraps every block, Mixtral included).
#include "common.cuh"
#include "petals_b200.h"

namespace pb {

struct LLPtrs { uint2* p[PB_MAX_PEERS]; };

__global__ void __launch_bounds__(256) ll_reduce_kernel(const __nv_bfloat16* __restrict__ x, LLPtrs parts, int R, const uint64_t* __restrict__ epoch,
                                                        uint32_t mul, uint32_t add, __nv_bfloat16* __restrict__ out, long n8, int* error_flag) {
  const uint32_t tag = static_cast<uint32_t>(*epoch) * mul + add;
  for (long v = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; v < n8; v += static_cast<long>(gridDim.x) * blockDim.x) {
    const long off = v << 3;
    const uint4 xv = __ldcg(reinterpret_cast<const uint4*>(x + off));
    float f[8] = {bf16_lo(xv.x), bf16_hi(xv.x), bf16_lo(xv.y), bf16_hi(xv.y), bf16_lo(xv.z), bf16_hi(xv.z), bf16_lo(xv.w), bf16_hi(xv.w)};
    for (int r = 0; r < R; ++r) {
      const uint4* src = reinterpret_cast<const uint4*>(parts.p[r] + (off >> 1));
      uint4 a, b;
      unsigned long long t0 = 0;
      for (unsigned spins = 0;; ++spins) {
        a = ld_relaxed_sys_v4(src);
        b = ld_relaxed_sys_v4(src + 1);
        if (a.y == tag && a.w == tag && b.y == tag && b.w == tag) break;
        if ((spins & 1023u) == 1023u) {
          const unsigned long long now = globaltimer_ns();
          if (t0 == 0) t0 = now;
          else if (now - t0 > PB_FLAG_TIMEOUT_NS) { if (error_flag != nullptr) atomicExch(error_flag, 1); break; }
        }
      }
      f[0] += bf16_lo(a.x); f[1] += bf16_hi(a.x); f[2] += bf16_lo(a.z); f[3] += bf16_hi(a.z);
      f[4] += bf16_lo(b.x); f[5] += bf16_hi(b.x); f[6] += bf16_lo(b.z); f[7] += bf16_hi(b.z);
    }
    uint4 o;
    o.x = pack_bf16(f[0], f[1]); o.y = pack_bf16(f[2], f[3]); o.z = pack_bf16(f[4], f[5]); o.w = pack_bf16(f[6], f[7]);
    *reinterpret_cast<uint4*>(out + off) = o;
  }
}

__global__ void __launch_bounds__(256) ll_push_kernel(const uint32_t* __restrict__ x, LLPtrs dst, int R, const uint64_t* __restrict__ epoch, uint32_t mul,
                                                      uint32_t add, long n2) {
  const uint32_t tag = static_cast<uint32_t>(*epoch) * mul + add;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n2; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const uint32_t v = x[i];
    for (int r = 0; r < R; ++r) st_relaxed_sys_v2(dst.p[r] + i, v, tag);
  }
}

}  // namespace pb

using namespace pb;

extern "C" int pb_ll_reduce(const void* x, void* const* parts, int R, const void* epoch, unsigned mul, unsigned add, void* out, long n_values,
                            void* error_flag, void* stream) {
  if (R < 1 || R > PB_MAX_PEERS || n_values <= 0 || (n_values & 7)) return PB_ERR_SHAPE;
  LLPtrs p{};
  for (int r = 0; r < R; ++r) p.p[r] = static_cast<uint2*>(parts[r]);
  const long n8 = n_values >> 3;
  const int grid = static_cast<int>((n8 + 255) / 256 < 64 ? (n8 + 255) / 256 : 64);
  ll_reduce_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(x), p, R, static_cast<const uint64_t*>(epoch), mul,
                                                                       add, static_cast<__nv_bfloat16*>(out), n8, static_cast<int*>(error_flag));
  return pb_check_launch("ll_reduce");
}

extern "C" int pb_ll_push(const void* x, void* const* dst, int R, const void* epoch, unsigned mul, unsigned add, long n_values, void* stream) {
  if (R < 1 || R > PB_MAX_PEERS || n_values <= 0 || (n_values & 1)) return PB_ERR_SHAPE;
  LLPtrs p{};
  for (int r = 0; r < R; ++r) p.p[r] = static_cast<uint2*>(dst[r]);
  const long n2 = n_values >> 1;
  const int grid = static_cast<int>((n2 + 255) / 256 < 64 ? (n2 + 255) / 256 : 64);
  ll_push_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint32_t*>(x), p, R, static_cast<const uint64_t*>(epoch), mul, add, n2);
  return pb_check_launch("ll_push");
}
