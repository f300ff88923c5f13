// petals_b200 — fused (bias) + rotary embedding + paged KV-cache append.
//
// Input is the fused QKV projection of the new tokens. One pass:
//   q  -> (+bias) -> RoPE -> q_out (token-major, ready for the attention kernel)
//   k  -> (+bias) -> RoPE -> written in place into the session's KV pages at position pos + t
//   v  -> (+bias)         -> written into the KV pages
// The cache position `pos` lives in device memory so that the whole decode step can be replayed
// from a CUDA graph. Pages are [Hkv, PAGE, D] slabs addressed through a per-session block table,
// which replaces the reference's per-step torch.cat of the entire past plus BLOOM<->Llama layout
// permutes (src/petals/models/llama/block.py:97-100,280-300,
// src/petals/server/backend.py:171-181) and its six per-op CUDA graphs (G1, G4, G5).
//
// RoPE uses the HF "rotate_half" convention with a precomputed fp32 cos/sin table [max_pos, D/2]
// (so Llama-3 frequency scaling is a host-side table property) and reproduces HF's bf16 rounding:
// bf16(bf16(x*cos) + bf16(rot(x)*sin)).
#include "common.cuh"
#include "petals_b200.h"

namespace pb {

struct RopeKvParams {
  const __nv_bfloat16* qkv;
  __nv_bfloat16* q_out;
  __nv_bfloat16* k_pool;
  __nv_bfloat16* v_pool;
  const int* block_table;
  const int* pos_ptr;
  const float* cos;
  const float* sin;
  const __nv_bfloat16* bias;
  int B, T, Hq, Hkv, D, page, max_pages, max_pos, interleaved, num_pages;
  int* error_flag;
};

PB_DEVICE float rbf(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// grid: (B*T, Hq + 2*Hkv); block: D/2 threads — thread i owns the rotary pair (i, i + D/2).
__global__ void rope_kv_kernel(const RopeKvParams p) {
  pdl_trigger();  // let the attention kernel's CTAs get scheduled behind ours
  pdl_wait();     // the fused QKV projection (previous kernel) must be complete and visible
  const int tok = blockIdx.x;            // b * T + t
  const int head = blockIdx.y;           // [0,Hq): q, [Hq,Hq+Hkv): k, rest: v
  const int b = tok / p.T, t = tok - b * p.T;
  const int i = threadIdx.x;
  const int half = p.D >> 1;
  const int pos = *p.pos_ptr + t;
  const int G = p.Hq / p.Hkv;

  // locate this head's D-vector in the fused projection output
  int kind, h;  // 0 q, 1 k, 2 v
  if (head < p.Hq) { kind = 0; h = head; }
  else if (head < p.Hq + p.Hkv) { kind = 1; h = head - p.Hq; }
  else { kind = 2; h = head - p.Hq - p.Hkv; }
  size_t col;
  if (!p.interleaved) {
    col = static_cast<size_t>(head) * p.D;  // [q heads | k heads | v heads]
  } else {
    // Falcon / BLOOM fused layout: per kv group [G q heads, 1 k head, 1 v head]
    if (kind == 0) col = (static_cast<size_t>(h / G) * (G + 2) + (h % G)) * p.D;
    else if (kind == 1) col = (static_cast<size_t>(h) * (G + 2) + G) * p.D;
    else col = (static_cast<size_t>(h) * (G + 2) + G + 1) * p.D;
  }
  const size_t width = static_cast<size_t>(p.Hq + 2 * p.Hkv) * p.D;
  const __nv_bfloat16* src = p.qkv + static_cast<size_t>(tok) * width + col;
  float x0 = __bfloat162float(src[i]);
  float x1 = __bfloat162float(src[i + half]);
  if (p.bias != nullptr) {
    x0 = rbf(x0 + __bfloat162float(p.bias[col + i]));
    x1 = rbf(x1 + __bfloat162float(p.bias[col + i + half]));
  }
  if (kind != 2 && p.cos != nullptr) {
    const int pp = pos < p.max_pos ? pos : p.max_pos - 1;
    const float c = rbf(p.cos[static_cast<size_t>(pp) * half + i]);
    const float s = rbf(p.sin[static_cast<size_t>(pp) * half + i]);
    const float y0 = rbf(rbf(x0 * c) + rbf(-x1 * s));
    const float y1 = rbf(rbf(x1 * c) + rbf(x0 * s));
    x0 = y0; x1 = y1;
  }
  if (kind == 0) {
    __nv_bfloat16* dst = p.q_out + (static_cast<size_t>(tok) * p.Hq + h) * p.D;
    dst[i] = __float2bfloat16_rn(x0);
    dst[i + half] = __float2bfloat16_rn(x1);
  } else {
    const int pg_idx = pos / p.page;
    if (pg_idx >= p.max_pages) { if (i == 0) atomicExch(p.error_flag, 2); return; }
    const int pg = p.block_table[static_cast<size_t>(b) * p.max_pages + pg_idx];
    if (pg < 0 || pg >= p.num_pages) { if (i == 0) atomicExch(p.error_flag, 2); return; }  // never write outside the pool
    __nv_bfloat16* pool = kind == 1 ? p.k_pool : p.v_pool;
    __nv_bfloat16* dst = pool + ((static_cast<size_t>(pg) * p.Hkv + h) * p.page + (pos % p.page)) * p.D;
    dst[i] = __float2bfloat16_rn(x0);
    dst[i + half] = __float2bfloat16_rn(x1);
  }
}

// Copy whole KV pages (all layers / K and V slabs) — used for copy-on-write when beam search
// duplicates a hypothesis (src/petals/server/backend.py:154-158 gathers the entire
// cache instead).
__global__ void kv_copy_pages_kernel(uint4* pool, const int* __restrict__ src, const int* __restrict__ dst,
                                     long page_vecs, long slab_stride_vecs) {
  const int pair = blockIdx.x, slab = blockIdx.y;
  const uint4* s = pool + slab * slab_stride_vecs + static_cast<long>(src[pair]) * page_vecs;
  uint4* d = pool + slab * slab_stride_vecs + static_cast<long>(dst[pair]) * page_vecs;
  for (long v = threadIdx.x; v < page_vecs; v += blockDim.x) d[v] = s[v];
}

}  // namespace pb

extern "C" int pb_rope_kv(const PbRopeKvArgs* a, void* stream) {
  using namespace pb;
  if (a->D & 1 || a->D / 2 > 1024 || a->Hq % a->Hkv) return PB_ERR_SHAPE;
  if (a->B * a->T == 0) return PB_OK;
  RopeKvParams p{};
  p.qkv = static_cast<const __nv_bfloat16*>(a->qkv);
  p.q_out = static_cast<__nv_bfloat16*>(a->q_out);
  p.k_pool = static_cast<__nv_bfloat16*>(a->k_pool);
  p.v_pool = static_cast<__nv_bfloat16*>(a->v_pool);
  p.block_table = static_cast<const int*>(a->block_table);
  p.pos_ptr = static_cast<const int*>(a->pos_ptr);
  p.cos = static_cast<const float*>(a->cos);
  p.sin = static_cast<const float*>(a->sin);
  p.bias = static_cast<const __nv_bfloat16*>(a->qkv_bias);
  p.B = a->B; p.T = a->T; p.Hq = a->Hq; p.Hkv = a->Hkv; p.D = a->D; p.page = a->page;
  p.max_pages = a->max_pages; p.max_pos = a->max_pos; p.interleaved = a->interleaved_qkv;
  p.num_pages = a->num_pages > 0 ? a->num_pages : 0x7fffffff;
  p.error_flag = static_cast<int*>(a->error_flag);
  dim3 grid(a->B * a->T, a->Hq + 2 * a->Hkv);
  launch_pdl(kPdlRope, rope_kv_kernel, grid, dim3(a->D / 2), 0, static_cast<cudaStream_t>(stream), p);
  return pb_check_launch("rope_kv");
}

extern "C" int pb_kv_copy_pages(void* pool, const void* src_pages, const void* dst_pages, int n, long page_elems,
                                long slab_stride_elems, int n_slabs, void* stream) {
  using namespace pb;
  if (n == 0) return PB_OK;
  if ((page_elems & 7) || (slab_stride_elems & 7)) return PB_ERR_SHAPE;
  dim3 grid(n, n_slabs);
  kv_copy_pages_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<uint4*>(pool), static_cast<const int*>(src_pages), static_cast<const int*>(dst_pages),
      page_elems >> 3, slab_stride_elems >> 3);
  return pb_check_launch("rope_kv");
}
