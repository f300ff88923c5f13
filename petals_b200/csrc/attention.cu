// petals_b200 — flash attention over the paged KV cache (prefill, chunked prefill, decode, verify).
//
// One kernel covers every inference shape the reference serves through rpc_inference
// (src/petals/server/backend.py:111-144): T new tokens per sequence attend to `pos + T` cached
// positions (the new K/V were appended by rope_kv.cu just before). Design points:
//
//  * GQA packing: the M dimension of a CTA tile is (token, q-head-within-group) pairs of ONE kv head,
//    so K/V tiles are loaded once per group instead of once per q head — the reference instead
//    materialises `repeat_kv` copies and a [B,Hq,T,L] logits tensor.
//  * KV tile == one cache page (64 tokens): a tile is one contiguous 64·D·2-byte slab found through
//    the session's block table, streamed with cp.async into XOR-swizzled shared memory (double
//    buffered) and consumed with ldmatrix + mma.sync m16n8k16 (bf16, fp32 accumulate).
//  * online softmax in the exp2 domain; causal, sliding-window (Mixtral) and ALiBi (BLOOM/Falcon)
//    masks are generated from positions — no mask tensor is ever built
//    (reference builds a 4-D mask on the host per call: src/petals/models/llama/block.py:250-259).
//  * split-KV for decode: blockIdx.z owns a slice of the KV pages and writes an (O, LSE) partial;
//    a tiny combine kernel merges the slices. The cache length is read from device memory so the
//    launch is CUDA-graph replayable while the sequence grows.
//
// Prefill-sized tiles (>= 128 packed query rows per kv head) are dispatched to the tcgen05/TMEM kernel in attention_tc.cu; this
// mma.sync kernel serves decode and short verify steps, where one 16-row MMA tile per kv head is KV-bandwidth bound.
#include "common.cuh"
#include "petals_b200.h"

namespace pb {

struct AttnParams {
  const __nv_bfloat16* q;
  const __nv_bfloat16* k_pool;
  const __nv_bfloat16* v_pool;
  const int* block_table;
  const int* pos_ptr;
  __nv_bfloat16* out;
  float* partial_o;
  float* partial_lse;
  const float* alibi;
  float scale_log2;
  int B, T, Hq, Hkv, page, max_pages, window, splits, pos_static, num_pages;
  int* split_counter;   // [m_tiles * B * Hkv] zero-initialised; non-null fuses the split-KV combine into this kernel
  float* lse_out;       // optional [B*T*Hq]: log2-domain log-sum-exp of every query row (saved for the backward pass; splits == 1)
};

PB_DEVICE void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
PB_DEVICE void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
PB_DEVICE void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

PB_DEVICE void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
PB_DEVICE void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
PB_DEVICE void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// byte offset of 16-byte chunk `c` of row `r` in a [rows][D] bf16 tile with XOR swizzle
template <int D>
PB_DEVICE uint32_t swz(int r, int c) {
  return static_cast<uint32_t>(r * (D * 2) + ((c ^ (r & 7)) << 4));
}

template <int D>
__global__ void __launch_bounds__(128) attn_fwd_kernel(const AttnParams p) {
  pdl_trigger();
  pdl_wait();  // q / KV pages come from the RoPE+append kernel launched just before
  constexpr int BM = 64, BN = 64, CH = D / 8, KS = D / 16;
  constexpr int TILE_BYTES = BN * D * 2;
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sQ = smem_u32(smem);
  const uint32_t sK = sQ + BM * D * 2;
  const uint32_t sV = sK + 2 * TILE_BYTES;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mt = blockIdx.x, bh = blockIdx.y, split = blockIdx.z;
  const int b = bh / p.Hkv, kvh = bh - b * p.Hkv;
  const int G = p.Hq / p.Hkv;
  const int pos0 = p.pos_ptr != nullptr ? *p.pos_ptr : p.pos_static;
  const int rows_total = p.T * G;
  const int m0 = mt * BM;
  const int t_min = m0 / G;
  const int t_max = min(p.T - 1, (m0 + BM - 1) / G);
  const int kv_hi = pos0 + t_max + 1;
  const int n_tiles = (kv_hi + BN - 1) / BN;
  int tile_lo = 0;
  if (p.window > 0) {
    const int first = max(0, pos0 + t_min - p.window + 1);
    tile_lo = first / BN;
  }
  const int nt = max(n_tiles - tile_lo, 0);
  const int per = (nt + p.splits - 1) / p.splits;
  const int tb = tile_lo + split * per;
  const int te = min(n_tiles, tb + per);

  // ---- async loads -------------------------------------------------------------------------
  auto load_kv = [&](int tile, int buf) {
    int pg = p.block_table[static_cast<size_t>(b) * p.max_pages + tile];
    pg = min(max(pg, 0), p.num_pages - 1);  // a corrupt table entry must not read outside the pool
    const size_t base = (static_cast<size_t>(pg) * p.Hkv + kvh) * p.page * D;
    const __nv_bfloat16* ks = p.k_pool + base;
    const __nv_bfloat16* vs = p.v_pool + base;
#pragma unroll
    for (int i = 0; i < BN * CH / 128; ++i) {
      const int idx = tid + i * 128;
      const int r = idx / CH, c = idx % CH;
      cp_async16(sK + buf * TILE_BYTES + swz<D>(r, c), ks + r * D + c * 8, true);
      cp_async16(sV + buf * TILE_BYTES + swz<D>(r, c), vs + r * D + c * 8, true);
    }
  };
#pragma unroll
  for (int i = 0; i < BM * CH / 128; ++i) {
    const int idx = tid + i * 128;
    const int r = idx / CH, c = idx % CH;
    const int row = m0 + r;
    const bool ok = row < rows_total;
    const int t = ok ? row / G : 0, g = ok ? row % G : 0;
    const __nv_bfloat16* src = p.q + ((static_cast<size_t>(b) * p.T + t) * p.Hq + kvh * G + g) * D + c * 8;
    cp_async16(sQ + swz<D>(r, c), src, ok);
  }
  if (tb < te) load_kv(tb, 0);
  cp_async_commit();

  // ---- per-thread state ----------------------------------------------------------------------
  const int r_lo = warp * 16 + (lane >> 2);        // tile row of c0/c1; c2/c3 are r_lo + 8
  const bool warp_active = (m0 + warp * 16) < rows_total;
  float o[D / 8][4];
#pragma unroll
  for (int i = 0; i < D / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_i[2] = {-INFINITY, -INFINITY}, l_i[2] = {0.f, 0.f};
  uint32_t qf[KS][4];
  int qpos[2];
  float slope2 = 0.f;
  {
    const int row_a = m0 + r_lo, row_b = row_a + 8;
    qpos[0] = pos0 + min(row_a / G, p.T - 1);
    qpos[1] = pos0 + min(row_b / G, p.T - 1);
  }
  // ALiBi slope: rows of a thread may belong to different q heads (row % G) -> keep two slopes
  float slopes[2] = {0.f, 0.f};
  if (p.alibi != nullptr) {
    slopes[0] = p.alibi[kvh * G + (m0 + r_lo) % G] * 1.4426950408889634f;
    slopes[1] = p.alibi[kvh * G + (m0 + r_lo + 8) % G] * 1.4426950408889634f;
  }
  (void)slope2;

  for (int tile = tb; tile < te; ++tile) {
    const int buf = (tile - tb) & 1;
    if (tile + 1 < te) load_kv(tile + 1, buf ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();

    if (tile == tb) {
#pragma unroll
      for (int kk = 0; kk < KS; ++kk) {
        const int r = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int c = 2 * kk + (lane >> 4);
        ldsm_x4(sQ + swz<D>(r, c), qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3]);
      }
    }
    if (warp_active) {
      // ---- S = Q K^T -------------------------------------------------------------------------
      float s[8][4];
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
      const uint32_t kbase = sK + buf * TILE_BYTES;
#pragma unroll
      for (int kk = 0; kk < KS; ++kk) {
#pragma unroll
        for (int jn = 0; jn < 8; jn += 2) {
          const int r = jn * 8 + (lane >> 4) * 8 + (lane & 7);
          const int c = 2 * kk + ((lane >> 3) & 1);
          uint32_t b0, b1, b2, b3;
          ldsm_x4(kbase + swz<D>(r, c), b0, b1, b2, b3);
          mma_bf16(s[jn], qf[kk], b0, b1);
          mma_bf16(s[jn + 1], qf[kk], b2, b3);
        }
      }
      // ---- scale, mask, online softmax --------------------------------------------------------
      const int kv0 = tile * BN + 2 * (lane & 3);
      float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int hr = e >> 1;  // 0: row r_lo, 1: row r_lo + 8
          const int kvpos = kv0 + j * 8 + (e & 1);
          float v = s[j][e] * p.scale_log2;
          if (p.alibi != nullptr) v += slopes[hr] * static_cast<float>(kvpos - qpos[hr]);
          const bool ok = kvpos <= qpos[hr] && (p.window <= 0 || kvpos > qpos[hr] - p.window);
          v = ok ? v : -INFINITY;
          s[j][e] = v;
          mx[hr] = fmaxf(mx[hr], v);
        }
      }
      float alpha[2], m_use[2];
#pragma unroll
      for (int hr = 0; hr < 2; ++hr) {
        mx[hr] = fmaxf(mx[hr], __shfl_xor_sync(0xffffffffu, mx[hr], 1));
        mx[hr] = fmaxf(mx[hr], __shfl_xor_sync(0xffffffffu, mx[hr], 2));
        const float m_new = fmaxf(m_i[hr], mx[hr]);
        m_use[hr] = m_new == -INFINITY ? 0.f : m_new;
        alpha[hr] = exp2f(m_i[hr] - m_use[hr]);
        m_i[hr] = m_new;
      }
      float rs[2] = {0.f, 0.f};
      uint32_t pa[4][4];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float p0 = exp2f(s[j][0] - m_use[0]), p1 = exp2f(s[j][1] - m_use[0]);
        const float p2 = exp2f(s[j][2] - m_use[1]), p3 = exp2f(s[j][3] - m_use[1]);
        rs[0] += p0 + p1;
        rs[1] += p2 + p3;
        // C fragments of two adjacent 8-token tiles form one A fragment of a 16-token k-step
        pa[j >> 1][(j & 1) * 2 + 0] = pack_bf16(p0, p1);
        pa[j >> 1][(j & 1) * 2 + 1] = pack_bf16(p2, p3);
      }
      l_i[0] = l_i[0] * alpha[0] + rs[0];
      l_i[1] = l_i[1] * alpha[1] + rs[1];
#pragma unroll
      for (int dn = 0; dn < D / 8; ++dn) {
        o[dn][0] *= alpha[0]; o[dn][1] *= alpha[0];
        o[dn][2] *= alpha[1]; o[dn][3] *= alpha[1];
      }
      // ---- O += P V ----------------------------------------------------------------------------
      const uint32_t vbase = sV + buf * TILE_BYTES;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
        for (int dn = 0; dn < D / 8; dn += 2) {
          const int r = kk * 16 + ((lane >> 3) & 1) * 8 + (lane & 7);
          const int c = dn + (lane >> 4);
          uint32_t b0, b1, b2, b3;
          ldsm_x4_t(vbase + swz<D>(r, c), b0, b1, b2, b3);
          mma_bf16(o[dn], pa[kk], b0, b1);
          mma_bf16(o[dn + 1], pa[kk], b2, b3);
        }
      }
    }
    __syncthreads();
  }
  cp_async_wait<0>();

  // ---- finalize ---------------------------------------------------------------------------------
  if (warp_active) {
#pragma unroll
  for (int hr = 0; hr < 2; ++hr) {
    l_i[hr] += __shfl_xor_sync(0xffffffffu, l_i[hr], 1);
    l_i[hr] += __shfl_xor_sync(0xffffffffu, l_i[hr], 2);
  }
#pragma unroll
  for (int hr = 0; hr < 2; ++hr) {
    const int row = m0 + r_lo + hr * 8;
    if (row >= rows_total) continue;
    const int t = row / G, g = row % G;
    const size_t rowid = (static_cast<size_t>(b) * p.T + t) * p.Hq + kvh * G + g;
    const float inv = l_i[hr] > 0.f ? 1.f / l_i[hr] : 0.f;
    if (p.splits == 1) {
      __nv_bfloat16* dst = p.out + rowid * D + 2 * (lane & 3);
#pragma unroll
      for (int dn = 0; dn < D / 8; ++dn)
        *reinterpret_cast<uint32_t*>(dst + dn * 8) = pack_bf16(o[dn][hr * 2] * inv, o[dn][hr * 2 + 1] * inv);
      if (p.lse_out != nullptr && (lane & 3) == 0) p.lse_out[rowid] = l_i[hr] > 0.f ? m_i[hr] + log2f(l_i[hr]) : -INFINITY;
    } else {
      const size_t R = static_cast<size_t>(p.B) * p.T * p.Hq;
      float* dst = p.partial_o + (static_cast<size_t>(split) * R + rowid) * D + 2 * (lane & 3);
#pragma unroll
      for (int dn = 0; dn < D / 8; ++dn)
        *reinterpret_cast<float2*>(dst + dn * 8) = make_float2(o[dn][hr * 2] * inv, o[dn][hr * 2 + 1] * inv);
      if ((lane & 3) == 0)
        p.partial_lse[static_cast<size_t>(split) * R + rowid] = l_i[hr] > 0.f ? m_i[hr] + log2f(l_i[hr]) : -INFINITY;
    }
  }
  }  // warp_active

  // ---- fused split-KV combine: the last split CTA of this (m tile, sequence, kv head) to finish merges the partials, so the
  // decode step needs no separate combine launch (one kernel boundary less per block and token)
  if (p.splits > 1 && p.split_counter != nullptr) {
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      int* ctr = p.split_counter + (bh * gridDim.x + mt);
      const int prev = atomicAdd(ctr, 1);
      s_last = prev == p.splits - 1;
      if (s_last) *ctr = 0;  // self-resetting: ready for the next launch / graph replay
    }
    __syncthreads();
    if (s_last) {
      __threadfence();
      const size_t R = static_cast<size_t>(p.B) * p.T * p.Hq;
      const int rows_here = min(BM, rows_total - m0);
      for (int idx = tid; idx < rows_here * (D / 4); idx += 128) {
        const int rr = idx / (D / 4), c4 = idx - rr * (D / 4);
        const int row = m0 + rr;
        const int t = row / G, g = row - t * G;
        const size_t rowid = (static_cast<size_t>(b) * p.T + t) * p.Hq + kvh * G + g;
        float mx = -INFINITY;
        for (int sp = 0; sp < p.splits; ++sp) mx = fmaxf(mx, __ldcg(p.partial_lse + sp * R + rowid));
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        float wsum = 0.f;
        for (int sp = 0; sp < p.splits; ++sp) {
          const float l = __ldcg(p.partial_lse + sp * R + rowid);
          const float w = l == -INFINITY ? 0.f : exp2f(l - mx);
          const float4 v = __ldcg(reinterpret_cast<const float4*>(p.partial_o + (sp * R + rowid) * D) + c4);
          wsum += w;
          acc.x += w * v.x; acc.y += w * v.y; acc.z += w * v.z; acc.w += w * v.w;
        }
        const float inv = wsum > 0.f ? 1.f / wsum : 0.f;
        uint2 o2;
        o2.x = pack_bf16(acc.x * inv, acc.y * inv);
        o2.y = pack_bf16(acc.z * inv, acc.w * inv);
        *reinterpret_cast<uint2*>(p.out + rowid * D + c4 * 4) = o2;
      }
    }
  }
}

// merge split-KV partials: one CTA per (token, head) row, D threads
__global__ void attn_combine_kernel(const float* __restrict__ po, const float* __restrict__ plse,
                                    __nv_bfloat16* __restrict__ out, int splits, size_t R, int D) {
  pdl_trigger();
  pdl_wait();
  const size_t row = blockIdx.x;
  const int d = threadIdx.x;
  float mx = -INFINITY;
  for (int s = 0; s < splits; ++s) mx = fmaxf(mx, plse[s * R + row]);
  float acc = 0.f, wsum = 0.f;
  for (int s = 0; s < splits; ++s) {
    const float l = plse[s * R + row];
    const float w = l == -INFINITY ? 0.f : exp2f(l - mx);
    wsum += w;
    acc += w * po[(s * R + row) * D + d];
  }
  out[row * D + d] = __float2bfloat16_rn(wsum > 0.f ? acc / wsum : 0.f);
}

template <int D>
static int launch_attn(const PbAttnArgs* a, cudaStream_t s) {
  AttnParams p{};
  p.q = static_cast<const __nv_bfloat16*>(a->q);
  p.k_pool = static_cast<const __nv_bfloat16*>(a->k_pool);
  p.v_pool = static_cast<const __nv_bfloat16*>(a->v_pool);
  p.block_table = static_cast<const int*>(a->block_table);
  p.pos_ptr = static_cast<const int*>(a->pos_ptr);
  p.out = static_cast<__nv_bfloat16*>(a->out);
  p.partial_o = static_cast<float*>(a->partial_o);
  p.partial_lse = static_cast<float*>(a->partial_lse);
  p.alibi = static_cast<const float*>(a->alibi_slopes);
  p.split_counter = static_cast<int*>(a->split_counter);
  p.lse_out = static_cast<float*>(a->lse_out);
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.B = a->B; p.T = a->T; p.Hq = a->Hq; p.Hkv = a->Hkv; p.page = a->page; p.max_pages = a->max_pages;
  p.num_pages = a->num_pages > 0 ? a->num_pages : 0x7fffffff;
  p.window = a->window; p.splits = a->splits < 1 ? 1 : a->splits; p.pos_static = a->pos_static;
  const int G = a->Hq / a->Hkv;
  const int m_tiles = (a->T * G + 63) / 64;
  const size_t smem = static_cast<size_t>(64 * D * 2) * 5;
  auto kern = attn_fwd_kernel<D>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess)
    return PB_ERR_CUDA;
  dim3 grid(m_tiles, a->B * a->Hkv, p.splits);
  launch_pdl(kPdlAttn, kern, grid, dim3(128), smem, s, p);
  if (pb_check_launch("attention") != PB_OK) return PB_ERR_CUDA;
  if (p.splits > 1 && p.split_counter == nullptr) {
    const size_t R = static_cast<size_t>(a->B) * a->T * a->Hq;
    launch_pdl(kPdlCombine, attn_combine_kernel, dim3(static_cast<unsigned>(R)), dim3(D), 0, s, p.partial_o, p.partial_lse, p.out, p.splits, R, D);
    if (pb_check_launch("attention") != PB_OK) return PB_ERR_CUDA;
  }
  return PB_OK;
}

}  // namespace pb

namespace pb { int attention_tc_dispatch(const PbAttnArgs* a, cudaStream_t s); }

extern "C" int pb_attention(const PbAttnArgs* a, void* stream) {
  using namespace pb;
  if (a->page != 64 || a->Hq % a->Hkv) return PB_ERR_SHAPE;
  if (a->B * a->T == 0) return PB_OK;
  if (a->splits > 1 && (a->partial_o == nullptr || a->partial_lse == nullptr)) return PB_ERR_SHAPE;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  {
    // prefill-sized work (at least one full 128-row tile of packed query rows per kv head) runs on the tcgen05 kernel
    static const int env_tc = [] { const char* e = getenv("PETALS_B200_ATTN_TC"); return e ? atoi(e) : 1; }();
    const int G = a->Hq / a->Hkv;
    const bool tc_ok = a->splits <= 1 && a->num_pages > 0 && (a->D == 128 || a->D == 64);
    if (tc_ok && (a->impl == 2 || (a->impl == 0 && env_tc && a->T * G >= 128))) return attention_tc_dispatch(a, s);
    if (a->impl == 2) return PB_ERR_UNSUPPORTED;
  }
  if (a->D == 128) return launch_attn<128>(a, s);
  if (a->D == 64) return launch_attn<64>(a, s);
  return PB_ERR_UNSUPPORTED;
}
