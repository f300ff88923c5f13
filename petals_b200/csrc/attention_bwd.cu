// petals_b200 — backward of causal (GQA) attention for the prompt-tuning / fine-tuning path.
//
// Stages are frozen, so `rpc_backward` (reference: src/petals/server/block_functions.py:84-141, server/backend.py:106-109) only needs
// gradients with respect to activations: dQ, dK, dV given dO. The reference gets them from torch.autograd over an attention that
// materialises the [B, Hq, T, L] probabilities (models/llama/block.py:108-115). Here the probabilities are recomputed tile by
// tile from the saved log-sum-exp (flash-attention backward), with the same GQA row packing and paged K/V tiles as the forward
// kernel (attention.cu): the M dimension of a tile is (token, q-head-in-group) pairs of ONE kv head, so dK / dV of a kv head
// accumulate over its whole query group inside one CTA — no atomics anywhere.
//
//   delta[row]  = sum_d dO[row, d] * O[row, d]                                   (attn_delta_kernel)
//   P           = exp2(scale*log2e * Q K^T - LSE)                                 (recomputed)
//   dV          = P^T dO                      dP = dO V^T
//   dS          = P o (dP - delta)            dQ = scale * dS K        dK = scale * dS^T Q
//
// Two kernels instead of FA2's single kernel with fp32 atomics on dQ: `attn_bwd_dq_kernel` owns 64 packed query rows and loops
// over KV pages; `attn_bwd_dkdv_kernel` owns one KV page of one kv head and loops over the packed query rows that can see it,
// working on the TRANSPOSED tiles (S^T = K Q^T, dP^T = V dO^T) so that P^T / dS^T come out of the MMA directly in the
// accumulator layout that converts to an A operand. Deterministic, and each kernel is the forward's loop with one more MMA.
// Tensor-core path: mma.sync m16n8k16 bf16 + ldmatrix on XOR-swizzled tiles (same fragments as attention.cu).
#include "common.cuh"
#include "petals_b200.h"

extern "C" int pb_set_error(const char* msg);

namespace pb {
namespace bwd {

struct Params {
  const __nv_bfloat16* q; const __nv_bfloat16* k_pool; const __nv_bfloat16* v_pool; const int* block_table;
  const __nv_bfloat16* out; const __nv_bfloat16* d_out; const float* lse; float* delta;
  __nv_bfloat16* dq; __nv_bfloat16* dk; __nv_bfloat16* dv;
  float scale, scale_log2;
  int B, T, Hq, Hkv, max_pages, num_pages;
};

PB_DEVICE void cp16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
PB_DEVICE void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
PB_DEVICE void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
PB_DEVICE void ldsm4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
PB_DEVICE void ldsm4t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
PB_DEVICE void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <int D>
PB_DEVICE uint32_t swz(int r, int c) { return static_cast<uint32_t>(r * (D * 2) + ((c ^ (r & 7)) << 4)); }

// ---- delta = rowsum(dO * O): one warp per (token, head) row --------------------------------------------------------------
__global__ void attn_delta_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ d_o, float* __restrict__ delta, long rows, int D) {
  const long row = static_cast<long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  float s = 0.f;
  for (int d = lane * 2; d < D; d += 64) {
    const uint32_t a = *reinterpret_cast<const uint32_t*>(o + row * D + d), b = *reinterpret_cast<const uint32_t*>(d_o + row * D + d);
    s += bf16_lo(a) * bf16_lo(b) + bf16_hi(a) * bf16_hi(b);
  }
  s = warp_sum(s);
  if (lane == 0) delta[row] = s;
}

// Load 64 packed query rows (token, head-in-group) of `src` [B*T, Hq*D] into a swizzled [64][D] tile.
template <int D>
PB_DEVICE void load_rows(uint32_t dst, const __nv_bfloat16* src, const Params& p, int b, int kvh, int G, int m0, int nrows, int rows_total, int tid) {
  constexpr int CH = D / 8;
  for (int idx = tid; idx < nrows * CH; idx += 128) {
    const int r = idx / CH, c = idx - r * CH;
    const int row = m0 + r;
    const bool ok = row < rows_total;
    const int t = ok ? row / G : 0, g = ok ? row - t * G : 0;
    cp16(dst + swz<D>(r, c), src + ((static_cast<size_t>(b) * p.T + t) * p.Hq + kvh * G + g) * D + c * 8, ok);
  }
}
template <int D>
PB_DEVICE void load_page(uint32_t dst, const __nv_bfloat16* pool, const Params& p, int b, int kvh, int tile, int tid) {
  constexpr int CH = D / 8;
  int pg = tile < p.max_pages ? p.block_table[static_cast<size_t>(b) * p.max_pages + tile] : 0;
  pg = min(max(pg, 0), p.num_pages - 1);
  const __nv_bfloat16* src = pool + (static_cast<size_t>(pg) * p.Hkv + kvh) * 64 * D;
  for (int idx = tid; idx < 64 * CH; idx += 128) {
    const int r = idx / CH, c = idx - r * CH;
    cp16(dst + swz<D>(r, c), src + r * D + c * 8, true);
  }
}

// ---- dQ: CTA = 64 packed query rows of one (sequence, kv head); loop over the KV pages they can see -------------------------
template <int D>
__global__ void __launch_bounds__(128) attn_bwd_dq_kernel(const Params p) {
  constexpr int BM = 64, BN = 64, KS = D / 16, TILE = BN * D * 2;
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sQ = smem_u32(smem), sdO = sQ + TILE, sK = sdO + TILE, sV = sK + 2 * TILE;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mt = blockIdx.x, bh = blockIdx.y;
  const int b = bh / p.Hkv, kvh = bh - b * p.Hkv, G = p.Hq / p.Hkv;
  const int rows_total = p.T * G, m0 = mt * BM;
  const int t_max = min(p.T - 1, (m0 + BM - 1) / G);
  const int n_tiles = t_max / BN + 1;

  load_rows<D>(sQ, p.q, p, b, kvh, G, m0, BM, rows_total, tid);
  load_rows<D>(sdO, p.d_out, p, b, kvh, G, m0, BM, rows_total, tid);
  load_page<D>(sK, p.k_pool, p, b, kvh, 0, tid);
  load_page<D>(sV, p.v_pool, p, b, kvh, 0, tid);
  cp_commit();

  const int r_lo = warp * 16 + (lane >> 2);
  const bool warp_active = (m0 + warp * 16) < rows_total;
  float acc[D / 8][4];
#pragma unroll
  for (int i = 0; i < D / 8; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  float lse[2], dl[2];
  int qpos[2];
#pragma unroll
  for (int hr = 0; hr < 2; ++hr) {
    const int row = m0 + r_lo + hr * 8;
    const bool ok = row < rows_total;
    const int t = ok ? row / G : 0, g = ok ? row - t * G : 0;
    const size_t rowid = (static_cast<size_t>(b) * p.T + t) * p.Hq + kvh * G + g;
    lse[hr] = ok ? p.lse[rowid] : INFINITY;   // +inf: every probability of a padding row becomes 0
    dl[hr] = ok ? p.delta[rowid] : 0.f;
    qpos[hr] = ok ? t : -1;
  }
  uint32_t qf[KS][4];

  for (int tile = 0; tile < n_tiles; ++tile) {
    const int buf = tile & 1;
    if (tile + 1 < n_tiles) {
      load_page<D>(sK + (buf ^ 1) * TILE, p.k_pool, p, b, kvh, tile + 1, tid);
      load_page<D>(sV + (buf ^ 1) * TILE, p.v_pool, p, b, kvh, tile + 1, tid);
    }
    cp_commit();
    cp_wait<1>();
    __syncthreads();
    if (tile == 0) {
#pragma unroll
      for (int kk = 0; kk < KS; ++kk) {
        const int r = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, c = 2 * kk + (lane >> 4);
        ldsm4(sQ + swz<D>(r, c), qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3]);
      }
    }
    if (warp_active) {
      const uint32_t kb = sK + buf * TILE, vb = sV + buf * TILE;
      float s[8][4], dp[8][4];
#pragma unroll
      for (int j = 0; j < 8; ++j) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f; dp[j][0] = dp[j][1] = dp[j][2] = dp[j][3] = 0.f; }
#pragma unroll
      for (int kk = 0; kk < KS; ++kk) {
        uint32_t df[4];
        {
          const int r = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, c = 2 * kk + (lane >> 4);
          ldsm4(sdO + swz<D>(r, c), df[0], df[1], df[2], df[3]);
        }
#pragma unroll
        for (int jn = 0; jn < 8; jn += 2) {
          const int r = jn * 8 + (lane >> 4) * 8 + (lane & 7), c = 2 * kk + ((lane >> 3) & 1);
          uint32_t b0, b1, b2, b3;
          ldsm4(kb + swz<D>(r, c), b0, b1, b2, b3);
          mma16816(s[jn], qf[kk], b0, b1);
          mma16816(s[jn + 1], qf[kk], b2, b3);
          ldsm4(vb + swz<D>(r, c), b0, b1, b2, b3);
          mma16816(dp[jn], df, b0, b1);
          mma16816(dp[jn + 1], df, b2, b3);
        }
      }
      // dS = P o (dP - delta), as bf16 A fragments
      uint32_t dsa[4][4];
      const int kv0 = tile * BN + 2 * (lane & 3);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float ds[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int hr = e >> 1, kvpos = kv0 + j * 8 + (e & 1);
          const bool ok = kvpos <= qpos[hr];  // select, never multiply: a masked key's page row may hold anything (0 * NaN = NaN)
          ds[e] = ok ? exp2f(s[j][e] * p.scale_log2 - lse[hr]) * (dp[j][e] - dl[hr]) : 0.f;
        }
        dsa[j >> 1][(j & 1) * 2 + 0] = pack_bf16(ds[0], ds[1]);
        dsa[j >> 1][(j & 1) * 2 + 1] = pack_bf16(ds[2], ds[3]);
      }
      // dQ += dS K
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
        for (int dn = 0; dn < D / 8; dn += 2) {
          const int r = kk * 16 + ((lane >> 3) & 1) * 8 + (lane & 7), c = dn + (lane >> 4);
          uint32_t b0, b1, b2, b3;
          ldsm4t(kb + swz<D>(r, c), b0, b1, b2, b3);
          mma16816(acc[dn], dsa[kk], b0, b1);
          mma16816(acc[dn + 1], dsa[kk], b2, b3);
        }
      }
    }
    __syncthreads();
  }
  cp_wait<0>();
  if (warp_active) {
#pragma unroll
    for (int hr = 0; hr < 2; ++hr) {
      const int row = m0 + r_lo + hr * 8;
      if (row >= rows_total) continue;
      const int t = row / G, g = row - t * G;
      __nv_bfloat16* dst = p.dq + ((static_cast<size_t>(b) * p.T + t) * p.Hq + kvh * G + g) * D + 2 * (lane & 3);
#pragma unroll
      for (int dn = 0; dn < D / 8; ++dn)
        *reinterpret_cast<uint32_t*>(dst + dn * 8) = pack_bf16(acc[dn][hr * 2] * p.scale, acc[dn][hr * 2 + 1] * p.scale);
    }
  }
}

// ---- dK, dV: CTA = one KV page (64 keys) of one (sequence, kv head); loop over 32-row tiles of the packed query rows that see it --
template <int D>
__global__ void __launch_bounds__(128) attn_bwd_dkdv_kernel(const Params p) {
  constexpr int BQ = 32, KS = D / 16, TILE = 64 * D * 2, QTILE = BQ * D * 2;
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sK = smem_u32(smem), sV = sK + TILE, sQ = sV + TILE, sdO = sQ + 2 * QTILE;
  float* s_lse = reinterpret_cast<float*>(smem + 2 * TILE + 4 * QTILE);  // [2][BQ]
  float* s_dl = s_lse + 2 * BQ;                                          // [2][BQ]
  int* s_qpos = reinterpret_cast<int*>(s_dl + 2 * BQ);                   // [2][BQ]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nt = blockIdx.x, bh = blockIdx.y;
  const int b = bh / p.Hkv, kvh = bh - b * p.Hkv, G = p.Hq / p.Hkv;
  const int rows_total = p.T * G, n0 = nt * 64;
  const int q_lo = (n0 * G) / BQ, q_hi = (rows_total + BQ - 1) / BQ;  // causal: token t >= n0 sees this page

  auto load_q = [&](int qt, int buf) {
    load_rows<D>(sQ + buf * QTILE, p.q, p, b, kvh, G, qt * BQ, BQ, rows_total, tid);
    load_rows<D>(sdO + buf * QTILE, p.d_out, p, b, kvh, G, qt * BQ, BQ, rows_total, tid);
    if (tid < BQ) {
      const int row = qt * BQ + tid;
      const bool ok = row < rows_total;
      const int t = ok ? row / G : 0, g = ok ? row - t * G : 0;
      const size_t rowid = (static_cast<size_t>(b) * p.T + t) * p.Hq + kvh * G + g;
      s_lse[buf * BQ + tid] = ok ? p.lse[rowid] : INFINITY;
      s_dl[buf * BQ + tid] = ok ? p.delta[rowid] : 0.f;
      s_qpos[buf * BQ + tid] = ok ? t : -1;
    }
  };
  load_page<D>(sK, p.k_pool, p, b, kvh, nt, tid);
  load_page<D>(sV, p.v_pool, p, b, kvh, nt, tid);
  if (q_lo < q_hi) load_q(q_lo, 0);
  cp_commit();

  float dk[D / 8][4], dv[D / 8][4];
#pragma unroll
  for (int i = 0; i < D / 8; ++i) { dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f; dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f; }
  const int k_lo = warp * 16 + (lane >> 2);  // key rows of this thread's accumulator elements: k_lo, k_lo + 8
  const int kpos[2] = {n0 + k_lo, n0 + k_lo + 8};

  for (int qt = q_lo; qt < q_hi; ++qt) {
    const int buf = (qt - q_lo) & 1;
    if (qt + 1 < q_hi) load_q(qt + 1, buf ^ 1);
    cp_commit();
    cp_wait<1>();
    __syncthreads();
    const uint32_t qb = sQ + buf * QTILE, ob = sdO + buf * QTILE;
    // S^T = K Q^T and dP^T = V dO^T : [16 keys of this warp] x [32 query rows]
    float st[4][4], dpt[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { st[j][0] = st[j][1] = st[j][2] = st[j][3] = 0.f; dpt[j][0] = dpt[j][1] = dpt[j][2] = dpt[j][3] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < KS; ++kk) {
      uint32_t kf[4], vf[4];
      {
        const int r = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, c = 2 * kk + (lane >> 4);
        ldsm4(sK + swz<D>(r, c), kf[0], kf[1], kf[2], kf[3]);
        ldsm4(sV + swz<D>(r, c), vf[0], vf[1], vf[2], vf[3]);
      }
#pragma unroll
      for (int jn = 0; jn < 4; jn += 2) {
        const int r = jn * 8 + (lane >> 4) * 8 + (lane & 7), c = 2 * kk + ((lane >> 3) & 1);
        uint32_t b0, b1, b2, b3;
        ldsm4(qb + swz<D>(r, c), b0, b1, b2, b3);
        mma16816(st[jn], kf, b0, b1);
        mma16816(st[jn + 1], kf, b2, b3);
        ldsm4(ob + swz<D>(r, c), b0, b1, b2, b3);
        mma16816(dpt[jn], vf, b0, b1);
        mma16816(dpt[jn + 1], vf, b2, b3);
      }
    }
    // P^T and dS^T as A fragments (k dimension = query rows)
    uint32_t pa[2][4], dsa[2][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float pr[4], ds[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int hr = e >> 1, col = j * 8 + 2 * (lane & 3) + (e & 1);
        const float l = s_lse[buf * BQ + col];
        const bool ok = kpos[hr] <= s_qpos[buf * BQ + col];
        pr[e] = ok ? exp2f(st[j][e] * p.scale_log2 - l) : 0.f;
        ds[e] = ok ? pr[e] * (dpt[j][e] - s_dl[buf * BQ + col]) : 0.f;
      }
      pa[j >> 1][(j & 1) * 2 + 0] = pack_bf16(pr[0], pr[1]);
      pa[j >> 1][(j & 1) * 2 + 1] = pack_bf16(pr[2], pr[3]);
      dsa[j >> 1][(j & 1) * 2 + 0] = pack_bf16(ds[0], ds[1]);
      dsa[j >> 1][(j & 1) * 2 + 1] = pack_bf16(ds[2], ds[3]);
    }
    // dV += P^T dO ; dK += dS^T Q      (B operands: [query row = k][d = n] -> transposed ldmatrix)
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
#pragma unroll
      for (int dn = 0; dn < D / 8; dn += 2) {
        const int r = kk * 16 + ((lane >> 3) & 1) * 8 + (lane & 7), c = dn + (lane >> 4);
        uint32_t b0, b1, b2, b3;
        ldsm4t(ob + swz<D>(r, c), b0, b1, b2, b3);
        mma16816(dv[dn], pa[kk], b0, b1);
        mma16816(dv[dn + 1], pa[kk], b2, b3);
        ldsm4t(qb + swz<D>(r, c), b0, b1, b2, b3);
        mma16816(dk[dn], dsa[kk], b0, b1);
        mma16816(dk[dn + 1], dsa[kk], b2, b3);
      }
    }
    __syncthreads();
  }
  cp_wait<0>();
#pragma unroll
  for (int hr = 0; hr < 2; ++hr) {
    const int t = kpos[hr];
    if (t >= p.T) continue;
    const size_t o = ((static_cast<size_t>(b) * p.T + t) * p.Hkv + kvh) * D + 2 * (lane & 3);
#pragma unroll
    for (int dn = 0; dn < D / 8; ++dn) {
      *reinterpret_cast<uint32_t*>(p.dk + o + dn * 8) = pack_bf16(dk[dn][hr * 2] * p.scale, dk[dn][hr * 2 + 1] * p.scale);
      *reinterpret_cast<uint32_t*>(p.dv + o + dn * 8) = pack_bf16(dv[dn][hr * 2], dv[dn][hr * 2 + 1]);
    }
  }
}

template <int D>
static int launch(const PbAttnBwdArgs* a, cudaStream_t s) {
  Params p{};
  p.q = static_cast<const __nv_bfloat16*>(a->q); p.k_pool = static_cast<const __nv_bfloat16*>(a->k_pool);
  p.v_pool = static_cast<const __nv_bfloat16*>(a->v_pool); p.block_table = static_cast<const int*>(a->block_table);
  p.out = static_cast<const __nv_bfloat16*>(a->out); p.d_out = static_cast<const __nv_bfloat16*>(a->d_out);
  p.lse = static_cast<const float*>(a->lse); p.delta = static_cast<float*>(a->delta);
  p.dq = static_cast<__nv_bfloat16*>(a->dq); p.dk = static_cast<__nv_bfloat16*>(a->dk); p.dv = static_cast<__nv_bfloat16*>(a->dv);
  p.scale = a->scale; p.scale_log2 = a->scale * 1.4426950408889634f;
  p.B = a->B; p.T = a->T; p.Hq = a->Hq; p.Hkv = a->Hkv; p.max_pages = a->max_pages; p.num_pages = a->num_pages > 0 ? a->num_pages : 0x7fffffff;
  const long rows = static_cast<long>(a->B) * a->T * a->Hq;
  attn_delta_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, s>>>(p.out, p.d_out, p.delta, rows, D);
  if (pb_check_launch("attn_delta") != PB_OK) return PB_ERR_CUDA;
  const int G = a->Hq / a->Hkv;
  {
    const int smem = 6 * 64 * D * 2;
    auto k = attn_bwd_dq_kernel<D>;
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return PB_ERR_CUDA;
    k<<<dim3((a->T * G + 63) / 64, a->B * a->Hkv), 128, smem, s>>>(p);
    if (pb_check_launch("attn_bwd_dq") != PB_OK) return PB_ERR_CUDA;
  }
  {
    const int smem = 2 * 64 * D * 2 + 4 * 32 * D * 2 + 6 * 32 * 4;
    auto k = attn_bwd_dkdv_kernel<D>;
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return PB_ERR_CUDA;
    k<<<dim3((a->T + 63) / 64, a->B * a->Hkv), 128, smem, s>>>(p);
    if (pb_check_launch("attn_bwd_dkdv") != PB_OK) return PB_ERR_CUDA;
  }
  return PB_OK;
}

}  // namespace bwd
}  // namespace pb

extern "C" int pb_attention_bwd(const PbAttnBwdArgs* a, void* stream) {
  if (a->B * a->T == 0) return PB_OK;
  if (a->Hkv <= 0 || a->Hq % a->Hkv) { pb_set_error("attention_bwd: Hq must be a multiple of Hkv"); return PB_ERR_SHAPE; }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (a->D == 128) return pb::bwd::launch<128>(a, s);
  if (a->D == 64) return pb::bwd::launch<64>(a, s);
  pb_set_error("attention_bwd: head_dim must be 64 or 128");
  return PB_ERR_UNSUPPORTED;
}
