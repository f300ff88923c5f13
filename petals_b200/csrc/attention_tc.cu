// petals_b200 — prefill flash attention on the 5th-gen tensor cores (tcgen05 + TMEM + TMA) over the paged KV cache.
//
// Same contract as attention.cu (T new tokens per sequence attend to pos + T cached positions; causal / sliding-window /
// ALiBi masks generated from positions; GQA row packing), used when a kv-head has at least one full 128-row tile of
// (token, q-head-in-group) rows, i.e. prompt ingestion and chunked prefill. Decode / short verify steps stay on the
// split-KV mma.sync kernel.
//
// One CTA = 128 packed query rows of one (sequence, kv head); KV tile = 128 keys = two cache pages.
//
//   warp 0    TMA producer: per KV tile, 2 pages x D/64 boxes of [64 keys x 64 d] for K and for V (128B swizzle), 2 stages
//   warp 1    MMA issuer (one thread):  S = Q K^T   (M128 x N128 x K16 tcgen05.mma, D/16 steps, fp32 in TMEM, double buffered)
//                                       O' = P V    (M128 x N=D x K16, 8 steps; A = P from shared memory, B = V MN-major)
//   warps 2-9 softmax + accumulate: thread (row r, half h) pulls its 64 logits out of TMEM once (the S buffer is released
//             immediately), online softmax in the exp2 domain with the row maximum exchanged between the two halves through
//             shared memory, P (bf16) into 128B-swizzled shared memory; after the PV MMA, tcgen05.ld O' and
//             O = O * alpha + O' in registers (D/2 columns per thread; no TMEM read-modify-write) -> normalise, store bf16.
//
// Pipelining: S(j+1) is issued before P(j)V(j), so the tensor core computes the next logits tile while the softmax
// warps work on the current one; K/V stages are released by tcgen05.commit.
// Reference behaviour replaced: eager QK^T / softmax / PV with a materialised [B,Hq,T,L] logits tensor and repeat_kv copies
// (src/petals/models/llama/block.py:95-120).
#include "common.cuh"
#include "petals_b200.h"

namespace pb {

struct AttnTcParams {
  const __nv_bfloat16* q;
  const int* block_table;
  const int* pos_ptr;
  __nv_bfloat16* out;
  const float* alibi;
  float scale_log2;
  int B, T, Hq, Hkv, max_pages, num_pages, window, pos_static;
};

constexpr int kAtcThreads = 320;  // TMA warp + MMA warp + 8 softmax warps
constexpr int kAtcBM = 128, kAtcBN = 128;

PB_DEVICE float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int D>
__global__ void __launch_bounds__(kAtcThreads, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_k, const __grid_constant__ CUtensorMap tmap_v, const AttnTcParams p) {
  constexpr int NB = D / 64;                       // 64-element (128-byte) swizzle blocks along d
  constexpr int Q_BYTES = NB * kAtcBM * 128;       // [NB][128 rows][128 B]
  constexpr int P_BYTES = 2 * kAtcBM * 128;        // [2 key blocks][128 rows][128 B]
  constexpr int KV_BYTES = 2 * NB * 8192;          // per operand per stage: 2 pages x NB boxes of 8 KB
  constexpr int STAGE_BYTES = 2 * KV_BYTES;        // K then V
  constexpr int STAGES = 2;
  constexpr uint32_t TMEM_COLS = 512;              // S0 | S1 | O'  (128 + 128 + D <= 384 -> next power of two)

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sP = sQ + Q_BYTES;
  uint8_t* sKV = sP + P_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + STAGES * STAGE_BYTES);
  uint64_t* k_full = bars;             // [2]  TMA -> MMA
  uint64_t* k_empty = bars + 2;        // [2]  MMA (commit after S) -> TMA: the K stage is free as soon as the logits MMAs retire
  uint64_t* v_full = bars + 4;         // [2]  TMA -> MMA
  uint64_t* v_empty = bars + 6;        // [2]  MMA (commit after PV) -> TMA
  uint64_t* s_full = bars + 8;         // [2]  MMA -> softmax
  uint64_t* s_empty = bars + 10;       // [2]  softmax (8 warps) -> MMA
  uint64_t* p_full = bars + 12;        // [1]  softmax (8 warps) -> MMA
  uint64_t* o_full = bars + 13;        // [1]  MMA (commit after PV) -> softmax ; also means "P consumed"
  uint64_t* o_empty = bars + 14;       // [1]  softmax (8 warps) -> MMA
  uint64_t* q_full = bars + 15;        // [1]  softmax (8 warps) -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  float* sMax = reinterpret_cast<float*>(bars + 24);  // [2 parities][2 halves][128 rows] row-max exchange + [2][128] row sums

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // causal work grows with the tile index: schedule the long tiles first so the last wave is made of short ones
  const int lin = blockIdx.y * gridDim.x + blockIdx.x;  // dispatch order
  const int mt = gridDim.x - 1 - lin / static_cast<int>(gridDim.y), bh = lin % static_cast<int>(gridDim.y);
  const int b = bh / p.Hkv, kvh = bh - b * p.Hkv;
  const int G = p.Hq / p.Hkv;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_k);
    prefetch_tmap(&tmap_v);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], 8); }
    mbar_init(p_full, 8);
    mbar_init(o_full, 1);
    mbar_init(o_empty, 8);
    mbar_init(q_full, 8);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  pdl_trigger();
  pdl_wait();  // q and the appended K/V pages come from the RoPE kernel launched just before
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int pos0 = p.pos_ptr != nullptr ? *p.pos_ptr : p.pos_static;
  const int rows_total = p.T * G;
  const int m0 = mt * kAtcBM;
  const int t_min = m0 / G;
  const int t_max = min(p.T - 1, (m0 + kAtcBM - 1) / G);
  const int kv_hi = pos0 + t_max + 1;                       // keys [0, kv_hi) are visible to some row of this tile
  const int n_tiles = (kv_hi + kAtcBN - 1) / kAtcBN;
  int tile_lo = 0;
  if (p.window > 0) tile_lo = max(0, pos0 + t_min - p.window + 1) / kAtcBN;
  const int nt = n_tiles - tile_lo;                          // >= 1

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      const int* table = p.block_table + static_cast<size_t>(b) * p.max_pages;
      auto page_row = [&](int tile, int pg) {
        const int pidx = tile * 2 + pg;
        // the second page of the last tile may lie past the sequence's pages: fetch any valid page (its keys are masked)
        int page = pidx < p.max_pages ? table[pidx] : 0;
        if (page < 0 || page >= p.num_pages || pidx * 64 >= kv_hi) page = table[tile * 2];
        return (page * p.Hkv + kvh) * 64;
      };
      for (int j = 0; j < nt; ++j) {
        const int stage = j & 1;
        const uint32_t par = ((j >> 1) & 1) ^ 1;
        uint8_t* sk = sKV + stage * STAGE_BYTES;
        uint8_t* sv = sk + KV_BYTES;
        const int tile = tile_lo + j;
        mbar_wait(&k_empty[stage], par);
        mbar_expect_tx(&k_full[stage], KV_BYTES);
#pragma unroll
        for (int pg = 0; pg < 2; ++pg) {
          const int row = page_row(tile, pg);
#pragma unroll
          for (int db = 0; db < NB; ++db)
            tma_load_2d(sk + db * 16384 + pg * 8192, &tmap_k, &k_full[stage], db * 64, row);  // K: [d block][128 keys][128 B]
        }
        mbar_wait(&v_empty[stage], par);
        mbar_expect_tx(&v_full[stage], KV_BYTES);
#pragma unroll
        for (int pg = 0; pg < 2; ++pg) {
          const int row = page_row(tile, pg);
#pragma unroll
          for (int db = 0; db < NB; ++db)
            tma_load_2d(sv + (pg * NB + db) * 8192, &tmap_v, &v_full[stage], db * 64, row);     // V: [page][d block][64 keys][128 B]
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer =================================
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(kAtcBM, kAtcBN, 0, 0);  // A = Q (K-major), B = K (K-major)
      constexpr uint32_t idesc_o = umma_idesc_bf16(kAtcBM, D, 0, 1);       // A = P (K-major), B = V (MN-major)
      const uint32_t q_addr = smem_u32(sQ), p_addr = smem_u32(sP);
      const uint32_t tmem_o = tmem_base + 256;
      auto issue_s = [&](int j) {
        const int stage = j & 1, sb = j & 1;
        mbar_wait(&k_full[stage], (j >> 1) & 1);
        mbar_wait(&s_empty[sb], ((j >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(sKV + stage * STAGE_BYTES);
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          const uint64_t adesc = umma_desc_k_sw128(q_addr + (kk >> 2) * 16384 + (kk & 3) * 32);
          const uint64_t bdesc = umma_desc_k_sw128(k_addr + (kk >> 2) * 16384 + (kk & 3) * 32);
          tc_mma_f16(tmem_base + sb * 128, adesc, bdesc, idesc_s, kk != 0 ? 1u : 0u);
        }
        tc_commit(&s_full[sb]);
        tc_commit(&k_empty[stage]);
      };
      mbar_wait(q_full, 0);
      tc_fence_after();
      issue_s(0);
      for (int j = 0; j < nt; ++j) {
        if (j + 1 < nt) issue_s(j + 1);  // next logits tile runs under this tile's softmax
        const int stage = j & 1;
        mbar_wait(&v_full[stage], (j >> 1) & 1);
        mbar_wait(p_full, j & 1);
        mbar_wait(o_empty, (j & 1) ^ 1);
        tc_fence_after();
        const uint32_t v_addr = smem_u32(sKV + stage * STAGE_BYTES + KV_BYTES);
#pragma unroll
        for (int kk = 0; kk < kAtcBN / 16; ++kk) {
          const uint64_t adesc = umma_desc_k_sw128(p_addr + (kk >> 2) * 16384 + (kk & 3) * 32);
          const uint64_t bdesc = umma_desc_mn_sw128(v_addr + (kk >> 2) * NB * 8192 + (kk & 3) * 2048, 8192);
          tc_mma_f16(tmem_o, adesc, bdesc, idesc_o, kk != 0 ? 1u : 0u);
        }
        tc_commit(&v_empty[stage]);
        tc_commit(o_full);
      }
    }
  } else {
    // =============================== softmax / accumulate ========================
    // 8 warps: thread (row r, half h) owns keys [64h, 64h+64) of the logits tile and columns [h*D/2, (h+1)*D/2) of O.
    const int quarter = warp & 3;
    const int h = (warp - 2) >> 2;
    const int r = quarter * 32 + lane;           // tile row == TMEM lane
    const int row = m0 + r;
    const bool row_ok = row < rows_total;
    const int t = row_ok ? row / G : 0, g = row_ok ? row - (row / G) * G : 0;
    const int head = kvh * G + g;
    const int qpos = pos0 + t;
    const bool has_alibi = p.alibi != nullptr;
    const float slope2 = has_alibi ? p.alibi[head] * 1.4426950408889634f : 0.f;
    const float inv_scale = 1.f / p.scale_log2;
    constexpr int DH = D / 2;                    // O columns per thread

    // ---- Q: global -> swizzled shared (thread (r, h) copies half of row r) ----
    {
      const uint4* src = reinterpret_cast<const uint4*>(p.q + ((static_cast<size_t>(b) * p.T + t) * p.Hq + head) * D);
#pragma unroll
      for (int cc = 0; cc < D / 16; ++cc) {
        const int c = h * (D / 16) + cc;         // 16-byte chunk index within the row
        const uint4 v = row_ok ? __ldg(src + c) : make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(sQ + (c >> 3) * 16384 + r * 128 + (((c & 7) ^ (r & 7)) << 4)) = v;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(q_full);
    }

    float o[DH];
#pragma unroll
    for (int i = 0; i < DH; ++i) o[i] = 0.f;
    float m_i = -INFINITY, l_i = 0.f;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);

    auto fold_o = [&]() {
#pragma unroll
      for (int c = 0; c < DH; c += 16) {
        uint32_t rr[16];
        tmem_ld_32x16(lane_addr + 256 + h * DH + c, rr);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) o[c + i] += __uint_as_float(rr[i]);
      }
    };

    for (int j = 0; j < nt; ++j) {
      const int sb = j & 1;
      const int k0 = (tile_lo + j) * kAtcBN + h * 64;   // first key of this thread's half
      mbar_wait(&s_full[sb], (j >> 1) & 1);
      tc_fence_after();
      float sv[64];
#pragma unroll
      for (int c = 0; c < 64; c += 32) {
        uint32_t rr[32];
        tmem_ld_32x32(lane_addr + sb * 128 + h * 64 + c, rr);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) sv[c + i] = __uint_as_float(rr[i]);
      }
      // the logits now live in registers: hand the TMEM buffer back so S(j+2) can be issued
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[sb]);

      // scaled logits (exp2 domain) with position biases / masks; the common case (full tile, no ALiBi) is one FMUL
      const bool need_mask = (k0 + 63 > qpos) || (p.window > 0 && k0 < qpos - p.window + 1) || !row_ok;
      float m_loc = -INFINITY;
      if (!need_mask && !has_alibi) {
        // raw maximum with four independent chains; the scale is folded into the exponent FMA below
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int i = 0; i < 64; i += 4) {
          mx[0] = fmaxf(mx[0], sv[i]); mx[1] = fmaxf(mx[1], sv[i + 1]);
          mx[2] = fmaxf(mx[2], sv[i + 2]); mx[3] = fmaxf(mx[3], sv[i + 3]);
        }
        m_loc = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) * p.scale_log2;  // scale > 0
      } else {
#pragma unroll
        for (int i = 0; i < 64; ++i) {
          const int kpos = k0 + i;
          // stored back UNscaled-equivalent: (v / scale) so that the exponent FMA below is the same in both paths
          float v = sv[i] + (has_alibi ? slope2 * static_cast<float>(kpos - qpos) * inv_scale : 0.f);
          const bool ok = row_ok && kpos <= qpos && (p.window <= 0 || kpos > qpos - p.window);
          v = ok ? v : -INFINITY;
          sv[i] = v;
          m_loc = fmaxf(m_loc, v * p.scale_log2);
        }
      }
      // row maximum across the two halves
      sMax[((j & 1) * 2 + h) * 128 + r] = m_loc;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const float m_new = fmaxf(m_i, fmaxf(m_loc, sMax[((j & 1) * 2 + (h ^ 1)) * 128 + r]));
      const float m_safe = m_new == -INFINITY ? 0.f : m_new;
      const float alpha = fast_exp2(m_i - m_safe);  // m_i = -inf -> 0

      if (j > 0) {
        // O'(j-1) = P(j-1) V(j-1) was formed at the scale m_i; o_full also says P(j-1) has been consumed (sP reusable)
        mbar_wait(o_full, (j - 1) & 1);
        tc_fence_after();
        fold_o();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(o_empty);
      }
      if (__any_sync(0xffffffffu, alpha != 1.f)) {
#pragma unroll
        for (int i = 0; i < DH; ++i) o[i] *= alpha;
      }
      // probabilities -> bf16 -> row r of key block h of sP (128B swizzle), partial row sum
      float ls[4] = {0.f, 0.f, 0.f, 0.f};
      const float neg_m = -m_safe;
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8) {
        uint32_t pk[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float p0 = fast_exp2(fmaf(sv[c8 * 8 + 2 * e], p.scale_log2, neg_m));      // exp2(-inf) = 0 for masked keys
          const float p1 = fast_exp2(fmaf(sv[c8 * 8 + 2 * e + 1], p.scale_log2, neg_m));
          pk[e] = pack_bf16(p0, p1);
          ls[e] += p0 + p1;
        }
        *reinterpret_cast<uint4*>(sP + h * 16384 + r * 128 + ((c8 ^ (r & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
      l_i = l_i * alpha + ((ls[0] + ls[1]) + (ls[2] + ls[3]));
      m_i = m_new;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    // last O'
    mbar_wait(o_full, (nt - 1) & 1);
    tc_fence_after();
    fold_o();
    tc_fence_before();
    // total row sum = both halves
    sMax[(4 + h) * 128 + r] = l_i;  // dedicated slots: the partner may still be reading the last row-max exchange
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const float l_tot = l_i + sMax[(4 + (h ^ 1)) * 128 + r];
    if (row_ok) {
      const float inv = l_tot > 0.f ? 1.f / l_tot : 0.f;
      uint4* dst = reinterpret_cast<uint4*>(p.out + ((static_cast<size_t>(b) * p.T + t) * p.Hq + head) * D + h * DH);
#pragma unroll
      for (int c = 0; c < DH / 8; ++c) {
        uint4 v;
        v.x = pack_bf16(o[8 * c] * inv, o[8 * c + 1] * inv);
        v.y = pack_bf16(o[8 * c + 2] * inv, o[8 * c + 3] * inv);
        v.z = pack_bf16(o[8 * c + 4] * inv, o[8 * c + 5] * inv);
        v.w = pack_bf16(o[8 * c + 6] * inv, o[8 * c + 7] * inv);
        dst[c] = v;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

bool make_tmap_2d_bf16(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows, uint32_t box_cols);

template <int D>
static int launch_attn_tc(const PbAttnArgs* a, cudaStream_t s) {
  constexpr int NB = D / 64;
  const size_t smem = static_cast<size_t>(NB * 128 * 128 + 2 * 128 * 128 + 2 * 2 * (2 * NB * 8192)) + 1024 + 192 + 6 * 128 * 4 + 64;
  const uint64_t rows = static_cast<uint64_t>(a->num_pages) * a->Hkv * 64;
  CUtensorMap tk, tv;
  if (!make_tmap_2d_bf16(&tk, a->k_pool, rows, D, D, 64, 64)) return PB_ERR_DRIVER;
  if (!make_tmap_2d_bf16(&tv, a->v_pool, rows, D, D, 64, 64)) return PB_ERR_DRIVER;
  AttnTcParams p{};
  p.q = static_cast<const __nv_bfloat16*>(a->q);
  p.block_table = static_cast<const int*>(a->block_table);
  p.pos_ptr = static_cast<const int*>(a->pos_ptr);
  p.out = static_cast<__nv_bfloat16*>(a->out);
  p.alibi = static_cast<const float*>(a->alibi_slopes);
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.B = a->B; p.T = a->T; p.Hq = a->Hq; p.Hkv = a->Hkv; p.max_pages = a->max_pages; p.num_pages = a->num_pages;
  p.window = a->window; p.pos_static = a->pos_static;
  auto kern = attn_fwd_tc_kernel<D>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess) return PB_ERR_CUDA;
    attr_set = true;
  }
  const int G = a->Hq / a->Hkv;
  const int m_tiles = (a->T * G + kAtcBM - 1) / kAtcBM;
  dim3 grid(m_tiles, a->B * a->Hkv);
  launch_pdl(kPdlAttn, kern, grid, dim3(kAtcThreads), smem, s, tk, tv, p);
  return pb_check_launch("attention_tc");
}

int attention_tc_dispatch(const PbAttnArgs* a, cudaStream_t s) {
  if (a->D == 128) return launch_attn_tc<128>(a, s);
  if (a->D == 64) return launch_attn_tc<64>(a, s);
  return PB_ERR_SHAPE;
}

}  // namespace pb
