// petals_b200 — block-scaled FP8 GEMM on the 5th-generation tensor cores (sm_100a).
//
//   D[M, N] = epilogue( (A_q · 2^SFA) [M, K] · (B_q · 2^SFB)[N, K]^T )
//
// Both operands are MXFP8 (ops/quant.py): E4M3 payload, one UE8M0 power-of-two scale per 32 values along K. The tensor
// cores consume exactly that: `tcgen05.mma.kind::mxf8f6f4.block_scale` multiplies 128 x 256 x 32 tiles at twice the bf16
// rate and applies the per-32 scales of BOTH operands in hardware, reading them from tensor memory. SASS: UTCQMMA (the
// MMA), UTCCP (scale factors smem -> TMEM), UTMALDG (operand tiles), UBLKCP (scale-factor blocks).
//
// Same persistent warp-specialised shape as gemm_tcgen05.cu:
//   warp 0     : producer — per K block (128 values = 128 B rows, 128B swizzle): TMA tiles of A [128 x 128] and
//                B [256 x 128] plus the matching scale-factor blocks (512 B per 128 rows, 1-D bulk copies)
//   warp 1     : MMA issuer — copies the scale factors of the stage into TMEM (tcgen05.cp 32x128b.warpx4: 4 columns per
//                128 rows), then issues four K=32 MMAs whose instruction descriptors select the scale byte of their K
//                slice. tcgen05.cp and tcgen05.mma execute in issue order, so the same 12 scale columns serve every stage.
//   warps 2..5 : epilogue — TMEM -> registers, fused SwiGLU (B tile = 128 gate rows + 128 up rows) / residual, bf16 stores
// TMEM budget: one 128 x 256 fp32 accumulator (256 columns) + 12 scale columns; a second accumulator does not fit next to
// them, so the epilogue is not overlapped with the next tile's main loop (K >= 4096: < 10 % of a tile's time).
//
// Scale-factor layout in global memory (what `tcgen05.cp` wants to find in shared memory, so a block is one bulk copy):
//   [K / 128][ceil(rows / 128)][512 B]; inside a block the scale of (row r, K-slice c) sits at (r % 32) * 16 + (r / 32) * 4 + c.
// Written by quant_mxfp8_kernel for activations (fused with the RMSNorm that precedes the projection) and by
// ops/quant.py:pack_scales for weights.
//
// Replaces the bitsandbytes INT8 matmul of the reference's quantised serving (src/petals/utils/convert_block.py:76-115).
#include "common.cuh"
#include "petals_b200.h"

#include <cuda_fp8.h>

#include <atomic>
#include <mutex>

extern "C" int pb_set_error(const char* msg);

namespace pb {

bool make_tmap_2d_u8(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows, uint32_t box_cols);
bool make_tmap_sf_blocks(CUtensorMap* m, const void* base, uint64_t blocks, uint32_t box_rows);

namespace f8 {

constexpr int BM = 128;
constexpr int BN = 256;
constexpr int BK = 128;            // values (= bytes) per K block: one 128B-swizzle row
constexpr int kThreads = 192;
constexpr int kGroupM = 8;
constexpr int SF_BLOCK = 512;      // bytes of scale factors per 128 rows x 128 K values
constexpr int A_BYTES = BM * BK, B_BYTES = BN * BK, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int SF_STAGE = 3 * SF_BLOCK;   // A block + two B blocks
constexpr uint32_t ACC_COLS = BN, SFA_COL = ACC_COLS, SFB_COL = ACC_COLS + 4, TMEM_COLS = 512;

struct Params {
  const uint8_t* sfa;
  const uint8_t* sfb;
  const uint8_t* sfb2;
  const __nv_bfloat16* residual;
  __nv_bfloat16* out;
  int M, N, K;
  int ldo, ldres;
  int sfa_blocks, sfb_blocks;   // 128-row blocks per K block in the scale arrays
};

// kind::mxf8f6f4 block-scaled instruction descriptor: E4M3 x E4M3 -> fp32, UE8M0 scales, K-major operands.
__host__ __device__ constexpr uint32_t idesc_mxf8(uint32_t M, uint32_t N, uint32_t a_sf_id, uint32_t b_sf_id) {
  return (b_sf_id << 4) | (0u << 7) | (0u << 10) | ((N >> 3) << 17) | (1u << 23) | ((M >> 4) << 24) | (a_sf_id << 29);
}

// Shared-memory descriptor of one scale-factor block for tcgen05.cp: 32 rows x 16 B, no swizzle, 8-row groups 128 B apart.
PB_DEVICE uint64_t sf_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3ffffu) >> 4);
  d |= static_cast<uint64_t>(128 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}

PB_DEVICE void tc_cp_sf(uint32_t tmem_dst, uint64_t desc) {
  asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(tmem_dst), "l"(desc) : "memory");
}

PB_DEVICE void tc_mma_mxf8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate, uint32_t tmem_sfa,
                           uint32_t tmem_sfb) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::mxf8f6f4.block_scale [%0], %1, %2, %3, [%5], [%6], p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(tmem_sfa), "r"(tmem_sfb)
      : "memory");
}

PB_DEVICE float silu_f(float x) { return x / (1.f + __expf(-x)); }
PB_DEVICE float round_bf16(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

PB_DEVICE void tile_coords(int tile, int m_blocks, int n_blocks, int& m_blk, int& n_blk) {
  const int per_group = kGroupM * n_blocks;
  const int g = tile / per_group;
  const int first_m = g * kGroupM;
  const int gsize = min(kGroupM, m_blocks - first_m);
  const int r = tile - g * per_group;
  m_blk = first_m + r % gsize;
  n_blk = r / gsize;
}

// One warp's share of a finished accumulator tile: 32 rows (this lane = one row) x OUT_BN columns, 32 columns at a time.
template <bool DUAL>
PB_DEVICE void epilogue_rows(const Params& p, uint32_t taddr, int row, bool row_ok, int n_blk) {
  constexpr int OUT_BN = DUAL ? BN / 2 : BN;
#pragma unroll 1
  for (int c = 0; c < OUT_BN; c += 32) {
    const int col0 = n_blk * OUT_BN + c;
    if (col0 >= p.N) break;  // warp-uniform
    uint32_t r[32];
    float v[32];
    tmem_ld_32x32(taddr + c, r);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
    if (DUAL) {
      uint32_t r2[32];
      tmem_ld_32x32(taddr + OUT_BN + c, r2);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = round_bf16(silu_f(round_bf16(v[i]))) * round_bf16(__uint_as_float(r2[i]));
    }
    if (row_ok) {
      const bool full = col0 + 32 <= p.N;
      if (p.residual != nullptr) {
        const __nv_bfloat16* rp = p.residual + static_cast<size_t>(row) * p.ldres + col0;
        if (full) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint4 rv = *reinterpret_cast<const uint4*>(rp + q * 8);
            const uint32_t w[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              v[q * 8 + 2 * j] = round_bf16(v[q * 8 + 2 * j]) + bf16_lo(w[j]);
              v[q * 8 + 2 * j + 1] = round_bf16(v[q * 8 + 2 * j + 1]) + bf16_hi(w[j]);
            }
          }
        } else {
          for (int i = 0; i < 32 && col0 + i < p.N; ++i) v[i] = round_bf16(v[i]) + __bfloat162float(rp[i]);
        }
      }
      __nv_bfloat16* op = p.out + static_cast<size_t>(row) * p.ldo + col0;
      if (full) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 pk;
          pk.x = pack_bf16(v[q * 8 + 0], v[q * 8 + 1]);
          pk.y = pack_bf16(v[q * 8 + 2], v[q * 8 + 3]);
          pk.z = pack_bf16(v[q * 8 + 4], v[q * 8 + 5]);
          pk.w = pack_bf16(v[q * 8 + 6], v[q * 8 + 7]);
          reinterpret_cast<uint4*>(op)[q] = pk;
        }
      } else {
        for (int i = 0; i < 32 && col0 + i < p.N; ++i) op[i] = __float2bfloat16_rn(v[i]);
      }
    }
  }
}

// DUAL: the B tile is 128 rows of `b` (gate) + 128 rows of `b2` (up); the epilogue emits silu(gate) * up, 128 columns per tile.
template <bool DUAL, int STAGES, bool SFPIPE>
__global__ void __launch_bounds__(kThreads, 1)
gemm_mxfp8_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const __grid_constant__ CUtensorMap tmap_b2, const Params p) {
  constexpr int OUT_BN = DUAL ? BN / 2 : BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sf_smem = smem + STAGES * STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sf_smem + STAGES * SF_STAGE);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_blocks = (p.M + BM - 1) / BM;
  const int n_blocks = (p.N + OUT_BN - 1) / OUT_BN;
  const int num_tiles = m_blocks * n_blocks;
  const int num_kb = p.K / BK;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    if (DUAL) prefetch_tmap(&tmap_b2);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, 4);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =============================== producer ===============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int m_blk, n_blk;
        tile_coords(tile, m_blocks, n_blocks, m_blk, n_blk);
        // scale-factor blocks of this tile's B rows: DUAL -> block n_blk of each tensor; else blocks 2 n_blk (and 2 n_blk + 1 if it exists)
        const int nb0 = DUAL ? n_blk : 2 * n_blk;
        const bool second = DUAL || (nb0 + 1 < p.sfb_blocks);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          uint8_t* ss = sf_smem + stage * SF_STAGE;
          mbar_expect_tx(&full_bar[stage], STAGE_BYTES + (second ? 3 : 2) * SF_BLOCK);
          tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * BK, m_blk * BM);
          if (!DUAL) {
            tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * BK, n_blk * BN);
          } else {
            tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * BK, n_blk * OUT_BN);
            tma_load_2d(sb + B_BYTES / 2, &tmap_b2, &full_bar[stage], kb * BK, n_blk * OUT_BN);
          }
          bulk_load_1d(ss, p.sfa + (static_cast<size_t>(kb) * p.sfa_blocks + m_blk) * SF_BLOCK, SF_BLOCK, &full_bar[stage]);
          const size_t boff = (static_cast<size_t>(kb) * p.sfb_blocks + nb0) * SF_BLOCK;
          bulk_load_1d(ss + SF_BLOCK, p.sfb + boff, SF_BLOCK, &full_bar[stage]);
          if (second) bulk_load_1d(ss + 2 * SF_BLOCK, DUAL ? p.sfb2 + boff : p.sfb + boff + SF_BLOCK, SF_BLOCK, &full_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer =================================
    // The K blocks of all this CTA's tiles form one stream g = 0, 1, 2, ...: ring stage g % STAGES, scale columns of parity g & 1.
    // SFPIPE: the scale factors of block g + 1 are copied into the OTHER column set before the MMAs of block g are issued, so the
    // copy can overlap those MMAs instead of sitting between two MMA groups.
    if (lane == 0) {
      const int my_tiles = blockIdx.x < num_tiles ? (num_tiles - 1 - static_cast<int>(blockIdx.x)) / static_cast<int>(gridDim.x) + 1 : 0;
      const long total = static_cast<long>(my_tiles) * num_kb;
      auto sf_cols = [&](long g) { return tmem_base + SFA_COL + (SFPIPE ? static_cast<uint32_t>(g & 1) * 16u : 0u); };
      auto copy_scales = [&](long g) {
        const int s_ = static_cast<int>(g % STAGES);
        mbar_wait(&full_bar[s_], static_cast<uint32_t>((g / STAGES) & 1));
        tc_fence_after();
        const uint32_t ss = smem_u32(sf_smem + s_ * SF_STAGE);
        const uint32_t c = sf_cols(g);
        tc_cp_sf(c, sf_desc(ss));
        tc_cp_sf(c + 4, sf_desc(ss + SF_BLOCK));
        tc_cp_sf(c + 8, sf_desc(ss + 2 * SF_BLOCK));
      };
      long g = 0;
      if (SFPIPE && total > 0) copy_scales(0);
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        mbar_wait(tmem_empty, (it & 1) ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < num_kb; ++kb, ++g) {
          if (SFPIPE) {
            if (g + 1 < total) copy_scales(g + 1);
          } else {
            copy_scales(g);
          }
          const int stage = static_cast<int>(g % STAGES);
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t sb = sa + A_BYTES;
          const uint32_t c = sf_cols(g);
#pragma unroll
          for (int k = 0; k < BK / 32; ++k) {
            const uint64_t adesc = umma_desc_k_sw128(sa + k * 32);
            const uint64_t bdesc = umma_desc_k_sw128(sb + k * 32);
            tc_mma_mxf8(tmem_base, adesc, bdesc, idesc_mxf8(BM, BN, k, k), (kb | k) != 0 ? 1u : 0u, c, c + 4);
          }
          tc_commit(&empty_bar[stage]);
        }
        tc_commit(tmem_full);
      }
    }
  } else {
    // =============================== epilogue ====================================
    const int quarter = warp & 3;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      int m_blk, n_blk;
      tile_coords(tile, m_blocks, n_blocks, m_blk, n_blk);
      mbar_wait(tmem_full, it & 1);
      tc_fence_after();
      const int row = m_blk * BM + quarter * 32 + lane;
      const bool row_ok = row < p.M;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
      epilogue_rows<DUAL>(p, taddr, row, row_ok, n_blk);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---- 2-CTA variant: one 256 x 256 tile per SM pair (tcgen05.mma.cta_group::2.kind::mxf8f6f4.block_scale) -----------------------------
// Same idea as gemm_tcgen05_2cta.cu: each CTA stages its own 128 rows of A and HALF of the B tile, so the operand stream per SM drops
// from 48 KB to 32 KB per K block — this kernel is L2 -> SM bound at the FP8 math rate. Scale factors: each SM's tensor core needs the
// scales of its own 128 A rows and of ALL 256 B rows, so both CTAs load the same two B blocks; `tcgen05.cp.cta_group::2` copies each
// CTA's shared-memory blocks into that CTA's tensor memory. All loads are tensor-map TMA with `.cta_group::2` (completion bytes counted on
// the leader's barrier); the scale array is addressed as [blocks, 128] 32-bit words.
namespace pair {

constexpr int STAGES = 6;
constexpr int A2_BYTES = BM * BK, B2_BYTES = (BN / 2) * BK, STAGE2_BYTES = A2_BYTES + B2_BYTES;   // 16 KB + 16 KB
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;

PB_DEVICE uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
PB_DEVICE void cluster_sync() { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
PB_DEVICE void tma_pair(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
PB_DEVICE void tc_cp_sf_pair(uint32_t tmem_dst, uint64_t desc) {
  asm volatile("tcgen05.cp.cta_group::2.32x128b.warpx4 [%0], %1;" ::"r"(tmem_dst), "l"(desc) : "memory");
}
PB_DEVICE void tc_mma_mxf8_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate, uint32_t tmem_sfa, uint32_t tmem_sfb) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::mxf8f6f4.block_scale [%0], %1, %2, %3, [%5], [%6], p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(tmem_sfa), "r"(tmem_sfb)
      : "memory");
}
PB_DEVICE void tc_commit_pair(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
PB_DEVICE void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask) : "memory");
}

// groups of 4 tile pairs (1024 rows) along M, N fastest inside a group: the ~74 concurrent pairs share 4 row blocks of A and ~18 weight
// tiles, whatever K is (an M-fastest order would re-stream all of A every wave once A no longer fits in L2)
PB_DEVICE void pair_coords(int tile, int m_pairs, int n_blocks, int& m_pair, int& n_blk) {
  constexpr int kGroup = 4;
  const int per_group = kGroup * n_blocks;
  const int g = tile / per_group;
  const int first_m = g * kGroup;
  const int gsize = min(kGroup, m_pairs - first_m);
  const int r = tile - g * per_group;
  m_pair = first_m + r % gsize;
  n_blk = r / gsize;
}

template <bool DUAL>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm_mxfp8_2cta_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const __grid_constant__ CUtensorMap tmap_b2,
                       const __grid_constant__ CUtensorMap tmap_sfa, const __grid_constant__ CUtensorMap tmap_sfb, const __grid_constant__ CUtensorMap tmap_sfb2,
                       const Params p) {
  constexpr int OUT_BN = DUAL ? BN / 2 : BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sf_smem = smem + STAGES * STAGE2_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sf_smem + STAGES * SF_STAGE);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta = cluster_ctarank();
  const bool leader = cta == 0;
  const int m_pairs = (p.M + 2 * BM - 1) / (2 * BM);
  const int n_blocks = (p.N + OUT_BN - 1) / OUT_BN;
  const int num_tiles = m_pairs * n_blocks;
  const int num_kb = p.K / BK;
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_a); prefetch_tmap(&tmap_b); prefetch_tmap(&tmap_sfa); prefetch_tmap(&tmap_sfb);
    if (DUAL) { prefetch_tmap(&tmap_b2); prefetch_tmap(&tmap_sfb2); }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, 8);   // 4 epilogue warps x 2 CTAs, on the leader's copy
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
        int m_pair, n_blk;
        pair_coords(tile, m_pairs, n_blocks, m_pair, n_blk);
        const int m_blk = m_pair * 2 + static_cast<int>(cta);        // this CTA's 128-row block of A
        const int nb0 = DUAL ? n_blk : 2 * n_blk;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE2_BYTES;
          uint8_t* sb = sa + A2_BYTES;
          uint8_t* ss = sf_smem + stage * SF_STAGE;
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * (STAGE2_BYTES + SF_STAGE));
          tma_pair(sa, &tmap_a, &full_bar[stage], kb * BK, m_blk * BM);
          if (!DUAL) {
            tma_pair(sb, &tmap_b, &full_bar[stage], kb * BK, n_blk * BN + static_cast<int>(cta) * (BN / 2));
          } else {
            tma_pair(sb, leader ? &tmap_b : &tmap_b2, &full_bar[stage], kb * BK, n_blk * OUT_BN);
          }
          tma_pair(ss, &tmap_sfa, &full_bar[stage], 0, kb * p.sfa_blocks + m_blk);
          if (!DUAL) {
            tma_pair(ss + SF_BLOCK, &tmap_sfb, &full_bar[stage], 0, kb * p.sfb_blocks + nb0);   // box of 2 blocks (the 2nd is zero-filled / unused past the end)
          } else {
            tma_pair(ss + SF_BLOCK, &tmap_sfb, &full_bar[stage], 0, kb * p.sfb_blocks + nb0);
            tma_pair(ss + 2 * SF_BLOCK, &tmap_sfb2, &full_bar[stage], 0, kb * p.sfb_blocks + nb0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (leader && lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += n_clusters, ++it) {
        mbar_wait(tmem_empty, (it & 1) ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * STAGE2_BYTES);
          const uint32_t sb = sa + A2_BYTES;
          const uint32_t ss = smem_u32(sf_smem + stage * SF_STAGE);
          tc_cp_sf_pair(tmem_base + SFA_COL, sf_desc(ss));
          tc_cp_sf_pair(tmem_base + SFB_COL, sf_desc(ss + SF_BLOCK));
          tc_cp_sf_pair(tmem_base + SFB_COL + 4, sf_desc(ss + 2 * SF_BLOCK));
#pragma unroll
          for (int k = 0; k < BK / 32; ++k)
            tc_mma_mxf8_pair(tmem_base, umma_desc_k_sw128(sa + k * 32), umma_desc_k_sw128(sb + k * 32), idesc_mxf8(2 * BM, BN, k, k), (kb | k) != 0 ? 1u : 0u,
                             tmem_base + SFA_COL, tmem_base + SFB_COL);
          tc_commit_pair(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit_pair(tmem_full);
      }
    }
  } else {
    const int quarter = warp & 3;
    int it = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += n_clusters, ++it) {
      int m_pair, n_blk;
      pair_coords(tile, m_pairs, n_blocks, m_pair, n_blk);
      mbar_wait(tmem_full, it & 1);
      tc_fence_after();
      const int row = (m_pair * 2 + static_cast<int>(cta)) * BM + quarter * 32 + lane;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
      epilogue_rows<DUAL>(p, taddr, row, row < p.M, n_blk);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(tmem_empty);
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

}  // namespace pair

// ---- activation quantisation (optionally fused with the RMSNorm in front of the projection) ---------------------------
// One warp per row. Pass 1 (only with a norm weight): sum of squares. Pass 2: 8 values per lane, 4 lanes per 32-value group:
// group maximum -> UE8M0 exponent (the smallest power of two that brings the group into E4M3 range) -> E4M3 payload.
__global__ void __launch_bounds__(256) quant_mxfp8_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ norm_w, float eps,
                                                          uint8_t* __restrict__ q, uint8_t* __restrict__ sf, int M, int K, int sf_blocks) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  if (row >= M) return;
  const __nv_bfloat16* xr = x + static_cast<size_t>(row) * K;
  float inv = 1.f;
  if (norm_w != nullptr) {
    float ss = 0.f;
    for (int i = lane * 8; i < K; i += 256) {
      const uint4 v = *reinterpret_cast<const uint4*>(xr + i);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) { const float a = bf16_lo(w[j]), b = bf16_hi(w[j]); ss += a * a + b * b; }
    }
    ss = warp_sum(ss);
    inv = rsqrtf(ss / static_cast<float>(K) + eps);
  }
  const int m_blk = row >> 7, r_in = row & 127;
  uint8_t* sf_row = sf + static_cast<size_t>(m_blk) * SF_BLOCK + (r_in & 31) * 16 + (r_in >> 5) * 4;
  for (int i = lane * 8; i < K; i += 256) {   // K % 128 == 0, so every lane of a 4-lane group is in range together
    const uint4 v = *reinterpret_cast<const uint4*>(xr + i);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    float f[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) { f[2 * j] = bf16_lo(w[j]); f[2 * j + 1] = bf16_hi(w[j]); }
    if (norm_w != nullptr) {
      const uint4 g = *reinterpret_cast<const uint4*>(norm_w + i);
      const uint32_t gw[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {   // HF rounding: bf16(x * inv) * w, rounded to bf16 again (what the bf16 path feeds its GEMM)
        f[2 * j] = round_bf16(round_bf16(f[2 * j] * inv) * bf16_lo(gw[j]));
        f[2 * j + 1] = round_bf16(round_bf16(f[2 * j + 1] * inv) * bf16_hi(gw[j]));
      }
    }
    float amax = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) amax = fmaxf(amax, fabsf(f[j]));
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2));
    int e = -127;
    if (amax > 0.f) {
      int ex;
      const float mant = frexpf(amax * (1.f / 448.f), &ex);   // amax / 448 = mant * 2^ex, mant in [0.5, 1)
      e = (mant == 0.5f) ? ex - 1 : ex;                        // ceil(log2(amax / 448))
      e = max(-127, min(127, e));
    }
    const float s = exp2f(static_cast<float>(-e));
    uint32_t lo = 0, hi = 0;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      lo |= static_cast<uint32_t>(__nv_cvt_float2_to_fp8x2(make_float2(f[2 * j] * s, f[2 * j + 1] * s), __NV_SATFINITE, __NV_E4M3)) << (16 * j);
      hi |= static_cast<uint32_t>(__nv_cvt_float2_to_fp8x2(make_float2(f[4 + 2 * j] * s, f[5 + 2 * j] * s), __NV_SATFINITE, __NV_E4M3)) << (16 * j);
    }
    *reinterpret_cast<uint2*>(q + static_cast<size_t>(row) * K + i) = make_uint2(lo, hi);
    if ((lane & 3) == 0) {
      const int g32 = i >> 5;   // 32-value group index along K
      sf_row[static_cast<size_t>(g32 >> 2) * sf_blocks * SF_BLOCK + (g32 & 3)] = static_cast<uint8_t>(e + 127);
    }
  }
}

template <bool DUAL>
static int launch(const PbGemmFp8Args* a, cudaStream_t stream) {
  constexpr int STAGES = 4;
  constexpr int OUT_BN = DUAL ? BN / 2 : BN;
  CUtensorMap ta, tb, tb2;
  if (!make_tmap_2d_u8(&ta, a->a_q, a->M, a->K, a->K, BM, BK)) return PB_ERR_CUDA;
  const uint32_t b_box = DUAL ? BN / 2 : BN;
  if (!make_tmap_2d_u8(&tb, a->b_q, a->N, a->K, a->K, b_box, BK)) return PB_ERR_CUDA;
  tb2 = tb;
  if (DUAL && !make_tmap_2d_u8(&tb2, a->b2_q, a->N, a->K, a->K, b_box, BK)) return PB_ERR_CUDA;
  Params p{};
  p.sfa = static_cast<const uint8_t*>(a->a_sf); p.sfb = static_cast<const uint8_t*>(a->b_sf); p.sfb2 = static_cast<const uint8_t*>(a->b2_sf);
  p.residual = static_cast<const __nv_bfloat16*>(a->residual); p.out = static_cast<__nv_bfloat16*>(a->out);
  p.M = a->M; p.N = a->N; p.K = a->K; p.ldo = a->ldo > 0 ? a->ldo : a->N; p.ldres = a->ldres > 0 ? a->ldres : a->N;
  p.sfa_blocks = (a->M + 127) / 128; p.sfb_blocks = (a->N + 127) / 128;
  const int smem = STAGES * (STAGE_BYTES + SF_STAGE) + 1024 + 256;
  // measured (profiles/r2_fp8_sfpipe_mn_group8.txt): copying one block ahead is 7-14 % SLOWER than copying right before the MMAs -> off
  static const bool sfpipe = [] { const char* e = getenv("PETALS_B200_FP8_SFPIPE"); return e != nullptr && atoi(e) != 0; }();
  auto kern = sfpipe ? gemm_mxfp8_kernel<DUAL, STAGES, true> : gemm_mxfp8_kernel<DUAL, STAGES, false>;
  static std::atomic<bool> attr_done[64];
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_done[dev & 63].load()) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return pb_check_launch("gemm_mxfp8 attr");
    attr_done[dev & 63].store(true);
  }
  const int tiles = ((a->M + BM - 1) / BM) * ((a->N + OUT_BN - 1) / OUT_BN);
  const int sms = a->num_sms > 0 ? a->num_sms : 148;
  const int grid = tiles < sms ? tiles : sms;
  kern<<<grid, kThreads, smem, stream>>>(ta, tb, tb2, p);
  return pb_check_launch("gemm_mxfp8");
}

template <bool DUAL>
static int launch_pair(const PbGemmFp8Args* a, cudaStream_t stream) {
  constexpr int OUT_BN = DUAL ? BN / 2 : BN;
  CUtensorMap ta, tb, tb2, tsa, tsb, tsb2;
  if (!make_tmap_2d_u8(&ta, a->a_q, a->M, a->K, a->K, BM, BK)) return PB_ERR_CUDA;
  if (!make_tmap_2d_u8(&tb, a->b_q, a->N, a->K, a->K, BN / 2, BK)) return PB_ERR_CUDA;
  tb2 = tb;
  if (DUAL && !make_tmap_2d_u8(&tb2, a->b2_q, a->N, a->K, a->K, BN / 2, BK)) return PB_ERR_CUDA;
  const int kb = a->K / BK, sfa_blocks = (a->M + 127) / 128, sfb_blocks = (a->N + 127) / 128;
  if (!make_tmap_sf_blocks(&tsa, a->a_sf, static_cast<uint64_t>(kb) * sfa_blocks, 1)) return PB_ERR_CUDA;
  if (!make_tmap_sf_blocks(&tsb, a->b_sf, static_cast<uint64_t>(kb) * sfb_blocks, DUAL ? 1 : 2)) return PB_ERR_CUDA;
  tsb2 = tsb;
  if (DUAL && !make_tmap_sf_blocks(&tsb2, a->b2_sf, static_cast<uint64_t>(kb) * sfb_blocks, 1)) return PB_ERR_CUDA;
  Params p{};
  p.residual = static_cast<const __nv_bfloat16*>(a->residual); p.out = static_cast<__nv_bfloat16*>(a->out);
  p.M = a->M; p.N = a->N; p.K = a->K; p.ldo = a->ldo > 0 ? a->ldo : a->N; p.ldres = a->ldres > 0 ? a->ldres : a->N;
  p.sfa_blocks = sfa_blocks; p.sfb_blocks = sfb_blocks;
  const int smem = pair::STAGES * (pair::STAGE2_BYTES + SF_STAGE) + 1024 + 256;
  auto kern = pair::gemm_mxfp8_2cta_kernel<DUAL>;
  static std::atomic<int> max_clusters[64];   // 0 = not asked yet
  int dev = 0;
  cudaGetDevice(&dev);
  const int sms = a->num_sms > 0 ? a->num_sms : 148;
  if (max_clusters[dev & 63].load() == 0) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return pb_check_launch("gemm_mxfp8_2cta attr");
    cudaLaunchConfig_t cfg = {};   // persistent grid = the co-resident pairs, not more
    cfg.gridDim = dim3(sms / 2 * 2); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) { cudaGetLastError(); n = sms / 2; }
    max_clusters[dev & 63].store(n < sms / 2 ? n : sms / 2);
  }
  const int tiles = ((a->M + 2 * BM - 1) / (2 * BM)) * ((a->N + OUT_BN - 1) / OUT_BN);
  int clusters = max_clusters[dev & 63].load();
  if (tiles < clusters) clusters = tiles;
  kern<<<2 * clusters, kThreads, smem, stream>>>(ta, tb, tb2, tsa, tsb, tsb2, p);
  return pb_check_launch("gemm_mxfp8_2cta");
}

}  // namespace f8
}  // namespace pb

using namespace pb;

extern "C" int pb_gemm_mxfp8(const PbGemmFp8Args* a, void* stream) {
  if (a == nullptr || a->M <= 0 || a->N <= 0 || a->K <= 0 || a->K % f8::BK != 0) { pb_set_error("gemm_mxfp8: K must be a positive multiple of 128"); return PB_ERR_SHAPE; }
  if (a->a_q == nullptr || a->a_sf == nullptr || a->b_q == nullptr || a->b_sf == nullptr || a->out == nullptr) return PB_ERR_SHAPE;
  if (a->act == 1) {
    if (a->b2_q == nullptr || a->b2_sf == nullptr) return PB_ERR_SHAPE;
    return f8::launch<true>(a, static_cast<cudaStream_t>(stream));
  }
  if (a->act != 0) { pb_set_error("gemm_mxfp8: only act 0 (none) and 1 (SwiGLU) are fused"); return PB_ERR_UNSUPPORTED; }
  return f8::launch<false>(a, static_cast<cudaStream_t>(stream));
}

// 2-CTA variant of pb_gemm_mxfp8 (same arguments): one 256 x 256 tile per SM pair.
extern "C" int pb_gemm_mxfp8_2cta(const PbGemmFp8Args* a, void* stream) {
  if (a == nullptr || a->M <= 0 || a->N <= 0 || a->K <= 0 || a->K % f8::BK != 0) { pb_set_error("gemm_mxfp8_2cta: K must be a positive multiple of 128"); return PB_ERR_SHAPE; }
  if (a->a_q == nullptr || a->a_sf == nullptr || a->b_q == nullptr || a->b_sf == nullptr || a->out == nullptr) return PB_ERR_SHAPE;
  if (a->act == 1) {
    if (a->b2_q == nullptr || a->b2_sf == nullptr) return PB_ERR_SHAPE;
    return f8::launch_pair<true>(a, static_cast<cudaStream_t>(stream));
  }
  if (a->act != 0) return PB_ERR_UNSUPPORTED;
  return f8::launch_pair<false>(a, static_cast<cudaStream_t>(stream));
}

extern "C" int pb_quant_mxfp8(const void* x, const void* norm_w, float eps, void* q, void* sf, int M, int K, void* stream) {
  if (M <= 0 || K <= 0 || K % 128 != 0) { pb_set_error("quant_mxfp8: K must be a positive multiple of 128"); return PB_ERR_SHAPE; }
  f8::quant_mxfp8_kernel<<<(M + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(norm_w), eps, static_cast<uint8_t*>(q), static_cast<uint8_t*>(sf), M, K,
      (M + 127) / 128);
  return pb_check_launch("quant_mxfp8");
}
