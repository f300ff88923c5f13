// petals_b200 — 2-CTA tcgen05 GEMM for the large prefill / training shapes (sm_100a).
//
//   D[M, N] = epilogue( A[M, K] · B[N, K]^T )       bf16 operands, fp32 accumulation in TMEM
//
// The 1-CTA kernel (gemm_tcgen05.cu) streams 48 KB of operands per 128 x 256 x 64 step into every SM: 96 B per clock per SM at the
// full tensor rate, which is what the L2 -> SM path delivers and no more (ncu: tensor pipe 84 % active, DRAM at 20 %). Here two CTAs
// of a cluster (an SM pair) compute ONE 256 x 256 tile with `tcgen05.mma.cta_group::2`: each CTA stages its own 128 rows of A and
// only HALF of the B tile (128 of the 256 weight rows), the tensor cores of both SMs read both halves, and each SM accumulates its
// 128 rows x 256 columns in its own tensor memory. Operand traffic per SM drops to 32 KB per step (64 B per clock).
//
// Roles per CTA (192 threads, persistent over tile pairs; cluster = {even CTA = leader, odd CTA = peer}):
//   warp 0     : producer — TMA loads with `.cta_group::2`: the completion bytes of BOTH CTAs' loads are counted on the LEADER's
//                full barrier (address with the peer bit cleared); each CTA waits on its own empty barrier before reusing a stage
//   warp 1     : MMA issuer, leader only — four 256 x 256 x 16 MMAs per stage; `tcgen05.commit ... multicast::cluster` releases the
//                stage in both CTAs and publishes the finished accumulator to both CTAs' epilogues
//   warps 2..5 : epilogue in both CTAs — tcgen05.ld of the CTA's own 128 accumulator rows, fused SwiGLU / residual, bf16 stores;
//                the 8 epilogue warps of the pair hand the accumulator stage back by arriving on the leader's barrier
// SwiGLU: the leader stages 128 gate rows, the peer 128 up rows of the weights — accumulator columns [0, 128) are gate, [128, 256) up.
//
// Taken by ops/functional.py:gemm for plain K-major GEMMs without peer pushes / flags / groups (those stay on the 1-CTA kernel).
// Reference behaviour: the cuBLAS GEMMs behind src/petals/models/llama/block.py:81-125.
#include "common.cuh"
#include "petals_b200.h"

#include <atomic>

extern "C" int pb_set_error(const char* msg);

namespace pb {

bool make_tmap_2d_bf16(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows, uint32_t box_cols);

namespace g2 {

constexpr int BM = 128;        // rows per CTA (256 per pair)
constexpr int BN = 256;        // accumulator columns (MMA N); each CTA stages BN / 2 weight rows
constexpr int BK = 64;
constexpr int kThreads = 192;
constexpr int kGroupM = 8;     // default tile pairs per rasterisation group along M (PETALS_B200_GEMM_GROUP_M overrides: measured in profiles/)
constexpr int A_BYTES = BM * BK * 2, B_BYTES = (BN / 2) * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int STAGES = 6;
constexpr uint32_t TMEM_COLS = 2 * BN;
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;   // clears the bit that selects the odd CTA of a pair in a shared::cluster address

struct Params {
  const __nv_bfloat16* residual;
  __nv_bfloat16* out;
  int M, N, K;
  int ldo, ldres;
  int group_m;   // tile pairs per rasterisation group along M: the concurrent pairs cover group_m x 256 rows of A and (pairs / group_m) weight tiles
  // optional prologue wait (sequence-parallel prefill: the A rows are all-gathered by peers, flag = epoch * per_epoch when they landed)
  const uint64_t* wait_flag;
  uint64_t wait_per_epoch;
  const uint64_t* epoch;
  int* error_flag;
};

PB_DEVICE uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
PB_DEVICE void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
PB_DEVICE void tmem_alloc2(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols) : "memory");
}
PB_DEVICE void tmem_relinquish2() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
PB_DEVICE void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// 2-D TMA load whose completion bytes are counted on the LEADER CTA's mbarrier
PB_DEVICE void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
PB_DEVICE void tc_mma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  const uint32_t z = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(z)
      : "memory");
}
// arrive (once all earlier MMAs of this thread completed) on the barrier at the same shared-memory offset in BOTH CTAs of the pair
PB_DEVICE void tc_commit_pair(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask)
               : "memory");
}
PB_DEVICE void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask) : "memory");
}

PB_DEVICE float silu_f(float x) { return x / (1.f + __expf(-x)); }
PB_DEVICE float round_bf16(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

PB_DEVICE void tile_coords(int tile, int m_pairs, int n_blocks, int group_m, int& m_pair, int& n_blk) {
  const int per_group = group_m * n_blocks;
  const int g = tile / per_group;
  const int first_m = g * group_m;
  const int gsize = min(group_m, m_pairs - first_m);
  const int r = tile - g * per_group;
  m_pair = first_m + r % gsize;
  n_blk = r / gsize;
}

// B_MN: B is the [K, N] row-major view of the weight (the dgrad dX = dY . W on the untransposed nn.Linear weight): each CTA stages its
// 128-wide half of the N extent as two 64-column boxes per K block and the MMA reads B MN-major.
template <bool DUAL, bool B_MN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm_2cta_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const __grid_constant__ CUtensorMap tmap_b2,
                 const Params p) {
  constexpr int OUT_BN = DUAL ? BN / 2 : BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta = cluster_ctarank();
  const bool leader = cta == 0;
  const int m_pairs = (p.M + 2 * BM - 1) / (2 * BM);
  const int n_blocks = (p.N + OUT_BN - 1) / OUT_BN;
  const int num_tiles = m_pairs * n_blocks;
  const int num_kb = (p.K + BK - 1) / BK;
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    if (DUAL) prefetch_tmap(&tmap_b2);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);    // the leader's producer arrives (expect_tx); bytes of both CTAs complete on the leader's copy
      mbar_init(&empty_bar[s], 1);   // one multicast commit per use, in each CTA
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);   // one multicast commit per tile, in each CTA
      mbar_init(&tmem_empty[s], 8);  // 4 epilogue warps x 2 CTAs arrive on the leader's copy
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    tmem_alloc2(tmem_slot, TMEM_COLS);
    tmem_relinquish2();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();   // barriers of both CTAs are initialised and both allocations are done before any remote arrive / pair MMA
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =============================== producer (both CTAs) ===============================
    if (lane == 0) {
      if (p.wait_flag != nullptr) {
        if (!spin_wait_ge(p.wait_flag, *p.epoch * p.wait_per_epoch) && p.error_flag != nullptr) atomicExch(p.error_flag, 1);
        asm volatile("fence.proxy.async;" ::: "memory");
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
        int m_pair, n_blk;
        tile_coords(tile, m_pairs, n_blocks, p.group_m, m_pair, n_blk);
        const int a_row = m_pair * 2 * BM + static_cast<int>(cta) * BM;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
          tma_load_2d_pair(sa, &tmap_a, &full_bar[stage], kb * BK, a_row);
          if (B_MN) {
#pragma unroll
            for (int j = 0; j < BN / 2 / 64; ++j)   // [64 k-rows x 64 n-columns] boxes, 8 KB each
              tma_load_2d_pair(sb + j * 8192, &tmap_b, &full_bar[stage], n_blk * BN + static_cast<int>(cta) * (BN / 2) + j * 64, kb * BK);
          } else if (!DUAL) {
            tma_load_2d_pair(sb, &tmap_b, &full_bar[stage], kb * BK, n_blk * BN + static_cast<int>(cta) * (BN / 2));
          } else {  // leader: gate rows, peer: up rows of the same output columns
            tma_load_2d_pair(sb, leader ? &tmap_b : &tmap_b2, &full_bar[stage], kb * BK, n_blk * OUT_BN);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer (leader CTA only) =================================
    if (leader && lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * BM, BN, 0, B_MN ? 1 : 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += n_clusters, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t sb = sa + A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            tc_mma_f16_pair(d_tmem, umma_desc_k_sw128(sa + k * 32), B_MN ? umma_desc_mn_sw128(sb + k * 2048, 8192) : umma_desc_k_sw128(sb + k * 32), idesc,
                            (kb | k) != 0 ? 1u : 0u);
          tc_commit_pair(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit_pair(&tmem_full[as]);
      }
    }
  } else {
    // =============================== epilogue (both CTAs: own 128 rows) ====================================
    const int quarter = warp & 3;
    int it = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += n_clusters, ++it) {
      int m_pair, n_blk;
      tile_coords(tile, m_pairs, n_blocks, p.group_m, m_pair, n_blk);
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after();
      const int row = m_pair * 2 * BM + static_cast<int>(cta) * BM + quarter * 32 + lane;
      const bool row_ok = row < p.M;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * BN;
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
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&tmem_empty[as]);
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync();   // the peer may still be reading its accumulator / receiving commits: nobody frees anything before both are done
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, TMEM_COLS);
  }
}

template <bool DUAL, bool B_MN>
static int launch(const PbGemmArgs* a, cudaStream_t stream) {
  constexpr int OUT_BN = DUAL ? BN / 2 : BN;
  const int lda = a->lda > 0 ? a->lda : a->K, ldb = a->ldb > 0 ? a->ldb : a->K;
  CUtensorMap ta, tb, tb2;
  if (!make_tmap_2d_bf16(&ta, a->a, a->M, a->K, lda, BM, BK)) return PB_ERR_DRIVER;
  if (B_MN) {  // B is [K, N] row-major: box = 64 k-rows x 64 n-columns
    if (!make_tmap_2d_bf16(&tb, a->b, a->K, a->N, a->ldb > 0 ? a->ldb : a->N, BK, 64)) return PB_ERR_DRIVER;
  } else if (!make_tmap_2d_bf16(&tb, a->b, a->N, a->K, ldb, BN / 2, BK)) {
    return PB_ERR_DRIVER;
  }
  tb2 = tb;
  if (DUAL && !make_tmap_2d_bf16(&tb2, a->b2, a->N, a->K, ldb, BN / 2, BK)) return PB_ERR_DRIVER;
  Params p{};
  p.residual = static_cast<const __nv_bfloat16*>(a->residual); p.out = static_cast<__nv_bfloat16*>(a->out);
  p.M = a->M; p.N = a->N; p.K = a->K; p.ldo = a->ldo > 0 ? a->ldo : a->N; p.ldres = a->ldres > 0 ? a->ldres : a->N;
  static const int env_group = [] { const char* e = getenv("PETALS_B200_GEMM_GROUP_M"); const int v = e ? atoi(e) : 0; return v > 0 ? v : kGroupM; }();
  p.group_m = env_group;
  p.wait_flag = static_cast<const uint64_t*>(a->wait_flag); p.wait_per_epoch = a->wait_per_epoch; p.epoch = static_cast<const uint64_t*>(a->epoch);
  p.error_flag = static_cast<int*>(a->error_flag);
  const int smem = STAGES * STAGE_BYTES + 1024 + 256;
  auto kern = gemm_2cta_kernel<DUAL, B_MN>;
  static std::atomic<int> max_clusters[64];   // 0 = not asked yet
  int dev = 0;
  cudaGetDevice(&dev);
  const int sms = a->num_sms > 0 ? a->num_sms : 148;
  if (max_clusters[dev & 63].load() == 0) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return pb_check_launch("gemm_2cta attr");
    // a persistent grid must not exceed what is co-resident: a pair that waits for a free TPC would start its tiles when the others finish
    cudaLaunchConfig_t cfg = {};
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
  kern<<<2 * clusters, kThreads, smem, stream>>>(ta, tb, tb2, p);
  return pb_check_launch("gemm_2cta");
}

}  // namespace g2
}  // namespace pb

using namespace pb;

// Same argument block as pb_gemm_bf16; refuses (PB_ERR_UNSUPPORTED) what only the 1-CTA kernel does.
extern "C" int pb_gemm_bf16_2cta(const PbGemmArgs* a, void* stream) {
  if (a == nullptr || a->M <= 0 || a->N <= 0 || a->K <= 0 || (a->K & 7) || (a->N & 7)) return PB_ERR_SHAPE;
  if ((a->b_mn_major && a->act != 0) || a->n_push > 0 || (a->wait_flag != nullptr && a->epoch == nullptr) || a->grp != nullptr || a->bias != nullptr || a->bias2 != nullptr || a->out_fp32 ||
      a->accumulate || a->out == nullptr || (a->act != 0 && a->act != 1)) {
    pb_set_error("gemm_2cta: plain K-major GEMM (optional SwiGLU / residual) only");
    return PB_ERR_UNSUPPORTED;
  }
  if (a->act == 1) {
    if (a->b2 == nullptr) return PB_ERR_SHAPE;
    return g2::launch<true, false>(a, static_cast<cudaStream_t>(stream));
  }
  if (a->b_mn_major) return g2::launch<false, true>(a, static_cast<cudaStream_t>(stream));
  return g2::launch<false, false>(a, static_cast<cudaStream_t>(stream));
}
