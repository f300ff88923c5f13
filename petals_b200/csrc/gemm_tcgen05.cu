// petals_b200 — persistent warp-specialised tcgen05 GEMM for sm_100a.
//
//   D[M, N] = epilogue( A[M, K] · op(B) )      bf16 operands, fp32 accumulation in TMEM
//
// Roles inside one 192-thread CTA (one CTA per SM, persistent over output tiles):
//   warp 0     : TMA producer   — cp.async.bulk.tensor 2-D loads (128B swizzle) into a multi-stage
//                                 shared-memory ring, signalled through mbarrier complete_tx
//   warp 1     : MMA issuer     — one thread issues tcgen05.mma (128 x BN x 16) from shared-memory
//                                 descriptors into a double-buffered TMEM accumulator and releases
//                                 ring slots / publishes accumulators with tcgen05.commit
//   warps 2..5 : epilogue       — tcgen05.ld TMEM -> registers, fused bias / GELU / SwiGLU /
//                                 residual, bf16 (or fp32) stores to local HBM and, for the fused
//                                 stage hop / tensor-parallel push, straight into peer GPUs' buffers
//                                 over NVLink followed by a release-increment of the peers' flags.
//
// B can be the nn.Linear weight as stored by HF ([N, K], K-major: forward pass) or the same tensor
// consumed transposed ([K, N], MN-major: the dX = dY·W dgrad of the prompt-tuning backward pass,
// SURVEY.md §2.5 L20) — no transposed weight copy is kept.
// With `act == 1` the B tile is half gate_proj rows, half up_proj rows and the epilogue emits
// silu(gate)·up directly, so the [T, 2I] intermediate never touches HBM.
//
// Replaces every cuBLAS call the reference makes per block and the
// inter-server activation RPC (X1-X4). Reference behaviour: src/petals/models/llama/block.py:81-125.
#include "common.cuh"
#include "petals_b200.h"

#include <atomic>
#include <mutex>

namespace pb {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kGemmThreads = 192;
constexpr bool kPushTransposed = true;  // peer-push epilogue through a shared-memory transpose (full-sector NVLink stores)
constexpr int kGroupM = 8;

struct GemmParams {
  const __nv_bfloat16* bias;
  const __nv_bfloat16* bias2;
  const __nv_bfloat16* residual;
  void* out;
  int M, N, K;          // N = number of output columns
  int ldo, ldres;
  int act, out_fp32;
  int n_push;
  void* push_out[PB_MAX_PEERS];
  uint64_t* push_flag[PB_MAX_PEERS];       // per-tile release-increment (lets a consumer start on finished row blocks)
  uint64_t* push_done_flag[PB_MAX_PEERS];  // ONE release-increment per launch, by the last CTA to finish all its tiles
  unsigned int* done_counter;              // local self-resetting counter electing that last CTA
  int push_rows_per_owner;                 // 0: every tile goes to all n_push peers (broadcast / stage hop);
                                           // >0: reduce-scatter routing — row r goes ONLY to peer r / rows_per_owner, at local row r % rows_per_owner
  const uint64_t* wait_flag;
  uint64_t wait_per_epoch;
  const uint64_t* epoch;
  int* error_flag;
  // Grouped (ragged-M) mode for sparse MoE layers: the M dimension is a concatenation of per-expert row groups and every group
  // multiplies its own weight matrix (the experts' weights are stacked along N: expert e owns rows [e * grp_n, (e+1) * grp_n) of B).
  // `grp` is a DEVICE table built by moe_plan_kernel — {n_mtiles, expert[], first row[], valid rows[]} — so the host never learns
  // (or waits for) the routing: the persistent CTAs read the tile count when they start.
  const int* grp;
  int grp_cap;   // capacity of each table array
  int grp_n;     // B rows per expert
};

PB_DEVICE float gelu_tanh_f(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  return 0.5f * x * (1.f + tanhf(k0 * (x + k1 * x * x * x)));
}
PB_DEVICE float gelu_erf_f(float x) { return 0.5f * x * (1.f + erff(x * 0.7071067811865475f)); }
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

// BN   : accumulator width (MMA N). With DUAL the output tile is BN/2 columns wide.
// B_MN : B operand is MN-major ([K, N] row-major in global memory).
// DUAL : fused SwiGLU (two B tensors).
template <int BN, bool B_MN, bool DUAL, int STAGES>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_b2, const GemmParams p) {
  constexpr int A_BYTES = BM * BK * 2;
  constexpr int B_BYTES = BN * BK * 2;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int OUT_BN = DUAL ? BN / 2 : BN;
  constexpr uint32_t TMEM_COLS = 2 * BN;
  static_assert(TMEM_COLS <= 512, "TMEM overflow");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  // per-epilogue-warp transpose buffer for peer pushes: 32 rows x (64 B payload + 16 B pad)
  uint8_t* push_stage = reinterpret_cast<uint8_t*>(tmem_slot) + 64;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool grouped = p.grp != nullptr;
  const int m_blocks = grouped ? min(p.grp[0], p.grp_cap) : (p.M + BM - 1) / BM;
  const int* g_expert = p.grp + 1;
  const int* g_row0 = g_expert + p.grp_cap;
  const int* g_rows = g_row0 + p.grp_cap;
  const int n_blocks = (p.N + OUT_BN - 1) / OUT_BN;
  const int num_tiles = m_blocks * n_blocks;
  const int num_kb = (p.K + BK - 1) / BK;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    if (DUAL) prefetch_tmap(&tmap_b2);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 4);
    }
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
    // =============================== TMA producer ===============================
    if (lane == 0) {
      if (p.wait_flag != nullptr) {
        const uint64_t target = *p.epoch * p.wait_per_epoch;
        if (!spin_wait_ge(p.wait_flag, target)) atomicExch(p.error_flag, 1);
        asm volatile("fence.proxy.async;" ::: "memory");
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int m_blk, n_blk;
        tile_coords(tile, m_blocks, n_blocks, m_blk, n_blk);
        const int a_row = grouped ? g_row0[m_blk] : m_blk * BM;
        const int b_row = grouped ? g_expert[m_blk] * p.grp_n : 0;   // this group's expert: its slab of the stacked weights
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
          tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * BK, a_row);
          if (!B_MN) {
            if (!DUAL) {
              tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * BK, b_row + n_blk * BN);
            } else {
              tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * BK, b_row + n_blk * OUT_BN);
              tma_load_2d(sb + B_BYTES / 2, &tmap_b2, &full_bar[stage], kb * BK, b_row + n_blk * OUT_BN);
            }
          } else {
            // MN-major: BN/64 boxes of [64 k-rows x 64 n-elements], 8 KB each
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_2d(sb + j * 8192, &tmap_b, &full_bar[stage], n_blk * BN + j * 64, kb * BK);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer =================================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, 0, B_MN ? 1 : 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
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
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t adesc = umma_desc_k_sw128(sa + k * 32);
            const uint64_t bdesc = B_MN ? umma_desc_mn_sw128(sb + k * 2048, 8192)
                                        : umma_desc_k_sw128(sb + k * 32);
            tc_mma_f16(d_tmem, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          tc_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit(&tmem_full[as]);
      }
    }
  } else {
    // =============================== epilogue ====================================
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      int m_blk, n_blk;
      tile_coords(tile, m_blocks, n_blocks, m_blk, n_blk);
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after();
      const int tile_row0 = grouped ? g_row0[m_blk] : m_blk * BM;
      const int row = tile_row0 + quarter * 32 + lane;
      const bool row_ok = grouped ? (quarter * 32 + lane) < g_rows[m_blk] : row < p.M;   // grouped: rows past the group's end belong to the next expert
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
        if (p.bias != nullptr) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (col0 + i < p.N) v[i] += __bfloat162float(__ldg(p.bias + col0 + i));
        }
        if (DUAL) {
          uint32_t r2[32];
          tmem_ld_32x32(taddr + OUT_BN + c, r2);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float u = __uint_as_float(r2[i]);
            if (p.bias2 != nullptr && col0 + i < p.N) u += __bfloat162float(__ldg(p.bias2 + col0 + i));
            // HF rounding: bf16(silu(bf16(gate))) * bf16(up)
            v[i] = round_bf16(silu_f(round_bf16(v[i]))) * round_bf16(u);
          }
        } else if (p.act == 2) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = gelu_tanh_f(v[i]);
        } else if (p.act == 3) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = gelu_erf_f(v[i]);
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
              for (int i = 0; i < 32 && col0 + i < p.N; ++i)
                v[i] = round_bf16(v[i]) + __bfloat162float(rp[i]);
            }
          }
          if (p.out_fp32) {
            float* op = static_cast<float*>(p.out) + static_cast<size_t>(row) * p.ldo + col0;
            if (full) {
#pragma unroll
              for (int q = 0; q < 8; ++q)
                *reinterpret_cast<float4*>(op + q * 4) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            } else {
              for (int i = 0; i < 32 && col0 + i < p.N; ++i) op[i] = v[i];
            }
          } else {
            uint4 pk[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              pk[q].x = pack_bf16(v[q * 8 + 0], v[q * 8 + 1]);
              pk[q].y = pack_bf16(v[q * 8 + 2], v[q * 8 + 3]);
              pk[q].z = pack_bf16(v[q * 8 + 4], v[q * 8 + 5]);
              pk[q].w = pack_bf16(v[q * 8 + 6], v[q * 8 + 7]);
            }
            const size_t off = static_cast<size_t>(row) * p.ldo + col0;
            if (full) {
              if (p.out != nullptr) {
                uint4* op = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + off);
#pragma unroll
                for (int q = 0; q < 4; ++q) op[q] = pk[q];
              }
              if (p.n_push > 0 && kPushTransposed) {
                uint4* st = reinterpret_cast<uint4*>(push_stage + (warp - 2) * (32 * 80) + lane * 80);
#pragma unroll
                for (int q = 0; q < 4; ++q) st[q] = pk[q];
              }
              if (p.n_push > 0 && !kPushTransposed) {
                if (p.push_rows_per_owner > 0) {
                  const int owner = row / p.push_rows_per_owner;
                  const size_t roff = static_cast<size_t>(row - owner * p.push_rows_per_owner) * p.ldo + col0;
                  uint4* op = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.push_out[owner]) + roff);
#pragma unroll
                  for (int q = 0; q < 4; ++q) op[q] = pk[q];
                } else {
                  for (int rnk = 0; rnk < p.n_push; ++rnk) {
                    uint4* op = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.push_out[rnk]) + off);
#pragma unroll
                    for (int q = 0; q < 4; ++q) op[q] = pk[q];
                  }
                }
              }
            } else {
              for (int i = 0; i < 32 && col0 + i < p.N; ++i) {
                const __nv_bfloat16 bv = __float2bfloat16_rn(v[i]);
                if (p.out != nullptr) static_cast<__nv_bfloat16*>(p.out)[off + i] = bv;
                if (p.push_rows_per_owner > 0) {
                  const int owner = row / p.push_rows_per_owner;
                  static_cast<__nv_bfloat16*>(p.push_out[owner])[static_cast<size_t>(row - owner * p.push_rows_per_owner) * p.ldo + col0 + i] = bv;
                } else {
                  for (int rnk = 0; rnk < p.n_push; ++rnk)
                    static_cast<__nv_bfloat16*>(p.push_out[rnk])[off + i] = bv;
                }
              }
            }
          }
        }
        if (kPushTransposed && p.n_push > 0 && !p.out_fp32 && col0 + 32 <= p.N) {
          // Peer stores, transposed through shared memory: instead of 32 lanes writing 16 bytes to 32 different rows (partial
          // sectors on NVLink), 4 consecutive lanes write one row's 64 contiguous bytes -> 8 rows x 2 full sectors per instruction.
          __syncwarp();
          const uint8_t* sbase = push_stage + (warp - 2) * (32 * 80);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rr = 8 * i + (lane >> 2), piece = lane & 3;
            const int grow = tile_row0 + quarter * 32 + rr;
            if (grouped ? (quarter * 32 + rr) < g_rows[m_blk] : grow < p.M) {
              const uint4 val = *reinterpret_cast<const uint4*>(sbase + rr * 80 + piece * 16);
              if (p.push_rows_per_owner > 0) {
                const int owner = grow / p.push_rows_per_owner;
                const size_t roff = static_cast<size_t>(grow - owner * p.push_rows_per_owner) * p.ldo + col0 + piece * 8;
                *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.push_out[owner]) + roff) = val;
              } else {
                const size_t goff = static_cast<size_t>(grow) * p.ldo + col0 + piece * 8;
                for (int rnk = 0; rnk < p.n_push; ++rnk)
                  *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.push_out[rnk]) + goff) = val;
              }
            }
          }
          __syncwarp();  // the staging rows are rewritten by the next 32-column chunk
        }
      }
      // accumulator fully read: hand the TMEM stage back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[as]);
      if (p.n_push > 0) {
        // all 4 epilogue warps finished this tile's peer stores -> one release per peer
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (warp == 2 && lane == 0) {
          __threadfence_system();
          for (int rnk = 0; rnk < p.n_push; ++rnk)
            if (p.push_flag[rnk] != nullptr) red_release_sys_add(p.push_flag[rnk], 1ull);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
  if (p.n_push > 0 && p.done_counter != nullptr && threadIdx.x == 64) {
    // every peer store of this CTA happened before the __syncthreads above; publish completion of the whole GEMM once
    __threadfence_system();
    const unsigned int prev = atomicAdd(p.done_counter, 1u);
    if (prev == gridDim.x - 1) {
      __threadfence_system();
      *p.done_counter = 0u;
      for (int rnk = 0; rnk < p.n_push; ++rnk)
        if (p.push_done_flag[rnk] != nullptr) red_release_sys_add(p.push_done_flag[rnk], 1ull);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// host side: tensor maps + launch
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  });
  return fn;
}

// 2-D bf16 tensor [rows, cols] with leading dimension ld (elements); box = [box_rows, box_cols].
static bool make_tmap_2d(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                         uint32_t box_rows, uint32_t box_cols) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return false;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// 2-D byte tensor (FP8 payloads) with the same 128B swizzle: one box row = box_cols bytes (gemm_mxfp8.cu)
bool make_tmap_2d_u8(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows, uint32_t box_cols) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return false;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// scale-factor blocks (gemm_mxfp8.cu, 2-CTA variant): the packed scale array as [blocks, 128] 32-bit words, no swizzle, box = `box_rows` blocks
bool make_tmap_sf_blocks(CUtensorMap* m, const void* base, uint64_t blocks, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return false;
  cuuint64_t gdim[2] = {128, blocks};
  cuuint64_t gstride[1] = {512};
  cuuint32_t box[2] = {128, box_rows};
  cuuint32_t estr[2] = {1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// exported to the other translation units that build TMA descriptors (attention_tc.cu)
bool make_tmap_2d_bf16(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows, uint32_t box_cols) {
  return make_tmap_2d(m, base, rows, cols, ld, box_rows, box_cols);
}

template <int BN, bool B_MN, bool DUAL>
static int launch_gemm(const PbGemmArgs* a, cudaStream_t stream) {
  constexpr int STAGE_BYTES = (BM + BN) * BK * 2;
  constexpr int STAGES = (200 * 1024) / STAGE_BYTES > 8 ? 8 : (200 * 1024) / STAGE_BYTES;
  constexpr int OUT_BN = DUAL ? BN / 2 : BN;
  const size_t smem = static_cast<size_t>(STAGES) * STAGE_BYTES + 1024 + 256 + 4 * 32 * 80;  // + per-warp push transpose buffers

  const int lda = a->lda > 0 ? a->lda : a->K;
  CUtensorMap ta, tb, tb2;
  if (!make_tmap_2d(&ta, a->a, a->M, a->K, lda, BM, BK)) return PB_ERR_DRIVER;
  const uint64_t b_rows = a->grp != nullptr ? static_cast<uint64_t>(a->grp_experts) * a->N : a->N;   // grouped: experts stacked along N
  if (!B_MN) {
    const int ldb = a->ldb > 0 ? a->ldb : a->K;
    const uint32_t box_rows = DUAL ? BN / 2 : BN;
    if (!make_tmap_2d(&tb, a->b, b_rows, a->K, ldb, box_rows, BK)) return PB_ERR_DRIVER;
    if (DUAL) {
      if (!make_tmap_2d(&tb2, a->b2, b_rows, a->K, ldb, box_rows, BK)) return PB_ERR_DRIVER;
    } else {
      tb2 = tb;
    }
  } else {
    const int ldb = a->ldb > 0 ? a->ldb : a->N;
    if (!make_tmap_2d(&tb, a->b, a->K, a->N, ldb, BK, 64)) return PB_ERR_DRIVER;
    tb2 = tb;
  }

  GemmParams p{};
  p.bias = static_cast<const __nv_bfloat16*>(a->bias);
  p.bias2 = static_cast<const __nv_bfloat16*>(a->bias2);
  p.residual = static_cast<const __nv_bfloat16*>(a->residual);
  p.out = a->out;
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.ldo = a->ldo > 0 ? a->ldo : a->N;
  p.ldres = a->ldres > 0 ? a->ldres : a->N;
  p.act = a->act; p.out_fp32 = a->out_fp32;
  p.n_push = a->n_push;
  for (int i = 0; i < a->n_push; ++i) {
    p.push_out[i] = a->push_out[i];
    p.push_flag[i] = static_cast<uint64_t*>(a->push_flag[i]);
    p.push_done_flag[i] = static_cast<uint64_t*>(a->push_done_flag[i]);
  }
  p.done_counter = static_cast<unsigned int*>(a->done_counter);
  p.push_rows_per_owner = a->push_rows_per_owner;
  p.wait_flag = static_cast<const uint64_t*>(a->wait_flag);
  p.wait_per_epoch = a->wait_per_epoch;
  p.epoch = static_cast<const uint64_t*>(a->epoch);
  p.error_flag = static_cast<int*>(a->error_flag);
  p.grp = static_cast<const int*>(a->grp); p.grp_cap = a->grp_cap; p.grp_n = a->N;

  auto kern = gemm_tcgen05_kernel<BN, B_MN, DUAL, STAGES>;
  // once per (instantiation, device): function attributes are per device, and several stage threads may get here at once
  static std::atomic<uint64_t> attr_done{0};
  int dev = 0;
  cudaGetDevice(&dev);
  const uint64_t bit = 1ull << (dev & 63);
  if (!(attr_done.load(std::memory_order_acquire) & bit)) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess)
      return PB_ERR_CUDA;
    attr_done.fetch_or(bit, std::memory_order_release);
  }
  const int sms = a->num_sms > 0 ? a->num_sms : 148;
  const int m_blocks = (a->M + BM - 1) / BM, n_blocks = (a->N + OUT_BN - 1) / OUT_BN;
  const int tiles = (a->grp != nullptr ? a->grp_cap : m_blocks) * n_blocks;   // grouped: the real count lives on the device; size the grid for the cap
  const int grid = tiles < sms ? tiles : sms;
  kern<<<grid, kGemmThreads, smem, stream>>>(ta, tb, tb2, p);
  return pb_check_launch("gemm_tcgen05");
}

}  // namespace pb

extern "C" int pb_gemm_tiles(int M, int N, int block_n, int dual) {
  const int out_bn = dual ? block_n / 2 : block_n;
  return ((M + pb::BM - 1) / pb::BM) * ((N + out_bn - 1) / out_bn);
}

extern "C" int pb_gemm_bf16(const PbGemmArgs* a, void* stream) {
  using namespace pb;
  if (a->M <= 0 || a->N <= 0 || a->K <= 0) return PB_ERR_SHAPE;
  if ((a->K & 7) || (a->N & 7)) return PB_ERR_SHAPE;  // 16-byte global strides for TMA / vector stores
  if (a->n_push > PB_MAX_PEERS) return PB_ERR_SHAPE;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool dual = a->act == 1;
  if (dual && (a->b2 == nullptr || a->b_mn_major)) return PB_ERR_SHAPE;
  if (a->grp != nullptr && (a->b_mn_major || a->n_push > 0 || a->grp_cap <= 0 || a->grp_experts <= 0)) return PB_ERR_SHAPE;
  int bn = a->block_n;
  if (bn == 0) {
    // Enough tiles to fill the machine with the widest accumulator that still does so.
    const int sms = a->num_sms > 0 ? a->num_sms : 148;
    const int m_blocks = (a->M + BM - 1) / BM;
    bn = 256;
    if (dual) {
      bn = 256;
    } else {
      while (bn > 64 && static_cast<long>(m_blocks) * ((a->N + bn - 1) / bn) < sms) bn >>= 1;
    }
  }
  if (dual) {
    if (bn == 256) return launch_gemm<256, false, true>(a, s);
    if (bn == 128) return launch_gemm<128, false, true>(a, s);
    return PB_ERR_SHAPE;
  }
  if (a->b_mn_major) {
    if (bn == 256) return launch_gemm<256, true, false>(a, s);
    if (bn == 128) return launch_gemm<128, true, false>(a, s);
    if (bn == 64) return launch_gemm<64, true, false>(a, s);
    return PB_ERR_SHAPE;
  }
  if (bn == 256) return launch_gemm<256, false, false>(a, s);
  if (bn == 128) return launch_gemm<128, false, false>(a, s);
  if (bn == 64) return launch_gemm<64, false, false>(a, s);
  return PB_ERR_SHAPE;
}
