// Persistent, warp-specialised tcgen05 GEMM engine for sm_100a.
//
//   acc[128 x BN] (fp32, TMEM) = A_tile * B_tile^t   over a K range, A/B tiles staged by TMA into
//   128-byte-swizzled shared memory, `tcgen05.mma` issued by one thread, accumulators double
//   buffered in TMEM so the epilogue of tile i overlaps the main loop of tile i+1.
//
// Warp roles (384 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM
// allocator, warps 4..11 = epilogue: warp w reads TMEM lanes 32*(w%4).. and the column half
// (w-4)/4 of the tile, so two warps share each row.
//
// Operand precision: bf16 (kind::f16, one product) or fp32 carried as a (hi, lo) pair of tf32
// arrays with three products hi*hi + hi*lo + lo*hi (kind::tf32) -- fp32-level accuracy.
// Operand layouts: K-major ([rows, K] row-major) or MN-major ([K, rows] row-major), both through
// the canonical SWIZZLE_128B shared-memory layouts.
//
// The epilogue is a functor template parameter: row/column softmax statistics, on-the-fly softmax
// gradient, or plain scaled store / split-K accumulate.
#pragma once

#include "ce_common.cuh"

namespace ce {

constexpr int kBM = 128;           // tile rows (UMMA M)
constexpr int kSwizzleBytes = 128; // one swizzle atom = BLOCK_K
constexpr int kGemmThreads = 384;   // 4 control warps + 8 epilogue warps
constexpr int kEpiThreads = 256;

struct GemmShape {
  int M, N, K;             // problem extents (rows of A, rows of B, reduction)
  int num_m_blk, num_n_blk;
  int k_splits;            // split-K factor (work item = tile x split)
  int kblk_total;          // reduction length in BLOCK_K units
  int kblk_per_split;
  int a_mn, b_mn;          // operand majorness: 0 = K-major, 1 = MN-major
};

template <bool TF32X3, int BN, int EPI_STAGE = 0>   // EPI_STAGE: bytes of epilogue staging smem
struct GemmCfg {
  static constexpr int kElemBytes = TF32X3 ? 4 : 2;
  static constexpr int kBK = kSwizzleBytes / kElemBytes;     // elements of K per stage: 32 / 64
  static constexpr int kUmmaK = 32 / kElemBytes;             // K per instruction: 8 / 16
  static constexpr int kParts = TF32X3 ? 2 : 1;              // hi, lo
  static constexpr int kABytes = kBM * kSwizzleBytes;        // one A part per stage
  static constexpr int kBBytes = BN * kSwizzleBytes;
  static constexpr int kStageBytes = kParts * (kABytes + kBBytes);
  static constexpr int kBudget = 216 * 1024 - EPI_STAGE;
  static constexpr int kStages = kBudget / kStageBytes > 6 ? 6 : kBudget / kStageBytes;
  static constexpr int kEpiStageBytes = EPI_STAGE;
  static constexpr int kTmemCols = 2 * BN;                   // double-buffered accumulator
  static constexpr int kEpiFloats = 4 * BN;                  // per-tile column data for the epilogue
  static constexpr size_t kSmemBytes = 1024 /*align slack*/ + (size_t)kStages * kStageBytes +
                                       sizeof(float) * kEpiFloats + 256 /*barriers*/ + EPI_STAGE;
  static_assert(kStages >= 2, "pipeline needs at least two stages");
  static_assert(kTmemCols <= 512 && (kTmemCols & (kTmemCols - 1)) == 0, "TMEM columns");
};

struct TmapSet {
  CUtensorMap a[2];  // hi, lo (lo unused for bf16)
  CUtensorMap b[2];
};

// Epilogue interface:
//   struct Epi { Params p;
//     __device__ void tile_begin(float* s_epi, int m_blk, int n_blk, int et /*0..127*/);   (all 128 epilogue threads, followed by a 128-thread barrier)
//     __device__ void row_begin(int row /*global row*/, bool row_ok);
//     __device__ void chunk(const float* acc /*32 cols*/, int col0 /*global col of acc[0]*/, int lcol0 /*tile-local*/, const float* s_epi);
//     __device__ void row_end(int row, bool row_ok, int m_blk, int n_blk, int k_split, float* s_epi, int et);
//   };

template <bool TF32X3, int BN, class Epi>
__global__ void __launch_bounds__(kGemmThreads, 1)
umma_gemm_kernel(const __grid_constant__ TmapSet tm, const GemmShape gs, const typename Epi::Params ep) {
  using Cfg = GemmCfg<TF32X3, BN, Epi::kStageBytesPerWarp * 8>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* stage_base = smem;
  float* s_epi = reinterpret_cast<float*>(smem + (size_t)Cfg::kStages * Cfg::kStageBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_epi + Cfg::kEpiFloats);
  uint64_t* full_bar = bars;                       // [kStages]
  uint64_t* empty_bar = bars + Cfg::kStages;       // [kStages]
  uint64_t* tfull_bar = bars + 2 * Cfg::kStages;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;            // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  uint8_t* epi_stage = reinterpret_cast<uint8_t*>(bars) + 256;   // [8 warps][kStageBytesPerWarp]

  const int warp = warp_id(), lane = lane_id();
  const int num_tiles = gs.num_m_blk * gs.num_n_blk;
  const int num_items = num_tiles * gs.k_splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm.a[0]);
    tma_prefetch_desc(&tm.b[0]);
    if (TF32X3) { tma_prefetch_desc(&tm.a[1]); tma_prefetch_desc(&tm.b[1]); }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < Cfg::kStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 8); }
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ================= TMA producer =================
    int stage = 0; uint32_t phase = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const int tile = item % num_tiles, ksp = item / num_tiles;
      const int m_blk = tile % gs.num_m_blk, n_blk = tile / gs.num_m_blk;
      const int kb0 = ksp * gs.kblk_per_split;
      const int kb1 = min(kb0 + gs.kblk_per_split, gs.kblk_total);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* st = stage_base + (size_t)stage * Cfg::kStageBytes;
        mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
#pragma unroll
        for (int part = 0; part < Cfg::kParts; ++part) {
          uint8_t* sa = st + part * Cfg::kABytes;
          uint8_t* sb = st + Cfg::kParts * Cfg::kABytes + part * Cfg::kBBytes;
          if (!gs.a_mn) {
            tma_load_2d(sa, &tm.a[part], &full_bar[stage], kb * Cfg::kBK, m_blk * kBM);
          } else {  // [K, M] row-major: boxes of (kBK rows of K) x (one 128-byte run of M)
#pragma unroll
            for (int c = 0; c < kBM / Cfg::kBK; ++c)
              tma_load_2d(sa + c * Cfg::kBK * kSwizzleBytes, &tm.a[part], &full_bar[stage],
                          m_blk * kBM + c * Cfg::kBK, kb * Cfg::kBK);
          }
          if (!gs.b_mn) {
            tma_load_2d(sb, &tm.b[part], &full_bar[stage], kb * Cfg::kBK, n_blk * BN);
          } else {
#pragma unroll
            for (int c = 0; c < BN / Cfg::kBK; ++c)
              tma_load_2d(sb + c * Cfg::kBK * kSwizzleBytes, &tm.b[part], &full_bar[stage],
                          n_blk * BN + c * Cfg::kBK, kb * Cfg::kBK);
          }
        }
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ================= MMA issuer =================
    const uint32_t idesc = umma_idesc(TF32X3, kBM, BN, gs.a_mn != 0, gs.b_mn != 0);
    // K-major: 8-row groups 1024 B apart, K advances 32 B inside the swizzle atom.
    // MN-major: 8-K-row groups 1024 B apart, 128-byte MN runs kBK*128 B apart, K advances by rows.
    const uint32_t a_lbo = gs.a_mn ? Cfg::kBK * kSwizzleBytes : 16, b_lbo = gs.b_mn ? Cfg::kBK * kSwizzleBytes : 16;
    const uint32_t a_kstep = gs.a_mn ? Cfg::kUmmaK * kSwizzleBytes : 32;
    const uint32_t b_kstep = gs.b_mn ? Cfg::kUmmaK * kSwizzleBytes : 32;
    // tf32 MN-major operands only exist in the 32-byte-atom flavour of the 128-byte swizzle
    // (4-row K groups, 512 B apart); everything else uses the plain 128-byte swizzle.
    const uint32_t a_lt = (TF32X3 && gs.a_mn) ? 1u : 2u, b_lt = (TF32X3 && gs.b_mn) ? 1u : 2u;
    const uint32_t a_sbo = (TF32X3 && gs.a_mn) ? 512u : 1024u, b_sbo = (TF32X3 && gs.b_mn) ? 512u : 1024u;
    int stage = 0; uint32_t phase = 0;
    int acc_stage = 0; uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const int ksp = item / num_tiles;
      const int kb0 = ksp * gs.kblk_per_split;
      const int kb1 = min(kb0 + gs.kblk_per_split, gs.kblk_total);
      mbar_wait(&tempty_bar[acc_stage], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc_stage * BN;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t st = smem_u32(stage_base + (size_t)stage * Cfg::kStageBytes);
        const uint32_t sa_hi = st, sa_lo = st + Cfg::kABytes;
        const uint32_t sb_hi = st + Cfg::kParts * Cfg::kABytes, sb_lo = sb_hi + Cfg::kBBytes;
#pragma unroll
        for (int kk = 0; kk < Cfg::kBK / Cfg::kUmmaK; ++kk) {
          const uint32_t first = (kb == kb0 && kk == 0) ? 0u : 1u;
          const uint64_t da_hi = umma_smem_desc(sa_hi + kk * a_kstep, a_lbo, a_sbo, a_lt);
          const uint64_t db_hi = umma_smem_desc(sb_hi + kk * b_kstep, b_lbo, b_sbo, b_lt);
          if constexpr (TF32X3) {
            const uint64_t da_lo = umma_smem_desc(sa_lo + kk * a_kstep, a_lbo, a_sbo, a_lt);
            const uint64_t db_lo = umma_smem_desc(sb_lo + kk * b_kstep, b_lbo, b_sbo, b_lt);
            umma<true>(tmem_d, da_lo, db_hi, idesc, first);
            umma<true>(tmem_d, da_hi, db_lo, idesc, 1u);
            umma<true>(tmem_d, da_hi, db_hi, idesc, 1u);
          } else {
            umma<false>(tmem_d, da_hi, db_hi, idesc, first);
          }
        }
        umma_commit(&empty_bar[stage]);  // frees the smem stage once these MMAs retire
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
      umma_commit(&tfull_bar[acc_stage]);  // accumulator ready for the epilogue
      if (++acc_stage == 2) { acc_stage = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4) {
    // ================= epilogue =================
    Epi epi;
    epi.p = ep;
    epi.stage = epi_stage + (warp - 4) * Epi::kStageBytesPerWarp;
    const int q = (warp - 4) & 3;       // TMEM lane quadrant
    const int half = (warp - 4) >> 2;   // column half of the tile
    const int et = q * 32 + lane;       // 0..127 = row inside the tile
    const int et2 = (warp - 4) * 32 + lane;  // 0..255 among the epilogue threads
    constexpr int kChunksPerHalf = BN / 64;
    int acc_stage = 0; uint32_t acc_phase = 0;
    // Per-tile scalars (column scales, row scale / LSE / label) are fetched one tile AHEAD so their
    // global-load latency hides behind the current tile's chunk loop; the column data is double
    // buffered in shared memory, so one 256-thread barrier per tile is enough.
    epi.init();
    if ((int)blockIdx.x < num_items) {
      const int tile = (int)blockIdx.x % num_tiles;
      const int m_blk = tile % gs.num_m_blk, n_blk = tile / gs.num_m_blk;
      epi.prefetch(n_blk, et2, m_blk * kBM + et, m_blk * kBM + et < gs.M);
    }
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const int tile = item % num_tiles, ksp = item / num_tiles;
      const int m_blk = tile % gs.num_m_blk, n_blk = tile / gs.num_m_blk;
      float* se = s_epi + acc_stage * (2 * BN);
      epi.tile_begin(se, et2);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const int row = m_blk * kBM + et;
      const bool row_ok = row < gs.M;
      epi.row_begin(row, row_ok);
      {
        const int nitem = item + gridDim.x;
        if (nitem < num_items) {
          const int ntile = nitem % num_tiles;
          const int nm = ntile % gs.num_m_blk, nn = ntile / gs.num_m_blk;
          epi.prefetch(nn, et2, nm * kBM + et, nm * kBM + et < gs.M);
        }
      }
      mbar_wait(&tfull_bar[acc_stage], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc_stage * BN;
#pragma unroll 1
      for (int c = half * kChunksPerHalf; c < (half + 1) * kChunksPerHalf; ++c) {
        float acc[32];
        tmem_ld32(taddr + c * 32, acc);
        tmem_ld_wait();
        epi.chunk(acc, n_blk * BN + c * 32, c * 32, se, row, row_ok);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc_stage]);
      epi.row_end(row, row_ok, m_blk, n_blk, ksp, se, et, half);
      if (++acc_stage == 2) { acc_stage = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
// 2-D tensor map over a row-major [outer, inner] matrix with a 128-byte-wide box.
int make_tmap(CUtensorMap* out, const void* ptr, bool fp32, uint64_t inner, uint64_t outer,
              uint64_t row_stride_elems, uint32_t box_inner, uint32_t box_outer,
              bool atom32 = false);

struct GemmOperand {
  const void* ptr[2];  // hi, lo (lo null for bf16)
  int rows;            // M or N extent
  int64_t ld;          // leading dimension in elements
  int mn_major;        // 0: [rows, K] row-major; 1: [K, rows] row-major
};

template <bool TF32X3, int BN>
int build_tmaps(TmapSet* tm, GemmShape* gs, const GemmOperand& A, const GemmOperand& B, int K,
                int k_splits) {
  using Cfg = GemmCfg<TF32X3, BN>;
  gs->M = A.rows; gs->N = B.rows; gs->K = K;
  gs->num_m_blk = (A.rows + kBM - 1) / kBM;
  gs->num_n_blk = (B.rows + BN - 1) / BN;
  gs->kblk_total = (K + Cfg::kBK - 1) / Cfg::kBK;
  if (k_splits < 1) k_splits = 1;
  if (k_splits > gs->kblk_total) k_splits = gs->kblk_total;
  gs->kblk_per_split = (gs->kblk_total + k_splits - 1) / k_splits;
  gs->k_splits = (gs->kblk_total + gs->kblk_per_split - 1) / gs->kblk_per_split;
  gs->a_mn = A.mn_major; gs->b_mn = B.mn_major;
  for (int part = 0; part < Cfg::kParts; ++part) {
    if (!A.mn_major) CE_TRY(make_tmap(&tm->a[part], A.ptr[part], TF32X3, K, A.rows, A.ld, Cfg::kBK, kBM));
    else CE_TRY(make_tmap(&tm->a[part], A.ptr[part], TF32X3, A.rows, K, A.ld, Cfg::kBK, Cfg::kBK, TF32X3));
    if (!B.mn_major) CE_TRY(make_tmap(&tm->b[part], B.ptr[part], TF32X3, K, B.rows, B.ld, Cfg::kBK, BN));
    else CE_TRY(make_tmap(&tm->b[part], B.ptr[part], TF32X3, B.rows, K, B.ld, Cfg::kBK, Cfg::kBK, TF32X3));
  }
  if (Cfg::kParts == 1) { tm->a[1] = tm->a[0]; tm->b[1] = tm->b[0]; }
  return CE_OK;
}

template <bool TF32X3, int BN, class Epi>
int launch_gemm(const GemmOperand& A, const GemmOperand& B, int K, int k_splits,
                const typename Epi::Params& ep, cudaStream_t st, int* items_out = nullptr) {
  using Cfg = GemmCfg<TF32X3, BN, Epi::kStageBytesPerWarp * 8>;
  TmapSet tm;
  GemmShape gs;
  CE_TRY((build_tmaps<TF32X3, BN>(&tm, &gs, A, B, K, k_splits)));
  const int items = gs.num_m_blk * gs.num_n_blk * gs.k_splits;
  if (items_out) *items_out = items;
  if (items == 0) return CE_OK;
  auto kern = umma_gemm_kernel<TF32X3, BN, Epi>;
  CE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmemBytes));
  int grid = items < num_sms() ? items : num_sms();
  kern<<<grid, kGemmThreads, Cfg::kSmemBytes, st>>>(tm, gs, ep);
  CE_LAUNCH_CHECK();
  return CE_OK;
}

}  // namespace ce
