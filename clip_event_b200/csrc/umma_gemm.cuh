// Persistent, warp-specialised tcgen05 GEMM engine for sm_100a.
//
//   acc[128 x BN] (fp32, TMEM) = A_tile * B_tile^t   over a K range, A/B tiles staged by TMA into
//   128-byte-swizzled shared memory, `tcgen05.mma` issued by one elected lane, accumulators double
//   buffered in TMEM so the epilogue of tile i overlaps the main loop of tile i+1.
//
// Warp roles (384 threads): warp 0 = TMA producer, warp 1 = MMA issuer (both run their loops with
// all lanes converged, one elected lane issues), warp 2 = TMEM allocator, warps 4..11 = epilogue:
// warp w reads TMEM lanes 32*(w%4).. and the column half (w-4)/4 of the tile, so two warps share
// each row.  CG = 2 runs the same roles on a CTA pair: one 256 x BN tile per pair, the leader CTA
// issues `tcgen05.mma.cta_group::2` for both, each CTA stages half of the B tile.
//
// Operand precision: bf16 (kind::f16, one product) or fp32 carried as a (hi, lo) pair of tf32
// arrays with three products hi*hi + hi*lo + lo*hi (kind::tf32) -- fp32-level accuracy.
// Operand layouts: K-major ([rows, K] row-major) or MN-major ([K, rows] row-major), both through
// the canonical SWIZZLE_128B shared-memory layouts.
//
// The epilogue is a functor template parameter: row softmax statistics, on-the-fly softmax
// gradient (TMA-stored), or plain scaled store / split-K accumulate.
#pragma once

#include <stdlib.h>

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
  int raster_n;            // tile order: 0 = row units fastest, 1 = column blocks fastest
  long long* trace;        // -DCE_GEMM_TRACE builds only: per-tile clock64 stamps of CTA 0
};

// Pipeline tracing (tuning aid, compiled in with -DCE_GEMM_TRACE): CTA 0 records, per work item,
// clock64 at the hand-over points of its producer / MMA / epilogue roles: trace[(item_seq * 4 + role) * 8 + k].
#ifdef CE_GEMM_TRACE
#define CE_TRACE(role, k)                                                                     \
  do {                                                                                        \
    if (gs.trace != nullptr && blockIdx.x == 0 && seq < 64)                                   \
      gs.trace[(seq * 4 + (role)) * 8 + (k)] = clock64();                                     \
  } while (0)
#else
#define CE_TRACE(role, k) do { } while (0)
#endif

// EPI_STAGE: bytes of epilogue staging smem.  CG: CTAs per tile group -- 1, or 2 = a CTA pair on one
// 256 x BN tile (`cta_group::2`): each CTA stages its 128 rows of A and HALF of the B tile, so the
// shared-memory traffic per flop halves.
template <bool TF32X3, int BN, int EPI_STAGE = 0, int CG = 1>
struct GemmCfg {
  static constexpr int kElemBytes = TF32X3 ? 4 : 2;
  static constexpr int kBK = kSwizzleBytes / kElemBytes;     // elements of K per stage: 32 / 64
  static constexpr int kUmmaK = 32 / kElemBytes;             // K per instruction: 8 / 16
  static constexpr int kParts = TF32X3 ? 2 : 1;              // hi, lo
  static constexpr int kABytes = kBM * kSwizzleBytes;        // one A part per stage
  static constexpr int kBRows = BN / CG;                     // rows of the B tile staged by one CTA
  static constexpr int kBBytes = kBRows * kSwizzleBytes;
  static constexpr int kStageBytes = kParts * (kABytes + kBBytes);
  static constexpr int kEpiFloats = 2 * BN;                  // per-tile column data, double buffered
  static constexpr int kMaxSmem = 232448;                    // 227 KB opt-in limit per CTA
  static constexpr int kBudget = kMaxSmem - EPI_STAGE - (int)sizeof(float) * kEpiFloats - 256 /*barriers*/;
  static constexpr int kMaxStages = CG == 2 ? 8 : 6;
  static constexpr int kStages = kBudget / kStageBytes > kMaxStages ? kMaxStages : kBudget / kStageBytes;
  static constexpr int kEpiStageBytes = EPI_STAGE;
  static constexpr int kTmemCols = 2 * BN;                   // double-buffered accumulator
  // layout: stages | epilogue staging (1024-byte aligned like the stages) | column data | barriers;
  // no alignment slack: the dynamic shared-memory window itself is 1024-byte aligned (checked)
  static constexpr int kEpiStageOffset = kStages * kStageBytes;
  static constexpr int kEpiDataOffset = kEpiStageOffset + EPI_STAGE;
  static constexpr size_t kSmemBytes = (size_t)kEpiDataOffset + sizeof(float) * kEpiFloats + 256;
  static_assert(kStages >= 2, "pipeline needs at least two stages");
  static_assert(CG == 1 || (CG == 2 && !TF32X3), "CTA pairs are wired for the bf16 path only");
  static_assert(kTmemCols <= 512 && (kTmemCols & (kTmemCols - 1)) == 0, "TMEM columns");
};

struct TmapSet {
  CUtensorMap a[2];  // hi, lo (lo unused for bf16)
  CUtensorMap b[2];
  CUtensorMap out;   // optional: output matrix for epilogues that store their tile with TMA
};

// Walks the work items of one worker (item, item + W, ...) without a division per tile:
// item = ksp * num_tiles + tile, tile = outer * n_inner + inner.  The inner (fastest) tile index is
// the row unit by default -- concurrent CTAs then share B tiles through L2 -- or the column block
// (`raster_n`) when there are only a few column blocks and the A operand is the big one.
struct TileWalk {
  int n_inner, num_tiles, W, w_in, w_out, raster_n;
  int item, tile, ksp, in, out, mu, nb;
  __host__ __device__ __forceinline__ void assign() { mu = raster_n ? out : in; nb = raster_n ? in : out; }
  __host__ __device__ __forceinline__ void init(int item0, int nmu, int nnb, int W_, int raster_n_) {
    raster_n = raster_n_;
    n_inner = raster_n ? nnb : nmu; num_tiles = nmu * nnb; W = W_;
    w_out = W / n_inner; w_in = W - w_out * n_inner;
    item = item0; ksp = item0 / num_tiles; tile = item0 - ksp * num_tiles;
    out = tile / n_inner; in = tile - out * n_inner;
    assign();
  }
  __host__ __device__ __forceinline__ void next() {
    item += W; tile += W; in += w_in; out += w_out;
    if (in >= n_inner) { in -= n_inner; ++out; }
    if (tile >= num_tiles) {   // next K split: rare, re-derive
      do { tile -= num_tiles; ++ksp; } while (tile >= num_tiles);
      out = tile / n_inner; in = tile - out * n_inner;
    }
    assign();
  }
};

// Epilogue interface (one object per epilogue thread; `et2` = 0..255 among the epilogue threads,
// `half` = which 128-column half of the tile this warp handles):
//   struct Epi { struct Params {...}; Params p; uint8_t* stage; const CUtensorMap* out_map;
//     static constexpr int kStageBytesPerWarp;     // shared-memory staging the kernel reserves per warp
//     void init();                                  // once per kernel
//     void prefetch(int n_blk, int et2, int row, bool row_ok);   // global loads for the NEXT tile
//     void tile_begin(float* s_epi, int et2);       // publish column data; a 256-thread barrier follows
//     void row_begin(int row, bool row_ok);
//     void chunk(const float* acc /*32 columns of this thread's row*/, int col0 /*global column*/,
//                int lcol0 /*tile-local column*/, const float* s_epi, int row, bool row_ok);
//     void row_end(int row, bool row_ok, int m_blk, int n_blk, int k_split, float* s_epi, int et, int half);
//     void finish();                                // once, after the last tile
//     static bool skip_all(const Params&);          // true (for the whole grid): the kernel returns at once
//   };

template <bool TF32X3, int BN, class Epi, int CG = 1>
__global__ void __launch_bounds__(kGemmThreads, 1)
umma_gemm_kernel(const __grid_constant__ TmapSet tm, const GemmShape gs, const typename Epi::Params ep) {
  using Cfg = GemmCfg<TF32X3, BN, Epi::kStageBytesPerWarp * 8, CG>;
  // plain pointer arithmetic on the shared array keeps the address space known to the compiler
  // (LDS/STS instead of generic LD/ST in the epilogues)
  // Device-side switch between two paths of a captured launch sequence (grid-uniform).  It may read only INPUTS of the
  // API call (logit_scale), which no kernel of the launch chain writes, so it is evaluated ahead of pdl_wait(); a
  // skipped launch still waits before it exits -- its successor's wait must imply the completion of everything
  // before it.
  if (Epi::skip_all(ep)) { pdl_wait(); return; }
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();   // swizzled TMA / UMMA tiles need 1024-byte alignment
  uint8_t* stage_base = smem;
  float* s_epi = reinterpret_cast<float*>(smem + Cfg::kEpiDataOffset);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_epi + Cfg::kEpiFloats);
  uint64_t* full_bar = bars;                       // [kStages]
  uint64_t* empty_bar = bars + Cfg::kStages;       // [kStages]
  uint64_t* tfull_bar = bars + 2 * Cfg::kStages;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;            // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  // [8 warps][kStageBytesPerWarp], 1024-byte aligned (swizzled TMA-store source)
  uint8_t* epi_stage = smem + Cfg::kEpiStageOffset;

  // The control warps run their loops with all 32 lanes converged and let one elected lane issue
  // the TMA / tcgen05 instructions: the operands are then provably warp-uniform and live in uniform
  // registers (issuing from a `lane == 0` branch costs an R2UR + election loop per instruction,
  // and the MMA issue path was as long as the MMAs themselves).
  const int warp = __shfl_sync(0xffffffffu, warp_id(), 0), lane = lane_id();
  // CG == 2: CTAs 2w and 2w+1 form worker w; CTA rank r owns rows [(2*unit + r) * 128, +128) of the
  // 256-row unit and stages columns [r * BN/2, +BN/2) of the B tile.  Rank 0 (the leader) issues the MMAs.
  const uint32_t cta_rank = CG == 2 ? __shfl_sync(0xffffffffu, cluster_ctarank(), 0) : 0u;
  const int worker = (int)blockIdx.x / CG, num_workers = (int)gridDim.x / CG;
  const int num_m_units = (gs.num_m_blk + CG - 1) / CG;
  const int num_tiles = num_m_units * gs.num_n_blk;
  const int num_items = num_tiles * gs.k_splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm.a[0]);
    tma_prefetch_desc(&tm.b[0]);
    if (TF32X3) { tma_prefetch_desc(&tm.a[1]); tma_prefetch_desc(&tm.b[1]); }
  }
  if (warp == 1 && lane == 0) {
    // pair mode: the leader's full barrier collects the TMA bytes of both CTAs (one local expect_tx
    // for the pair's bytes), its accumulator-empty barrier the epilogue warps of both CTAs
    for (int i = 0; i < Cfg::kStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 8 * CG); }
    mbar_fence_init();
  }
  if (warp == 2) {
    if constexpr (CG == 2) tmem_alloc_pair(tmem_slot, Cfg::kTmemCols);
    else tmem_alloc(tmem_slot, Cfg::kTmemCols);
  }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all();   // the peer's barriers must exist before anything lands on them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // everything above (barriers, TMEM, descriptor prefetch) touches no global data: it overlaps the tail of the
  // kernel before this one when the launch is programmatic
  pdl_wait();

  if (warp == 0) {
    // ================= TMA producer =================
    int stage = 0; uint32_t phase = 0;
    // pair mode: completion is signalled on the LEADER's full barrier (a shared::cluster address)
    auto load = [&](void* dst, const CUtensorMap* m, int st_idx, int c0, int c1) {
      if constexpr (CG == 2) tma_load_2d_pair(dst, m, map_to_cta(&full_bar[st_idx], 0), c0, c1);
      else tma_load_2d(dst, m, &full_bar[st_idx], c0, c1);
    };
    TileWalk tw;
    tw.init(worker, num_m_units, gs.num_n_blk, num_workers, gs.raster_n);
    for (int seq = 0; tw.item < num_items; tw.next(), ++seq) {
      const int ksp = tw.ksp;
      const int m_blk = tw.mu * CG + (int)cta_rank, n_blk = tw.nb;
      const int n0 = n_blk * BN + (int)cta_rank * Cfg::kBRows;   // first B row this CTA stages
      const int kb0 = ksp * gs.kblk_per_split;
      const int kb1 = min(kb0 + gs.kblk_per_split, gs.kblk_total);
      CE_TRACE(0, 0);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (kb == kb0) CE_TRACE(0, 1);
        if (elect_one()) {
        uint8_t* st = stage_base + (size_t)stage * Cfg::kStageBytes;
        // The peer never arrives on the leader's barrier: its bytes may land before the leader's
        // expect_tx (the transaction count goes negative for a moment), but the phase cannot
        // complete until the leader's arrival, and the peer cannot run a phase ahead because its
        // own empty barrier is released by the leader's commit.
        if (cta_rank == 0) mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes * CG);
#pragma unroll
        for (int part = 0; part < Cfg::kParts; ++part) {
          uint8_t* sa = st + part * Cfg::kABytes;
          uint8_t* sb = st + Cfg::kParts * Cfg::kABytes + part * Cfg::kBBytes;
          if (!gs.a_mn) {
            load(sa, &tm.a[part], stage, kb * Cfg::kBK, m_blk * kBM);
          } else {  // [K, M] row-major: boxes of (kBK rows of K) x (one 128-byte run of M)
#pragma unroll
            for (int c = 0; c < kBM / Cfg::kBK; ++c)
              load(sa + c * Cfg::kBK * kSwizzleBytes, &tm.a[part], stage, m_blk * kBM + c * Cfg::kBK,
                   kb * Cfg::kBK);
          }
          if (!gs.b_mn) {
            load(sb, &tm.b[part], stage, kb * Cfg::kBK, n0);
          } else {
#pragma unroll
            for (int c = 0; c < Cfg::kBRows / Cfg::kBK; ++c)
              load(sb + c * Cfg::kBK * kSwizzleBytes, &tm.b[part], stage, n0 + c * Cfg::kBK,
                   kb * Cfg::kBK);
          }
        }
        }
        __syncwarp();
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
      CE_TRACE(0, 2);
    }
    if constexpr (CG == 2) {
      // drain: every multicast commit aimed at this CTA's empty barriers must have landed before the
      // CTA may exit
      for (int i = 0; i < Cfg::kStages; ++i) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && cta_rank == 0) {
    // ================= MMA issuer (pair mode: the leader CTA only) =================
    const uint32_t idesc = umma_idesc(TF32X3, kBM * CG, BN, gs.a_mn != 0, gs.b_mn != 0);
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
    TileWalk tw;
    tw.init(worker, num_m_units, gs.num_n_blk, num_workers, gs.raster_n);
    for (int seq = 0; tw.item < num_items; tw.next(), ++seq) {
      const int ksp = tw.ksp;
      const int kb0 = ksp * gs.kblk_per_split;
      const int kb1 = min(kb0 + gs.kblk_per_split, gs.kblk_total);
      CE_TRACE(1, 0);
      mbar_wait(&tempty_bar[acc_stage], acc_phase ^ 1);
      tc_fence_after();
      CE_TRACE(1, 1);
      const uint32_t tmem_d = tmem_base + acc_stage * BN;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (kb == kb0) CE_TRACE(1, 2);
        if (kb == kb1 - 1) CE_TRACE(1, 3);
        if (elect_one()) {
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
          } else if constexpr (CG == 2) {
            umma_pair_bf16(tmem_d, da_hi, db_hi, idesc, first);
          } else {
            umma<false>(tmem_d, da_hi, db_hi, idesc, first);
          }
        }
        // frees the smem stage (in both CTAs of a pair) once these MMAs retire
        if constexpr (CG == 2) umma_commit_pair(&empty_bar[stage], 3);
        else umma_commit(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
      // accumulator ready for the epilogue
      if (elect_one()) {
        if constexpr (CG == 2) umma_commit_pair(&tfull_bar[acc_stage], 3);
        else umma_commit(&tfull_bar[acc_stage]);
      }
      __syncwarp();
      CE_TRACE(1, 4);
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
    epi.out_map = &tm.out;
    epi.init();
    TileWalk tw, nx;
    tw.init(worker, num_m_units, gs.num_n_blk, num_workers, gs.raster_n);
    nx = tw;
    if (worker < num_items) {
      const int m_blk = tw.mu * CG + (int)cta_rank;
      epi.prefetch(tw.nb, et2, m_blk * kBM + et, m_blk * kBM + et < gs.M);
    }
    for (int seq = 0; tw.item < num_items; tw = nx, ++seq) {
      const int ksp = tw.ksp;
      const int trole = 2 + half;   // warps 4 and 8 (lane 0) trace the two column halves of quadrant 0
#define CE_TRACE_E(k) do { if (q == 0 && lane == 0) CE_TRACE(trole, k); } while (0)
      CE_TRACE_E(0);
      const int m_blk = tw.mu * CG + (int)cta_rank, n_blk = tw.nb;
      const bool tile_ok = CG == 1 || m_blk < gs.num_m_blk;   // odd row-block count: the peer idles
      float* se = s_epi + acc_stage * BN;
      epi.tile_begin(se, et2);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      CE_TRACE_E(1);
      const int row = m_blk * kBM + et;
      const bool row_ok = row < gs.M;
      epi.row_begin(row, row_ok);
      nx.next();
      CE_TRACE_E(6);
      if (nx.item < num_items) {
        const int nm = nx.mu * CG + (int)cta_rank;
        epi.prefetch(nx.nb, et2, nm * kBM + et, nm * kBM + et < gs.M);
      }
      CE_TRACE_E(2);
      mbar_wait(&tfull_bar[acc_stage], acc_phase);
      tc_fence_after();
      CE_TRACE_E(3);
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc_stage * BN;
      if (tile_ok) {
#pragma unroll 1
        for (int c = half * kChunksPerHalf; c < (half + 1) * kChunksPerHalf; ++c) {
          float acc[32];
          tmem_ld32(taddr + c * 32, acc);
          tmem_ld_wait();
          epi.chunk(acc, n_blk * BN + c * 32, c * 32, se, row, row_ok);
        }
      }
      tc_fence_before();
      __syncwarp();
      CE_TRACE_E(4);
      if (lane == 0) {
        if constexpr (CG == 2) mbar_arrive_cluster(map_to_cta(&tempty_bar[acc_stage], 0));
        else mbar_arrive(&tempty_bar[acc_stage]);
      }
      if (tile_ok) epi.row_end(row, row_ok, m_blk, n_blk, ksp, se, et, half);
      CE_TRACE_E(5);
      if (++acc_stage == 2) { acc_stage = 0; acc_phase ^= 1; }
    }
    epi.finish();
  }

  tc_fence_before();
  if constexpr (CG == 2) {
    cluster_sync_all();   // neither CTA may release TMEM or exit while its peer still uses the pair
    if (warp == 2) tmem_dealloc_pair(tmem_base, Cfg::kTmemCols);
  } else {
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
// 2-D tensor map over a row-major [outer, inner] matrix.  swizzle: 0 = 128-byte (128-byte-wide
// box), 1 = 128-byte with 32-byte atoms (tf32 MN-major operands), 2 = 64-byte (64-byte-wide box).
int make_tmap(CUtensorMap* out, const void* ptr, bool fp32, uint64_t inner, uint64_t outer,
              uint64_t row_stride_elems, uint32_t box_inner, uint32_t box_outer,
              int swizzle = 0);

struct GemmOperand {
  const void* ptr[2];  // hi, lo (lo null for bf16)
  int rows;            // M or N extent
  int64_t ld;          // leading dimension in elements
  int mn_major;        // 0: [rows, K] row-major; 1: [K, rows] row-major
};

template <bool TF32X3, int BN, int CG = 1>
int build_tmaps(TmapSet* tm, GemmShape* gs, const GemmOperand& A, const GemmOperand& B, int K,
                int k_splits) {
  using Cfg = GemmCfg<TF32X3, BN, 0, CG>;
  gs->M = A.rows; gs->N = B.rows; gs->K = K;
  gs->num_m_blk = (A.rows + kBM - 1) / kBM;
  gs->num_n_blk = (B.rows + BN - 1) / BN;
  gs->kblk_total = (K + Cfg::kBK - 1) / Cfg::kBK;
  if (k_splits < 1) k_splits = 1;
  if (k_splits > gs->kblk_total) k_splits = gs->kblk_total;
  gs->kblk_per_split = (gs->kblk_total + k_splits - 1) / k_splits;
  gs->k_splits = (gs->kblk_total + gs->kblk_per_split - 1) / gs->kblk_per_split;
  gs->a_mn = A.mn_major; gs->b_mn = B.mn_major;
  // few column blocks over a tall A: run the column blocks of one row unit side by side, so that A
  // is fetched from HBM once (measured on G^t x img at c3: 616 MB of DRAM reads for a 302 MB G)
  gs->raster_n = (gs->num_n_blk > 1 && gs->num_n_blk * 8 <= gs->num_m_blk) ? 1 : 0;
  gs->trace = nullptr;
#ifdef CE_GEMM_TRACE
  {  // trace the CE_GEMM_TRACE_LAUNCH-th GEMM launch of this process into the buffer at CE_GEMM_TRACE_PTR
    static int launch_no = 0;
    const char* e = getenv("CE_GEMM_TRACE_PTR");
    const char* n = getenv("CE_GEMM_TRACE_LAUNCH");
    if (e != nullptr && ++launch_no == (n != nullptr ? atoi(n) : 1))
      gs->trace = reinterpret_cast<long long*>(strtoull(e, nullptr, 0));
  }
#endif
  for (int part = 0; part < Cfg::kParts; ++part) {
    if (!A.mn_major) CE_TRY(make_tmap(&tm->a[part], A.ptr[part], TF32X3, K, A.rows, A.ld, Cfg::kBK, kBM));
    else CE_TRY(make_tmap(&tm->a[part], A.ptr[part], TF32X3, A.rows, K, A.ld, Cfg::kBK, Cfg::kBK, TF32X3 ? 1 : 0));
    if (!B.mn_major) CE_TRY(make_tmap(&tm->b[part], B.ptr[part], TF32X3, K, B.rows, B.ld, Cfg::kBK, Cfg::kBRows));
    else CE_TRY(make_tmap(&tm->b[part], B.ptr[part], TF32X3, B.rows, K, B.ld, Cfg::kBK, Cfg::kBK, TF32X3 ? 1 : 0));
  }
  if (Cfg::kParts == 1) { tm->a[1] = tm->a[0]; tm->b[1] = tm->b[0]; }
  return CE_OK;
}

// bf16 [rows, cols] row-major output written by the epilogue through TMA in [32 x 64] boxes
struct GemmOut {
  void* ptr;
  int rows, cols;
  int64_t ld;
};

template <bool TF32X3, int BN, class Epi, int CG = 1>
int launch_gemm(const GemmOperand& A, const GemmOperand& B, int K, int k_splits,
                const typename Epi::Params& ep, cudaStream_t st, int* items_out = nullptr,
                const GemmOut* out = nullptr) {
  using Cfg = GemmCfg<TF32X3, BN, Epi::kStageBytesPerWarp * 8, CG>;
  TmapSet tm;
  GemmShape gs;
  CE_TRY((build_tmaps<TF32X3, BN, CG>(&tm, &gs, A, B, K, k_splits)));
  if (out != nullptr) CE_TRY(make_tmap(&tm.out, out->ptr, false, out->cols, out->rows, out->ld, 64, 32));
  else tm.out = tm.a[0];
  const int items = (gs.num_m_blk + CG - 1) / CG * gs.num_n_blk * gs.k_splits;
  if (items_out) *items_out = items;
  if (items == 0) return CE_OK;
  auto kern = umma_gemm_kernel<TF32X3, BN, Epi, CG>;
  CE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmemBytes));
  // a link of the programmatic-dependent-launch chain (ce_common.cuh): the kernel begins with pdl_wait()
  const bool pdl = pdl_use();
  if constexpr (CG == 1) {
    const int workers = items < num_sms() ? items : num_sms();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(workers);
    cfg.blockDim = dim3(kGemmThreads);
    cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    if (pdl) {
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
    }
    CE_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, tm, gs, ep));
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(num_sms() / CG * CG);
    cfg.blockDim = dim3(kGemmThreads);
    cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    // A persistent kernel must be fully co-resident: GPCs with an odd number of usable SMs leave
    // fewer than num_sms/2 slots for CTA pairs, and a pair that waits for a second wave would
    // serialise its whole share of the work.
    static thread_local int max_pairs = 0;
    if (max_pairs == 0) {
      int n = 0;
      CE_CUDA_TRY(cudaOccupancyMaxActiveClusters(&n, kern, &cfg));
      if (n < 1) return fail(CE_ERR_ARCH, "no CTA pair of the GEMM kernel fits on this device");
      max_pairs = n;
      if (getenv("CE_DEBUG")) fprintf(stderr, "clip_event_b200: %d co-resident CTA pairs\n", n);
    }
    const int workers = items < max_pairs ? items : max_pairs;
    cfg.gridDim = dim3(workers * CG);
    if (pdl) {
      attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[1].val.programmaticStreamSerializationAllowed = 1;
      cfg.numAttrs = 2;
    }
    CE_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, tm, gs, ep));
  }
  CE_LAUNCH_CHECK();
  pdl_mark();
  return CE_OK;
}

}  // namespace ce
