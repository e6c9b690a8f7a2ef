// OT gradient contraction for plans too large for the streaming kernel (bf16; c4: 32 x 257 x 768, the c5
// sweep): dy = -W^t x + ay*y, dx = -W y + ax*x  (model_ot.py:81-83 -- IPOT is not
// differentiated through, so the plan W enters as a constant).
//
// HBM-bound by construction: x, y are read once and dx, dy written once, 2 (M+N) D e bytes per sample; the
// contraction is 32 MACs per output element, so the kernel is written around its instruction count
// (the mma.sync predecessor `ot_grad_kernel` issued 15 instructions per MMA: 4-byte fragment loads, 4-byte
// global stores with their address arithmetic, 16-byte cp.async per thread).  Here
//   * persistent CTAs walk (sample, row split, column range) items; 64-column slabs of [x rows | y rows] arrive
//     through a 3-stage ring by TMA (one 3-D box per operand, 128-byte swizzle, out-of-range rows read as
//     zeros) and leave by TMA stores from the same ring: dy is formed IN PLACE over the y slab, dx in a small
//     side buffer -- no per-thread global load or store in the loop;
//   * fragments come from ldmatrix (x4) and go back with stmatrix (x4); the plan's A fragments for the dy
//     product and the diag(ay) fragments live in registers for the whole item;
//   * the plan W (fp32, written by the solver) of the NEXT item is prefetched by one bulk copy while the current
//     item runs, then converted to bf16 (sign folded in) once per item.
// Phases per slab: dx units (all warps read every y row) | barrier | dy tiles in place | barrier | stores.
#include <algorithm>

#include "ce_common.cuh"
#include "ot_fused.cuh"

namespace ce {

int make_tmap3d_bf16(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t rows, uint64_t batch,
                     uint64_t row_stride_elems, uint64_t batch_stride_elems, uint32_t box_inner, uint32_t box_rows);

namespace {

constexpr int kSlabCols = 64;      // bf16 columns per slab = one 128-byte swizzle row
constexpr int kRowB = 128;         // bytes of one staged row
constexpr int kMaxRing = 8;      // slab ring depth: as many stages as shared memory holds (plan_wide)

struct WideGradParams {
  int B, M, N, D;
  int tiles_per_cta;   // 16-row tiles of image nodes per row split
  int nsplit;          // row splits per sample (dx partials meet in dx_acc when > 1)
  int dsplit;          // column ranges per sample
  int nslab;           // D / 64
  int ybox, nybox;     // rows per y TMA box, boxes per stage
  int ring;            // stages of the slab ring
  int dbg;             // CE_OT_WIDE_DBG (timing aid): 1 = no dx phase, 2 = no dy phase, 3 = copy only (garbage out)
  const float* W;      // [B, N, MP] fp32 (scaled plan, model_ot.py:83 backward)
  const float* ax;     // [B, MP]
  const float* ay;     // [B, Nld]
  int Nld;
  float* dx_acc;       // [B, M, D] fp32 when nsplit > 1
};

__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(m), "r"(src),
               "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_load_3d_u32(uint32_t dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                                int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::
          "r"(dst),
      "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t* r, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t* r, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2_t(uint32_t* r, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
__device__ __forceinline__ void stsm_x4(uint32_t addr, const uint32_t* r) {
  asm volatile("stmatrix.sync.aligned.m8n8.x4.shared.b16 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3])
               : "memory");
}
__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float bf16_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

// byte offset of 16-byte chunk `chunk` of row `row` inside a 128-byte-swizzled tile whose base is 1024-byte aligned
__device__ __forceinline__ uint32_t swz(int row, int chunk) {
  return (uint32_t)(row * kRowB + ((chunk ^ (row & 7)) << 4));
}

struct WideLayout {
  uint32_t stage_bytes, ring, dxout, wf, ayf, axf, axs, wb, bars, total;
};
__host__ __device__ inline WideLayout wide_layout(int MP, int rows, int kRing) {
  WideLayout L;
  L.stage_bytes = (uint32_t)(MP + rows) * kRowB;
  uint32_t off = 0;
  L.ring = off; off += kRing * L.stage_bytes;
  L.dxout = off; off += 3u * MP * kRowB;
  L.wf = off; off += (uint32_t)rows * MP * 4;
  L.ayf = off; off += (uint32_t)rows * 4 + 64;
  L.axf = off; off += (uint32_t)MP * 4;
  L.axs = off; off += (uint32_t)MP * 4;
  L.wb = off; off += (uint32_t)rows * (MP + 8) * 2;
  L.bars = off; off += 8 * (kMaxRing + 1);
  L.total = off;
  return L;
}

// KT > 0: the dx product keeps the plan in REGISTERS (K-steps 0..KT-1 of 16 text rows x 16 image rows per warp,
// warp = (16 text rows, 16 columns)): shared-memory wavefronts per slab 1632 -> 544 (the first version was bound
// by the shared-memory pipe: 6 wavefronts per MMA).  Needs MP <= 32 and at most KT row tiles per CTA.
template <int MP, int TPW, int KT>
__global__ void __launch_bounds__(KT > 0 ? 320 : (TPW == 1 ? 576 : 512), 1)
ot_wide_grad_kernel(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmy,
                    const __grid_constant__ CUtensorMap tmdx, const __grid_constant__ CUtensorMap tmdy,
                    const WideGradParams a) {
  constexpr int LDWB = MP + 8;          // bf16 plan row stride (elements): conflict-free ldmatrix rows
  constexpr int KS = MP / 16;           // k-steps of the dy product
  constexpr int NOUT = (MP / 16) * 8;   // dx output units (16 text rows x 8 columns) per slab
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (sbase - smem_u32(smem_raw));
  const int rows = a.tiles_per_cta * 16;
  const int kRing = a.ring;
  const WideLayout L = wide_layout(MP, rows, kRing);
  float* Wf = reinterpret_cast<float*>(gbase + L.wf);
  float* ayf = reinterpret_cast<float*>(gbase + L.ayf);
  float* axf = reinterpret_cast<float*>(gbase + L.axf);
  float* axs = reinterpret_cast<float*>(gbase + L.axs);
  __nv_bfloat16* Wb = reinterpret_cast<__nv_bfloat16*>(gbase + L.wb);
  uint64_t* full = reinterpret_cast<uint64_t*>(gbase + L.bars);   // [kRing]
  uint64_t* wbar = full + kMaxRing;
  const uint32_t wb_u32 = sbase + L.wb;

  const int tid = threadIdx.x, NT = blockDim.x, w = tid >> 5, nw = NT >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int per_sample = a.nsplit * a.dsplit;
  const int items = a.B * per_sample;
  const int ntiles = (a.N + 15) / 16;

  auto geom = [&](int it, int& b, int& ns, int& c0, int& c1) {
    b = it / per_sample;
    const int rem = it - b * per_sample;
    ns = rem / a.dsplit;
    const int dh = rem - ns * a.dsplit;
    c0 = dh * a.nslab / a.dsplit;
    c1 = (dh + 1) * a.nslab / a.dsplit;
  };

  // Per-lane fragment offsets inside a swizzled tile, fixed for the whole kernel (row blocks are multiples of 8 rows,
  // so a tile / k-step only adds a multiple of 2048 bytes).  Pinned to registers: recomputing the XOR swizzle at
  // every ldmatrix was a third of the dy phase's instructions.
  const int q4 = lane >> 3;
  uint32_t lo4[4], so4[4];
#pragma unroll
  for (int j2 = 0; j2 < 4; ++j2) {
    lo4[j2] = swz(lane & 15, 2 * j2 + (lane >> 4));                      // B fragments: 16 rows x 2 chunks
    so4[j2] = swz((lane & 7) + (q4 & 1) * 8, 2 * j2 + (q4 >> 1));        // stmatrix of an accumulator pair
    asm volatile("" : "+r"(lo4[j2]), "+r"(so4[j2]));
  }

  if (tid == 0) {
    for (int s = 0; s < kRing; ++s) mbar_init(&full[s], 1);
    mbar_init(wbar, 1);
    mbar_fence_init();
  }
  __syncthreads();

  // ---- producer state (thread 0 only): the slab after the last one issued ----------------------------
  int l_it = blockIdx.x, l_b = 0, l_ns = 0, l_c = 0, l_c1 = 0;
  if (l_it < items) geom(l_it, l_b, l_ns, l_c, l_c1);
  auto issue_slab = [&](uint32_t q) {   // thread 0
    if (l_it >= items) return;
    const uint32_t s = q % kRing;
    const uint32_t st = sbase + L.ring + s * L.stage_bytes;
    mbar_expect_tx(&full[s], L.stage_bytes);
    tma_load_3d_u32(st, &tmx, &full[s], l_c * kSlabCols, 0, l_b);
    for (int k = 0; k < a.nybox; ++k)
      tma_load_3d_u32(st + (uint32_t)(MP + k * a.ybox) * kRowB, &tmy, &full[s], l_c * kSlabCols,
                      l_ns * rows + k * a.ybox, l_b);
    if (++l_c == l_c1) {
      l_it += gridDim.x;
      if (l_it < items) geom(l_it, l_b, l_ns, l_c, l_c1);
    }
  };
  auto issue_plan = [&](int it) {       // thread 0: W, ax, ay of item `it` into the fp32 staging buffers
    int b, ns, c0, c1;
    geom(it, b, ns, c0, c1);
    const int row0 = ns * rows;
    const int rv = min(rows, a.N - row0);
    const uint32_t wbytes = (uint32_t)rv * MP * 4, abytes = (uint32_t)min(rows, a.Nld - row0) * 4;
    mbar_expect_tx(wbar, wbytes + abytes + MP * 4);
    bulk_load(Wf, a.W + ((int64_t)b * a.N + row0) * MP, wbytes, wbar);
    bulk_load(ayf, a.ay + (int64_t)b * a.Nld + row0, abytes, wbar);
    bulk_load(axf, a.ax + (int64_t)b * MP, MP * 4, wbar);
  };
  if (tid == 0 && blockIdx.x < items) {
    issue_plan(blockIdx.x);
    for (int s = 0; s + 2 < kRing; ++s) issue_slab(s);
  }

  uint32_t q = 0;        // slabs consumed so far (ring position)
  int item_k = 0;
  for (int it = blockIdx.x; it < items; it += gridDim.x, ++item_k) {
    int b, ns, c0, c1;
    geom(it, b, ns, c0, c1);
    const int row0 = ns * rows;
    const int my_tiles = min(a.tiles_per_cta, ntiles - ns * a.tiles_per_cta);
    const int rows_valid = min(rows, a.N - row0);

    // ---- item start: plan -> bf16 (sign folded in), ax copy, fragments --------------------------------
    mbar_wait(wbar, item_k & 1);
    for (int idx = tid; idx < rows * (MP / 4); idx += NT) {
      const int r = idx / (MP / 4), cq = (idx - r * (MP / 4)) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < rows_valid) v = *reinterpret_cast<const float4*>(Wf + r * MP + cq);
      *reinterpret_cast<uint2*>(Wb + r * LDWB + cq) = make_uint2(pack2(-v.x, -v.y), pack2(-v.z, -v.w));
    }
    if (tid < MP) axs[tid] = axf[tid];
    uint32_t dgh[TPW][2], dgl[TPW][2];   // diag(ay) A fragments: registers 0 and 3 (1, 2 are zero)
#pragma unroll
    for (int i = 0; i < TPW; ++i) {
      const int r = (w + i * nw) * 16 + g;
      const float ay0 = r < rows_valid ? ayf[r] : 0.f, ay1 = r + 8 < rows_valid ? ayf[r + 8] : 0.f;
      const float h0 = bf16_round(ay0), h1 = bf16_round(ay1);
      const bool e0 = (2 * t == g), e1 = (2 * t + 1 == g);
      dgh[i][0] = pack2(e0 ? h0 : 0.f, e1 ? h0 : 0.f);
      dgh[i][1] = pack2(e0 ? h1 : 0.f, e1 ? h1 : 0.f);
      dgl[i][0] = pack2(e0 ? ay0 - h0 : 0.f, e1 ? ay0 - h0 : 0.f);
      dgl[i][1] = pack2(e0 ? ay1 - h1 : 0.f, e1 ? ay1 - h1 : 0.f);
    }
    __syncthreads();
    if (tid == 0 && it + (int)gridDim.x < items) issue_plan(it + gridDim.x);
    // A fragments of (-W^t) for the dy product (rows = image nodes of a tile, k = text nodes): per-lane address of
    // tile 0; the slab loop re-reads them (4 ldmatrix per tile) instead of holding 16 registers across the item
    const uint32_t awo = wb_u32 + (uint32_t)((((lane & 7) + ((lane >> 3) & 1) * 8) * LDWB + (lane >> 4) * 8) * 2);
    // register-resident plan for the dx product: B fragments (k = image rows, n = 16 text rows of this warp)
    constexpr int KTR = KT > 0 ? KT : 1;
    uint32_t wreg[KTR][4];
    float axr[2][2];
    const int mg = w >> 2, cb = w & 3;    // text-row group (16 rows), column block (16 columns) of this warp
    const bool dx_warp = KT > 0 && w < (MP / 16) * 4;
    if constexpr (KT > 0) {
      if (dx_warp) {
#pragma unroll
        for (int ks = 0; ks < KT; ++ks) {
          if (ks < my_tiles)
            ldsm_x4_t(wreg[ks], wb_u32 + (uint32_t)(((ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDWB + mg * 16 +
                                                      (lane >> 4) * 8) * 2));
          else wreg[ks][0] = wreg[ks][1] = wreg[ks][2] = wreg[ks][3] = 0u;
        }
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          axr[nt][0] = axs[mg * 16 + nt * 8 + 2 * t];
          axr[nt][1] = axs[mg * 16 + nt * 8 + 2 * t + 1];
        }
      }
    }

    for (int c = c0; c < c1; ++c, ++q) {
      const uint32_t s = q % kRing;
      const uint32_t st = sbase + L.ring + s * L.stage_bytes;       // x rows, then y rows
      const uint32_t sy = st + MP * kRowB;
      const uint32_t dxo = sbase + L.dxout + (q % 3) * (MP * kRowB);
      mbar_wait(&full[s], (q / kRing) & 1);

      // ---- phase A: dx, K = this CTA's image rows ------------------------------------------------------
      if constexpr (KT > 0) {
        if (dx_warp && !(a.dbg & 1)) {
          // D[d][m] = sum_n y[n][d] (-W)[m][n]: A = y^t out of the slab (transposed ldmatrix), B = the plan registers
          float acc[2][2][4];
#pragma unroll
          for (int e = 0; e < 2; ++e)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
              for (int cc = 0; cc < 4; ++cc) acc[e][nt][cc] = 0.f;
          uint32_t ao = sy + swz((lane & 7) + (q4 >> 1) * 8, 2 * cb + (q4 & 1));
          // K-steps past this item's tiles multiply zero plan registers; their addresses stay inside the stage
          const uint32_t kmax = (uint32_t)(a.tiles_per_cta - 1) * 2048u;
#pragma unroll
          for (int ks = 0; ks < KT; ++ks) {
            uint32_t aq[4];
            ldsm_x4_t(aq, ao + min((uint32_t)ks * 2048u, kmax));
            mma16816(acc[ks & 1][0], aq, wreg[ks][0], wreg[ks][1]);
            mma16816(acc[ks & 1][1], aq, wreg[ks][2], wreg[ks][3]);
          }
          // fragment (d = g / g + 8, m = 2t, 2t + 1) of text-row block nt; x comes in and dx goes out transposed
          const uint32_t xo = swz(mg * 16 + (q4 >> 1) * 8 + (lane & 7), 2 * cb + (q4 & 1));
          if (a.nsplit == 1) {
            uint32_t xq[4], oq[4];
            ldsm_x4_t(xq, st + xo);
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
              for (int db = 0; db < 2; ++db) {
                const uint32_t xv = xq[nt * 2 + db];
                oq[nt * 2 + db] = pack2(fmaf(axr[nt][0], bf_lo(xv), acc[0][nt][2 * db] + acc[1][nt][2 * db]),
                                        fmaf(axr[nt][1], bf_hi(xv), acc[0][nt][2 * db + 1] + acc[1][nt][2 * db + 1]));
              }
            asm volatile("stmatrix.sync.aligned.m8n8.x4.trans.shared.b16 [%0], {%1,%2,%3,%4};" ::"r"(dxo + xo),
                         "r"(oq[0]), "r"(oq[1]), "r"(oq[2]), "r"(oq[3])
                         : "memory");
          } else {
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
              for (int cc = 0; cc < 4; ++cc) {
                const int m = mg * 16 + nt * 8 + 2 * t + (cc & 1), d = c * kSlabCols + cb * 16 + g + (cc >> 1) * 8;
                if (m < a.M) atomicAdd(a.dx_acc + ((int64_t)b * a.M + m) * a.D + d, acc[0][nt][cc] + acc[1][nt][cc]);
              }
          }
        }
      } else {
      for (int u = w; u < NOUT; u += nw) {
        const int mi = u >> 3, j = u & 7;
        float acc0[4] = {0.f, 0.f, 0.f, 0.f}, acc1[4] = {0.f, 0.f, 0.f, 0.f};
        // A = W[m][n] out of Wb[n][m]: four transposed 8x8 blocks (m lo/hi x k lo/hi)
        const int blk = lane >> 3, br = lane & 7;
        uint32_t wa = wb_u32 + (uint32_t)(((br + (blk >> 1) * 8) * LDWB + 16 * mi + (blk & 1) * 8) * 2);
        int ks = 0;
        for (; ks + 1 < my_tiles; ks += 2) {
          uint32_t a0[4], a1[4], bq[4];
          ldsm_x4_t(a0, wa);
          ldsm_x4_t(a1, wa + 16 * LDWB * 2);
          ldsm_x4_t(bq, sy + swz(ks * 16 + lane, j));   // lanes 0-15: k-step ks, lanes 16-31: k-step ks + 1
          mma16816(acc0, a0, bq[0], bq[1]);
          mma16816(acc1, a1, bq[2], bq[3]);
          wa += 32 * LDWB * 2;
        }
        if (ks < my_tiles) {
          uint32_t a0[4], bq[2];
          ldsm_x4_t(a0, wa);
          ldsm_x2_t(bq, sy + swz(ks * 16 + (lane & 15), j));
          mma16816(acc0, a0, bq[0], bq[1]);
        }
        const int m = 16 * mi + g;
        if (a.nsplit == 1) {
          const uint32_t o0 = swz(m, j) + 4 * t, o1 = swz(m + 8, j) + 4 * t;
          const uint32_t u0 = lds32(st + o0), u1 = lds32(st + o1);
          const float ax0 = axs[m], ax1 = axs[m + 8];
          sts32(dxo + o0, pack2(fmaf(ax0, bf_lo(u0), acc0[0] + acc1[0]), fmaf(ax0, bf_hi(u0), acc0[1] + acc1[1])));
          sts32(dxo + o1, pack2(fmaf(ax1, bf_lo(u1), acc0[2] + acc1[2]), fmaf(ax1, bf_hi(u1), acc0[3] + acc1[3])));
        } else {
          float* dacc = a.dx_acc + ((int64_t)b * a.M) * a.D + c * kSlabCols + 8 * j + 2 * t;
          if (m < a.M) {
            atomicAdd(dacc + (int64_t)m * a.D, acc0[0] + acc1[0]);
            atomicAdd(dacc + (int64_t)m * a.D + 1, acc0[1] + acc1[1]);
          }
          if (m + 8 < a.M) {
            atomicAdd(dacc + (int64_t)(m + 8) * a.D, acc0[2] + acc1[2]);
            atomicAdd(dacc + (int64_t)(m + 8) * a.D + 1, acc0[3] + acc1[3]);
          }
        }
      }
      }
      // the stores of slab q - 2 have left shared memory (those of slab q - 1 may still be draining: waiting for
      // them here tied every slab to the drain time of the previous one)
      if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      __syncthreads();                       // every read of the y slab is done; stage (q - 2) % ring is free
      if (tid == 0) issue_slab(q + kRing - 2);

      // ---- phase B: dy tiles in place: (-W^t) x + diag(ay) y -------------------------------------------
#pragma unroll
      for (int i = 0; i < TPW; ++i) {
        const int tile = w + i * nw;
        if (tile < my_tiles && !(a.dbg & 2)) {
          float acc[8][4];
#pragma unroll
          for (int jj = 0; jj < 8; ++jj)
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) acc[jj][cc] = 0.f;
#pragma unroll
          for (int ks = 0; ks < KS; ++ks) {
            uint32_t aw[4];
            ldsm_x4(aw, awo + (uint32_t)((tile * 16 * LDWB + ks * 16) * 2));
#pragma unroll
            for (int j2 = 0; j2 < 4; ++j2) {
              uint32_t bq[4];
              ldsm_x4_t(bq, st + lo4[j2] + ks * 2048);
              mma16816(acc[2 * j2], aw, bq[0], bq[1]);
              mma16816(acc[2 * j2 + 1], aw, bq[2], bq[3]);
            }
          }
          const uint32_t dh[4] = {dgh[i][0], 0u, 0u, dgh[i][1]}, dl[4] = {dgl[i][0], 0u, 0u, dgl[i][1]};
#pragma unroll
          for (int j2 = 0; j2 < 4; ++j2) {
            uint32_t bq[4];
            ldsm_x4_t(bq, sy + lo4[j2] + tile * 2048);
            mma16816(acc[2 * j2], dh, bq[0], bq[1]);
            mma16816(acc[2 * j2], dl, bq[0], bq[1]);
            mma16816(acc[2 * j2 + 1], dh, bq[2], bq[3]);
            mma16816(acc[2 * j2 + 1], dl, bq[2], bq[3]);
          }
#pragma unroll
          for (int j2 = 0; j2 < 4; ++j2) {
            const uint32_t r4[4] = {pack2(acc[2 * j2][0], acc[2 * j2][1]), pack2(acc[2 * j2][2], acc[2 * j2][3]),
                                    pack2(acc[2 * j2 + 1][0], acc[2 * j2 + 1][1]),
                                    pack2(acc[2 * j2 + 1][2], acc[2 * j2 + 1][3])};
            stsm_x4(sy + so4[j2] + tile * 2048, r4);
          }
        }
      }
      fence_proxy_async();                   // generic writes (dy in place, dx side buffer) -> TMA stores
      __syncthreads();
      if (tid == 0) {
        if (a.nsplit == 1) tma_store_3d(&tmdx, dxo, c * kSlabCols, 0, b);
        for (int k = 0; k < a.nybox; ++k)
          if (k * a.ybox < rows_valid)
            tma_store_3d(&tmdy, sy + (uint32_t)(k * a.ybox) * kRowB, c * kSlabCols, row0 + k * a.ybox, b);
        tma_store_commit();
      }
    }
  }
  if (tid == 0) tma_store_wait_all();
}

// ------------------------------------------------------------------------------------------
// Cost contraction for the same plans: S[n][m] = <y_n, x_m> (fp32, [B, N, MP]) and the rows' sums of squares
// (model_ot.py:8-18 before the normalisation, which the solver applies).  Reads x, y once.  Same TMA ring as
// above, but nothing is stored from it, so there is no CTA-wide barrier at all: one extra warp produces (and takes
// the text rows' sums of squares from each slab), the tile warps release a stage through an mbarrier.
// ------------------------------------------------------------------------------------------
struct WideCostParams {
  int B, M, N, D;
  int tiles_per_cta, nsplit, nslab, ybox, nybox, ring;
  float* S;      // [B, N, MP]
  float* nx2;    // [B, MP]
  float* ny2;    // [B, Nld]
  int Nld;
};

template <int MP, int TPW>
__global__ void __launch_bounds__(TPW == 1 ? 576 : 352, 1)
ot_wide_cost_kernel(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmy,
                    const WideCostParams a) {
  constexpr int NJ = MP / 8;            // 8-column blocks of text nodes
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (sbase - smem_u32(smem_raw));
  const int rows = a.tiles_per_cta * 16;
  const uint32_t stage_bytes = (uint32_t)(MP + rows) * kRowB;
  const int R = a.ring;
  uint64_t* full = reinterpret_cast<uint64_t*>(gbase + (size_t)R * stage_bytes);
  uint64_t* empty = full + kMaxRing;
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int nw = (blockDim.x >> 5) - 1;          // tile warps; warp nw produces
  const int items = a.B * a.nsplit;
  const int ntiles = (a.N + 15) / 16;

  if (tid == 0) {
    for (int s = 0; s < R; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], nw + 1); }
    mbar_fence_init();
  }
  __syncthreads();

  if (w == nw) {
    // ---- producer + text-row norms ---------------------------------------------------------------------
    int l_it = blockIdx.x, l_c = 0;
    auto issue = [&](uint32_t p) {             // one lane
      if (l_it >= items) return;
      const uint32_t s = p % R, st = sbase + s * stage_bytes;
      const int b = l_it / a.nsplit, ns = l_it - b * a.nsplit;
      mbar_expect_tx(&full[s], stage_bytes);
      tma_load_3d_u32(st, &tmx, &full[s], l_c * kSlabCols, 0, b);
      for (int k = 0; k < a.nybox; ++k)
        tma_load_3d_u32(st + (uint32_t)(MP + k * a.ybox) * kRowB, &tmy, &full[s], l_c * kSlabCols,
                        ns * rows + k * a.ybox, b);
      if (++l_c == a.nslab) { l_c = 0; l_it += gridDim.x; }
    };
    if (lane == 0)
      for (int p = 0; p + 1 < R; ++p) issue(p);
    uint32_t q = 0;
    for (int it = blockIdx.x; it < items; it += gridDim.x) {
      const int b = it / a.nsplit, ns = it - b * a.nsplit;
      float ss[MP > 32 ? 2 : 1] = {};
      for (int c = 0; c < a.nslab; ++c, ++q) {
        if (lane == 0 && q > 0) {              // stage of slab q - 1: free once every warp has released it
          const uint32_t pq = q - 1;
          mbar_wait(&empty[pq % R], (pq / R) & 1);
          issue(q + R - 1);
        } else if (lane == 0) {
          // first slab: nothing to release, the prologue filled R - 1 stages; stage R - 1 is free
          issue(R - 1);
        }
        __syncwarp();
        const uint32_t s = q % R, st = sbase + s * stage_bytes;
        mbar_wait(&full[s], (q / R) & 1);
#pragma unroll
        for (int h = 0; h < (MP > 32 ? 2 : 1); ++h) {
          const int m = lane + 32 * h;
          if (m < MP) {
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) {
              uint4 v;
              asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                           : "r"(st + swz(m, ch)));
              const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float lo = bf_lo(wv[e]), hi = bf_hi(wv[e]);
                ss[h] = fmaf(lo, lo, ss[h]);
                ss[h] = fmaf(hi, hi, ss[h]);
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
      }
      if (ns == 0) {
#pragma unroll
        for (int h = 0; h < (MP > 32 ? 2 : 1); ++h)
          if (lane + 32 * h < MP) a.nx2[(int64_t)b * MP + lane + 32 * h] = ss[h];
      }
    }
    return;
  }

  // ---- tile warps ------------------------------------------------------------------------------------------
  uint32_t ao[TPW];
#pragma unroll
  for (int i = 0; i < TPW; ++i) {
    const int tile = min(w + i * nw, a.tiles_per_cta - 1);
    // A fragment (16 image rows x 16 columns): row = lane & 7 (+8 for matrices 1, 3), chunk = 2 ks + (lane >> 4)
    ao[i] = (uint32_t)(MP + tile * 16 + (lane & 7) + ((lane >> 3) & 1) * 8);
  }
  uint32_t q = 0;
  for (int it = blockIdx.x; it < items; it += gridDim.x) {
    const int b = it / a.nsplit, ns = it - b * a.nsplit;
    const int row0 = ns * rows;
    const int my_tiles = min(a.tiles_per_cta, ntiles - ns * a.tiles_per_cta);
    const int rows_valid = min(rows, a.N - row0);
    float acc[TPW][NJ][4], yd[TPW][2][4];
#pragma unroll
    for (int i = 0; i < TPW; ++i) {
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) yd[i][0][cc] = yd[i][1][cc] = 0.f;
#pragma unroll
      for (int j = 0; j < NJ; ++j)
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) acc[i][j][cc] = 0.f;
    }
    for (int c = 0; c < a.nslab; ++c, ++q) {
      const uint32_t s = q % R, st = sbase + s * stage_bytes;
      mbar_wait(&full[s], (q / R) & 1);
      if (w < my_tiles) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          // B fragments: text rows m on the n axis, k = columns; one x4 covers two 8-row blocks
          uint32_t bf[NJ / 2][4];
#pragma unroll
          for (int p2 = 0; p2 < NJ / 2; ++p2)
            ldsm_x4(bf[p2], st + swz(p2 * 16 + (lane & 7) + (lane >> 4) * 8, 2 * ks + ((lane >> 3) & 1)));
#pragma unroll
          for (int i = 0; i < TPW; ++i) {
            if (i == 0 || w + i * nw < my_tiles) {
              uint32_t af[4];
              ldsm_x4(af, st + swz((int)ao[i], 2 * ks + (lane >> 4)));
              mma16816(yd[i][0], af, af[0], af[2]);      // rows x rows 0-7 of the same tile: diagonal = |y|^2
              mma16816(yd[i][1], af, af[1], af[3]);
#pragma unroll
              for (int p2 = 0; p2 < NJ / 2; ++p2) {
                mma16816(acc[i][2 * p2], af, bf[p2][0], bf[p2][1]);
                mma16816(acc[i][2 * p2 + 1], af, bf[p2][2], bf[p2][3]);
              }
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
    }
    // ---- item end: S rows and |y|^2 of this warp's tiles ---------------------------------------------------
    float* Sg = a.S + ((int64_t)b * a.N + row0) * MP;
#pragma unroll
    for (int i = 0; i < TPW; ++i) {
      const int tile = w + i * nw;
      if (tile < my_tiles) {
        const int r = tile * 16 + g;
        if (t == (g >> 1)) {   // this thread holds the diagonal elements (g, g) and (g + 8, g + 8)
          const float s0 = (g & 1) ? yd[i][0][1] : yd[i][0][0];
          const float s1 = (g & 1) ? yd[i][1][3] : yd[i][1][2];
          if (r < rows_valid) a.ny2[(int64_t)b * a.Nld + row0 + r] = s0;
          if (r + 8 < rows_valid) a.ny2[(int64_t)b * a.Nld + row0 + r + 8] = s1;
        }
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          const int m = 8 * j + 2 * t;
          if (r < rows_valid) *reinterpret_cast<float2*>(Sg + (int64_t)r * MP + m) = make_float2(acc[i][j][0], acc[i][j][1]);
          if (r + 8 < rows_valid)
            *reinterpret_cast<float2*>(Sg + (int64_t)(r + 8) * MP + m) = make_float2(acc[i][j][2], acc[i][j][3]);
        }
      }
    }
  }
}

struct WidePlan {
  int MP, tiles_per_cta, nsplit, dsplit, ybox, nybox, tpw, nwarps, grid, ring, kt;
  size_t smem;
};

bool plan_wide(int B, int M, int N, int D, WidePlan* p) {
  if (M < 1 || M > 64 || N < 1 || N > 1024 || D < kSlabCols || D % kSlabCols != 0) return false;
  p->MP = M <= 16 ? 16 : (M <= 32 ? 32 : 64);
  const int ntiles = (N + 15) / 16;
  constexpr size_t kSmemMax = 226 * 1024;
  int cap = std::min(ntiles, 32);
  while (cap > 1 && wide_layout(p->MP, cap * 16, 4).total + 1024 > kSmemMax) --cap;
  p->nsplit = (ntiles + cap - 1) / cap;
  p->tiles_per_cta = (ntiles + p->nsplit - 1) / p->nsplit;
  const int rows = p->tiles_per_cta * 16;
  // ring depth: the loads in flight per SM hide the HBM latency (3 stages left ~1.5 slabs of lead: 3.8 TB/s at c4);
  // small plans share the SM between several CTAs instead of growing one ring
  p->ring = 4;   // the loop reloads the stage of slab q - 2: two slabs of lead need four stages
  while (p->ring < 7 && wide_layout(p->MP, rows, p->ring + 1).total + 1024 <= (rows >= 128 ? kSmemMax : 72 * 1024)) ++p->ring;
  {   // tuning aid: CE_OT_WIDE_RING forces the ring depth (when it fits)
    static const int ring_env = [] { const char* e = getenv("CE_OT_WIDE_RING"); return e != nullptr ? atoi(e) : 0; }();
    if (ring_env >= 3 && ring_env <= kMaxRing && wide_layout(p->MP, rows, ring_env).total + 1024 <= kSmemMax) p->ring = ring_env;
  }
  p->smem = wide_layout(p->MP, rows, p->ring).total + 1024;
  // TMA boxes hold at most 256 rows; two boxes of rows / 2 (a multiple of 8 rows: the second box stays
  // 1024-byte aligned) otherwise
  if (rows <= 256) { p->ybox = rows; p->nybox = 1; }
  else { p->ybox = rows / 2; p->nybox = 2; }
  // one warp per 16-row tile up to 18 tiles (c4: 17 warps, every phase balanced; 9 warps with two tiles each left
  // 2.25 warps per scheduler and 9 cycles between issues), two tiles per warp beyond
  static const int tpw1_max = [] { const char* e = getenv("CE_OT_WIDE_TPW1"); return e != nullptr ? atoi(e) : 18; }();
  p->tpw = p->tiles_per_cta > tpw1_max ? 2 : 1;
  p->nwarps = std::max(4, (p->tiles_per_cta + p->tpw - 1) / p->tpw);
  // register-resident plan for the dx product (KT = 17 or 20 K-steps): 9..20 row tiles, MP <= 32; two tiles per warp, and at least
  // the (MP / 16) x 4 warps of the dx phase
  static const bool res_on = [] { const char* e = getenv("CE_OT_WIDE_RES"); return e == nullptr || atoi(e) != 0; }();
  p->kt = 0;
  if (res_on && p->MP <= 32 && p->tiles_per_cta >= 9 && p->tiles_per_cta <= 20) {
    p->kt = p->tiles_per_cta <= 17 ? 17 : 20;
    p->tpw = 2;
    p->nwarps = std::max((p->tiles_per_cta + 1) / 2, (p->MP / 16) * 4);
  }
  int ctas_per_sm = std::max(1, std::min((int)((227 * 1024) / p->smem), 2048 / (p->nwarps * 32)));
  ctas_per_sm = std::min(ctas_per_sm, 4);
  const int slots = num_sms() * ctas_per_sm;
  // column ranges: fine enough that the item count divides evenly over the persistent CTAs, at least 3 slabs each
  const int nslab = D / kSlabCols;
  int best = 1;
  double best_waste = 1e9;
  for (int ds = 1; ds <= nslab; ++ds) {
    if (nslab % ds != 0 || (nslab / ds < 3 && ds > 1)) continue;
    const int64_t items = (int64_t)B * p->nsplit * ds;
    const int64_t per = (items + slots - 1) / slots;
    const double waste = (double)(per * slots) / (double)items;
    if (waste < best_waste - 0.02) { best_waste = waste; best = ds; }
  }
  p->dsplit = best;
  p->grid = (int)std::min<int64_t>((int64_t)B * p->nsplit * p->dsplit, slots);
  return true;
}

template <int MP, int TPW, int KT = 0>
int launch_cfg(const CUtensorMap& tx, const CUtensorMap& ty, const CUtensorMap& tdx, const CUtensorMap& tdy,
               const WideGradParams& a, const WidePlan& p, cudaStream_t st) {
  auto kern = ot_wide_grad_kernel<MP, TPW, KT>;
  CE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
  kern<<<p.grid, p.nwarps * 32, p.smem, st>>>(tx, ty, tdx, tdy, a);
  CE_LAUNCH_CHECK();
  return CE_OK;
}

}  // namespace

// Plans of more than 32 text rows keep the plan in shared memory (no register-resident dx product): up to 14 row
// tiles the mma.sync kernels of ot_kernels.cu are faster there (c5 sweep, batch 512, 50 iterations: 64 x 197 x 768
// 517 us against 680 us; from 64 x 257 on the TMA kernels win, 843 against 921 us).  CE_OT_WIDE=2 lifts the gate.
static bool wide_pays(int M, int N) {
  static const int mode = [] { const char* e = getenv("CE_OT_WIDE"); return e == nullptr ? 1 : atoi(e); }();
  return mode == 2 || !((M > 32 && N <= 224) || N <= 64);   // tiny plans (4 row tiles): 84 against 91 us at 16 x 50 x 768
}
bool ot_wide_supported(int M, int N, int D, int dtype) {
  static const bool on = [] { const char* e = getenv("CE_OT_WIDE"); return e == nullptr || atoi(e) != 0; }();
  WidePlan p;
  return on && dtype == CE_BF16 && wide_pays(M, N) && plan_wide(1, M, N, D, &p);
}
int ot_wide_nsplit(int M, int N, int D) {
  WidePlan p;
  return plan_wide(1, M, N, D, &p) ? p.nsplit : 1;
}

int launch_ot_wide_grad(const OtWideGradArgs& g, cudaStream_t st) {
  WidePlan p;
  if (!plan_wide(g.B, g.M, g.N, g.D, &p)) return fail(CE_ERR_SHAPE, "OT wide gradient: unsupported shape");
  CUtensorMap tx, ty, tdx, tdy;
  const int rows = p.tiles_per_cta * 16;
  (void)rows;
  CE_TRY(make_tmap3d_bf16(&tx, g.txt, g.D, g.M, g.B, g.D, g.txt_bs, kSlabCols, p.MP));
  CE_TRY(make_tmap3d_bf16(&ty, g.img, g.D, g.N, g.B, g.D, g.img_bs, kSlabCols, p.ybox));
  CE_TRY(make_tmap3d_bf16(&tdx, g.dtxt, g.D, g.M, g.B, g.D, g.txt_bs, kSlabCols, p.MP));
  CE_TRY(make_tmap3d_bf16(&tdy, g.dimg, g.D, g.N, g.B, g.D, g.img_bs, kSlabCols, p.ybox));
  WideGradParams a{};
  a.B = g.B; a.M = g.M; a.N = g.N; a.D = g.D;
  a.tiles_per_cta = p.tiles_per_cta; a.nsplit = p.nsplit; a.dsplit = p.dsplit; a.nslab = g.D / kSlabCols;
  a.ybox = p.ybox; a.nybox = p.nybox; a.ring = p.ring;
  { static const int dbg = [] { const char* e = getenv("CE_OT_WIDE_DBG"); return e != nullptr ? atoi(e) : 0; }(); a.dbg = dbg; }
  a.W = g.W; a.ax = g.ax; a.ay = g.ay; a.Nld = g.Nld; a.dx_acc = g.dx_acc;
  if (p.nsplit > 1 && g.dx_acc == nullptr) return fail(CE_ERR_ARG, "OT wide gradient: dx accumulator missing");
  if (p.kt == 17 && p.MP == 16) return launch_cfg<16, 2, 17>(tx, ty, tdx, tdy, a, p, st);
  if (p.kt == 17 && p.MP == 32) return launch_cfg<32, 2, 17>(tx, ty, tdx, tdy, a, p, st);
  if (p.kt == 20 && p.MP == 16) return launch_cfg<16, 2, 20>(tx, ty, tdx, tdy, a, p, st);
  if (p.kt == 20 && p.MP == 32) return launch_cfg<32, 2, 20>(tx, ty, tdx, tdy, a, p, st);
  switch (p.MP * 10 + p.tpw) {
    case 161: return launch_cfg<16, 1>(tx, ty, tdx, tdy, a, p, st);
    case 162: return launch_cfg<16, 2>(tx, ty, tdx, tdy, a, p, st);
    case 321: return launch_cfg<32, 1>(tx, ty, tdx, tdy, a, p, st);
    case 322: return launch_cfg<32, 2>(tx, ty, tdx, tdy, a, p, st);
    case 641: return launch_cfg<64, 1>(tx, ty, tdx, tdy, a, p, st);
    default: return launch_cfg<64, 2>(tx, ty, tdx, tdy, a, p, st);
  }
}

namespace {
struct WideCostPlan {
  int MP, tiles_per_cta, nsplit, ybox, nybox, tpw, nwarps, grid, ring;
  size_t smem;
};
bool plan_wide_cost(int B, int M, int N, int D, WideCostPlan* p) {
  if (M < 1 || M > 64 || N < 1 || N > 1024 || D < kSlabCols || D % kSlabCols != 0) return false;
  p->MP = M <= 16 ? 16 : (M <= 32 ? 32 : 64);
  const int ntiles = (N + 15) / 16;
  const int cap = 20;   // two tiles per warp on at most 10 tile warps (+ the producer): 352 threads, no spills at MP = 64
  p->nsplit = (ntiles + cap - 1) / cap;
  p->tiles_per_cta = (ntiles + p->nsplit - 1) / p->nsplit;
  const int rows = p->tiles_per_cta * 16;
  const size_t stage = (size_t)(p->MP + rows) * kRowB;
  const size_t budget = rows >= 128 ? 226 * 1024 : 72 * 1024;
  p->ring = (int)std::min<size_t>(kMaxRing, (budget - 1024 - 256) / stage);
  if (p->ring < 3) return false;
  p->smem = (size_t)p->ring * stage + 256 + 1024;
  if (rows <= 256) { p->ybox = rows; p->nybox = 1; }
  else { p->ybox = rows / 2; p->nybox = 2; }
  p->tpw = p->tiles_per_cta > 16 ? 2 : 1;
  p->nwarps = std::max(4, (p->tiles_per_cta + p->tpw - 1) / p->tpw);
  const int threads = (p->nwarps + 1) * 32;
  int ctas_per_sm = std::max(1, std::min((int)((227 * 1024) / p->smem), 2048 / threads));
  ctas_per_sm = std::min(ctas_per_sm, 3);
  p->grid = (int)std::min<int64_t>((int64_t)B * p->nsplit, (int64_t)num_sms() * ctas_per_sm);
  return true;
}
template <int MP, int TPW>
int launch_cost_cfg(const CUtensorMap& tx, const CUtensorMap& ty, const WideCostParams& a, const WideCostPlan& p,
                    cudaStream_t st) {
  auto kern = ot_wide_cost_kernel<MP, TPW>;
  CE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
  kern<<<p.grid, (p.nwarps + 1) * 32, p.smem, st>>>(tx, ty, a);
  CE_LAUNCH_CHECK();
  return CE_OK;
}
}  // namespace

bool ot_wide_cost_supported(int M, int N, int D, int dtype) {
  static const bool on = [] { const char* e = getenv("CE_OT_WIDE_COST"); return e == nullptr || atoi(e) != 0; }();
  WideCostPlan p;
  return on && dtype == CE_BF16 && wide_pays(M, N) && plan_wide_cost(1, M, N, D, &p);
}

int launch_ot_wide_cost(const OtWideCostArgs& g, cudaStream_t st) {
  WideCostPlan p;
  if (!plan_wide_cost(g.B, g.M, g.N, g.D, &p)) return fail(CE_ERR_SHAPE, "OT wide cost: unsupported shape");
  CUtensorMap tx, ty;
  CE_TRY(make_tmap3d_bf16(&tx, g.txt, g.D, g.M, g.B, g.D, g.txt_bs, kSlabCols, p.MP));
  CE_TRY(make_tmap3d_bf16(&ty, g.img, g.D, g.N, g.B, g.D, g.img_bs, kSlabCols, p.ybox));
  WideCostParams a{};
  a.B = g.B; a.M = g.M; a.N = g.N; a.D = g.D;
  a.tiles_per_cta = p.tiles_per_cta; a.nsplit = p.nsplit; a.nslab = g.D / kSlabCols;
  a.ybox = p.ybox; a.nybox = p.nybox; a.ring = p.ring;
  a.S = g.S; a.nx2 = g.nx2; a.ny2 = g.ny2; a.Nld = g.Nld;
  switch (p.MP * 10 + p.tpw) {
    case 161: return launch_cost_cfg<16, 1>(tx, ty, a, p, st);
    case 162: return launch_cost_cfg<16, 2>(tx, ty, a, p, st);
    case 321: return launch_cost_cfg<32, 1>(tx, ty, a, p, st);
    case 322: return launch_cost_cfg<32, 2>(tx, ty, a, p, st);
    case 641: return launch_cost_cfg<64, 1>(tx, ty, a, p, st);
    default: return launch_cost_cfg<64, 2>(tx, ty, a, p, st);
  }
}

}  // namespace ce
