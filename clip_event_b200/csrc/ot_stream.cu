// OT graph-alignment loss as ONE persistent, warp-specialised STREAMING kernel (bf16, small plans).
// Reference behaviour: src/clip-event/model_ot.py:8-84 (cost_matrix_cosine -> ipot -> trace) and
// src/clip-event/model_clip.py:679-715; gradient per SURVEY.md 8a-8.
//
// Why streaming.  The IPOT solve of one sample is a chain of `iters` dependent iterations (~500 cycles
// each on one warp, ~28 k cycles per sample at 50 iterations) while the sample's share of the HBM time
// is ~5.6 k cycles: at least five or six solves have to be in flight per SM.  A sample's node rows
// (x [M,D] + y [N,D], 68 KB at 16x50x512) cannot stay resident for that many samples, but its cost tile
// (3 KB) can.  So the rows pass through shared memory TWICE in 16-row chunks -- once for the cost
// contraction, once (normally out of L2) for the gradient contraction -- and only the per-sample
// scratch (cost tile, plan factors, W) lives across the solve:
//
//   warp 0        cost loader   bulk copies (cp.async.bulk) of x and the y chunks into the cost rings
//   warp 1        grad loader   the same rows again into the gradient rings
//   warp 2        storer        bulk stores of dy chunks / dx, formed in place in the gradient rings
//   warps 3..4    cost          S = y x^t per 16-row chunk pair (mma.sync bf16), row norms on the FMA pipe
//   warps 5..8    gradient      dy = -W x + ay y (in place over the chunk), dx += -W^t y (registers,
//                               a quarter of the D columns per warp), dx + ax x over the x buffer
//   warps 9..     IPOT          one warp per park: the solver of csrc/ot_fused.cu (register-resident
//                               factorised plan, lane = image rows l and l+32)
//
// Hand-overs are mbarriers; every ring is used strictly in sample order, so a phase is a division.
#include <type_traits>

#include "ot_fused.cuh"

namespace ce {
namespace {

constexpr int kCH = 16;                // rows per chunk
constexpr int kMP = 16;                // text nodes padded to one m16 / two n8 tiles
constexpr int kNR = 64;                // image-node rows covered by the solver warp
constexpr int kSLd = 16;               // floats per row of the S tile: dense 64-byte rows, the 16-byte chunk index is
                                       // XOR-swizzled with (row >> 1) & 3 (conflict-free LDS.128 by row, 1 KB per park saved)
constexpr int kWLd = 24;               // bf16 per row of the W tile (48 B: conflict-free ldmatrix)
constexpr int kPLd = 18;               // floats per lane row of the column-sum transpose
constexpr int kMaxParks = 6;
constexpr int kMaxRing = 6;
constexpr int kCostWarps = 2, kGradWarps = 4;
// Warp roles by warpgroup (setmaxnreg moves registers between warpgroups: the kernel is compiled for
// 128 registers x 512 threads, the loaders give most of theirs to the gradient and solver warps):
//   warps 0..3    cost loader, gradient loader, storer, (idle)      56 registers
//   warps 4..7    gradient                                          152
//   warps 8..9    cost          } one warpgroup                     152
//   warps 10..15  IPOT solvers  }
constexpr int kGradWarp0 = 4, kCostWarp0 = 8, kFirstSolver = 10;
constexpr int kThreads = 512;

struct ParkScratch {                   // per park: everything that lives across the solve
  float S[kNR * kSLd];                 // raw dots (fp32); later W as bf16 [kNR][kWLd]
  float yn2[kNR];                      // |y|^2, later ay
  float xn2[kMP];                      // |x|^2, later ax
  float P[32 * kPLd];                  // IPOT column-sum transpose
  float w[kMP];                        // v * sigma broadcast
  float v[kMP];                        // v broadcast (refold / epilogue)
};

struct WBuf {                          // W, ay, ax of the sample the gradient warps work on
  __nv_bfloat16 W[kNR * kWLd];
  float ay[kNR];
  float ax[kMP];
};

struct Bars {
  uint64_t cx_full[2], cx_empty[2];
  uint64_t cy_full[kMaxRing], cy_empty[kMaxRing];
  uint64_t gx_full[2], gx_out[2], gx_empty[2];
  uint64_t gy_full[kMaxRing], gy_out[kMaxRing], gy_empty[kMaxRing];
  uint64_t s_ready[kMaxParks], w_ready[kMaxParks], scr_free[kMaxParks];
};

__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
               "r"(smem_u32(src)), "r"(bytes)
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
__device__ __forceinline__ void stsm_x4(uint32_t addr, const uint32_t* r) {
  asm volatile("stmatrix.sync.aligned.m8n8.x4.shared.b16 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3])
               : "memory");
}
__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, "
      "{%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
// A fragment of diag(d) restricted to rows g (value vg) and g+8 (value vg8): see the mma m16n8k16 layout
__device__ __forceinline__ void diag_frag(uint32_t* af, float vg, float vg8, int g, int t) {
  af[0] = pack2(2 * t == g ? vg : 0.f, 2 * t + 1 == g ? vg : 0.f);
  af[1] = 0u;
  af[2] = 0u;
  af[3] = pack2(2 * t == g ? vg8 : 0.f, 2 * t + 1 == g ? vg8 : 0.f);
}
__device__ __forceinline__ float frcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ bool is_pad(const void* mask, int kind, int64_t idx) {
  if (kind == CE_MASK_NUM_I64) return reinterpret_cast<const int64_t*>(mask)[idx] == 0;
  return reinterpret_cast<const uint8_t*>(mask)[idx] != 0;
}
// One lane polls (with a back-off), the rest of the warp joins at the __syncwarp.  A pipeline bug traps
// after 4 s instead of hanging the GPU.
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {   // non-blocking probe
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void role_wait(uint64_t* bar, uint32_t parity, int lane, uint32_t sleep_ns = 200) {
  // Every lane waits (the warp never diverges here: a lane-0-only poll loop left the solver warps running
  // split for the whole sample, 3x slower).  try_wait suspends the warp in hardware until the phase flips
  // or its time limit expires, so a waiting role costs next to no issue slots; test_wait + nanosleep
  // polling was measured at half of all instructions the kernel issued.
  (void)lane;
  uint32_t spins = 0;
  while (!(sleep_ns > 0 ? mbar_test_wait(bar, parity) : mbar_try_wait(bar, parity))) {
    if (sleep_ns > 0) __nanosleep(sleep_ns);
    if (++spins > (1u << 22)) __trap();   // seconds: a pipeline bug traps instead of hanging the GPU
  }
  __syncwarp();
}
// the four gradient warps among themselves (named barrier 1)
__device__ __forceinline__ void grad_bar() {
  __syncwarp();
  asm volatile("bar.sync 1, 128;" ::: "memory");
}
// all lanes' earlier shared-memory accesses are ordered before the arrival
__device__ __forceinline__ void role_arrive(uint64_t* bar, int lane) {
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}

#define OT_TRACE(k, e)                                                                          \
  do {                                                                                          \
    if (a.trace != nullptr && blockIdx.x == 0 && (k) < 64 && lane == 0) a.trace[(k) * 32 + (e)] = clock64(); \
  } while (0)

template <int REGS>
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS)); }
template <int REGS>
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS)); }

// PACKED: variable-length node sets (offset arrays instead of masks); a separate instantiation so that the padded
// kernel does not carry its code (the kernel is sensitive to its instruction footprint)
template <bool PACKED>
__global__ void __launch_bounds__(kThreads, 1) ot_stream_kernel(const OtFusedArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int RS = a.D * 2 + 16;                       // row stride of a chunk buffer in bytes
  const int row_bytes = a.D * 2;
  const int chunk_bytes = kCH * RS;
  const int P = a.slots, CY = a.cy_depth, GY = a.gy_depth;
  const int NC = (a.N + kCH - 1) / kCH;              // y chunks per sample (1..4)
  uint8_t* cx = smem;                                // [2] x buffers of the cost stage
  uint8_t* cy = cx + 2 * chunk_bytes;                // [CY] y chunk ring of the cost stage
  uint8_t* gx = cy + (size_t)CY * chunk_bytes;       // [2] x buffers of the gradient stage (dx in place)
  uint8_t* gy = gx + 2 * chunk_bytes;                // [GY] y chunk ring of the gradient stage (dy in place)
  uint8_t* zero_row = gy + (size_t)GY * chunk_bytes; // D*2 bytes of zeros: the whole-image slot's gradient
  WBuf* wbuf = reinterpret_cast<WBuf*>(zero_row + RS);
  ParkScratch* scr = reinterpret_cast<ParkScratch*>(wbuf + 1);
  Bars* bars = reinterpret_cast<Bars*>(scr + P);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int count = (a.B - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const bool grads = a.dtxt != nullptr;
  constexpr bool packed = PACKED;
  // per-sample geometry: element offsets of the sample's first text / image row and its node counts
  auto sample_geom = [&](int64_t b, int64_t& xo, int64_t& yo, int& m, int& n) {
    if (packed) {
      const int t0 = __ldg(a.txt_off + b), t1 = __ldg(a.txt_off + b + 1), i0 = __ldg(a.img_off + b), i1 = __ldg(a.img_off + b + 1);
      xo = (int64_t)t0 * a.D; yo = (int64_t)i0 * a.D; m = t1 - t0; n = i1 - i0;
    } else {
      xo = b * a.txt_bs; yo = b * a.img_bs; m = a.M; n = a.N;
    }
  };
  const uint32_t sleep_ns = (uint32_t)a.poll_mode;

  // ---- one-time: rows that no load ever covers (beyond M / N) must read as zeros -----------------
  {
    const int total16 = (int)(reinterpret_cast<uint8_t*>(bars) - smem) / 16;
    uint4* p = reinterpret_cast<uint4*>(smem);
    for (int i = tid; i < total16; i += blockDim.x) p[i] = make_uint4(0u, 0u, 0u, 0u);
    if (tid == 0) {
      for (int i = 0; i < 2; ++i) {
        mbar_init(&bars->cx_full[i], 1); mbar_init(&bars->cx_empty[i], kCostWarps);
        mbar_init(&bars->gx_full[i], 1); mbar_init(&bars->gx_out[i], kGradWarps); mbar_init(&bars->gx_empty[i], 1);
      }
      for (int i = 0; i < kMaxRing; ++i) {
        mbar_init(&bars->cy_full[i], 1); mbar_init(&bars->cy_empty[i], 1);
        mbar_init(&bars->gy_full[i], 1); mbar_init(&bars->gy_out[i], kGradWarps); mbar_init(&bars->gy_empty[i], 1);
      }
      for (int i = 0; i < kMaxParks; ++i) {
        mbar_init(&bars->s_ready[i], kCostWarps); mbar_init(&bars->w_ready[i], 1);
        mbar_init(&bars->scr_free[i], kGradWarps);
      }
      mbar_fence_init();
    }
    fence_proxy_async();     // generic zero-fill before the async-proxy loads into the same bytes
    __syncthreads();
  }

  if (warp < 4) {
  reg_dec<56>();
  if (warp == 3) return;
  if (warp == 0 || warp == 1) {
    // ===================================== loaders ===============================================
    const bool cost_side = warp == 0;
    if (!cost_side && !grads) return;
    uint8_t* xb = cost_side ? cx : gx;
    uint8_t* yb = cost_side ? cy : gy;
    uint64_t* x_full = cost_side ? bars->cx_full : bars->gx_full;
    uint64_t* x_empty = cost_side ? bars->cx_empty : bars->gx_empty;
    uint64_t* y_full = cost_side ? bars->cy_full : bars->gy_full;
    uint64_t* y_empty = cost_side ? bars->cy_empty : bars->gy_empty;
    const int R = cost_side ? CY : GY;
    int q = 0;                                       // y chunk sequence number: ring slot q % R, use q / R
    int qs = 0, qu = 0;                              // q % R and (q / R) & 1 without the divisions
    // the geometry of sample k+1 is fetched while sample k is being loaded (packed layout: two offsets per side;
    // their global-load latency would otherwise sit in this serial chain once per sample)
    int64_t xo_n = 0, yo_n = 0; int m_n = 0, n_n = 0;
    sample_geom((int64_t)blockIdx.x, xo_n, yo_n, m_n, n_n);
    for (int k = 0; k < count; ++k) {
      const int64_t xo = xo_n, yo = yo_n; const int m_b = m_n, n_b = n_n;
      if (k + 1 < count) sample_geom((int64_t)blockIdx.x + (int64_t)(k + 1) * gridDim.x, xo_n, yo_n, m_n, n_n);
      const uint8_t* xg = reinterpret_cast<const uint8_t*>(a.txt) + xo * 2;
      const uint8_t* yg = reinterpret_cast<const uint8_t*>(a.img) + yo * 2;
      const int xs = k & 1;
      role_wait(&x_empty[xs], ((k >> 1) & 1) ^ 1, lane, sleep_ns);
      if (cost_side) OT_TRACE(k, 0); else OT_TRACE(k, 1);
      if (lane == 0) mbar_expect_tx(&x_full[xs], (uint32_t)(m_b * row_bytes));
      __syncwarp();
      if (lane < m_b) bulk_load(xb + (size_t)xs * chunk_bytes + (size_t)lane * RS, xg + (int64_t)lane * row_bytes, (uint32_t)row_bytes, &x_full[xs]);
      for (int c = 0; c < NC; ++c, ++q) {
        const int rows = max(0, min(kCH, n_b - c * kCH));   // packed layout: a short sample's last chunks are empty
        role_wait(&y_empty[qs], qu ^ 1, lane, sleep_ns);
        if (lane == 0) mbar_expect_tx(&y_full[qs], (uint32_t)(rows * row_bytes));
        __syncwarp();
        if (lane < rows)
          bulk_load(yb + (size_t)qs * chunk_bytes + (size_t)lane * RS, yg + (int64_t)(c * kCH + lane) * row_bytes, (uint32_t)row_bytes, &y_full[qs]);
        if (++qs == R) { qs = 0; qu ^= 1; }
      }
    }
  } else if (warp == 2) {
    // ===================================== storer ================================================
    if (!grads) return;
    int qs = 0, qu = 0;
    int64_t xo_n = 0, yo_n = 0; int m_n = 0, n_n = 0;
    sample_geom((int64_t)blockIdx.x, xo_n, yo_n, m_n, n_n);
    for (int k = 0; k < count; ++k) {
      const int64_t xo = xo_n, yo = yo_n; const int m_b = m_n, n_b = n_n;
      if (k + 1 < count) sample_geom((int64_t)blockIdx.x + (int64_t)(k + 1) * gridDim.x, xo_n, yo_n, m_n, n_n);
      uint8_t* dxg = reinterpret_cast<uint8_t*>(a.dtxt) + xo * 2;
      uint8_t* dyg = reinterpret_cast<uint8_t*>(a.dimg) + yo * 2;
      for (int c = 0; c < NC; ++c) {
        const int rows = max(0, min(kCH, n_b - c * kCH));
        role_wait(&bars->gy_out[qs], qu, lane, sleep_ns);
        if (lane < rows) {
          bulk_store(dyg + (int64_t)(c * kCH + lane) * row_bytes, gy + (size_t)qs * chunk_bytes + (size_t)lane * RS, (uint32_t)row_bytes);
          tma_store_commit();
          tma_store_wait_read();
        }
        role_arrive(&bars->gy_empty[qs], lane);
        if (++qs == GY) { qs = 0; qu ^= 1; }
      }
      const int xs = k & 1;
      role_wait(&bars->gx_out[xs], (k >> 1) & 1, lane, sleep_ns);
      OT_TRACE(k, 8);
      if (lane < m_b) {
        bulk_store(dxg + (int64_t)lane * row_bytes, gx + (size_t)xs * chunk_bytes + (size_t)lane * RS, (uint32_t)row_bytes);
        tma_store_commit();
        tma_store_wait_read();
      } else if (lane == 31 && a.dslot0 != nullptr) {   // the dropped whole-image slot's gradient is zero
        bulk_store(reinterpret_cast<uint8_t*>(a.dslot0) + yo * 2, zero_row, (uint32_t)row_bytes);
        tma_store_commit();
        tma_store_wait_read();
      }
      role_arrive(&bars->gx_empty[xs], lane);
    }
    tma_store_wait_all();
  }
  } else if (warp >= kCostWarp0) {
  reg_inc<152>();
  if (warp < kFirstSolver) {
    // ===================================== cost warps ============================================
    // warp j contracts the chunk pair (2j, 2j+1) against x when the sample has three or four chunks,
    // chunk j alone otherwise (one x fragment load serves both chunks); both warps report to the park's
    // s_ready barrier.  Row norms come from the tensor core as the diagonals of chunk * chunk^t (an A
    // fragment re-read as B fragments), so the loop is ldmatrix + mma only; fragments are double-buffered
    // by hand (the asm statements keep their program order, so the loads of step ks+1 are written before
    // the MMAs of step ks).
    const int j = warp - kCostWarp0;
    const int lrow = lane & 15, lcol = (lane >> 4) * 8;
    const int ksteps = a.D / 16;                         // even (D is a multiple of 64)
    const bool paired = NC > 2;
    const int cA = paired ? 2 * j : j, cB = paired ? 2 * j + 1 : NC;   // cB == NC: absent
    const bool hasA = cA < NC, hasB = cB < NC;
    const bool diag = t == (g >> 1);                     // this lane holds G[g][g] and G[g+8][g+8]
    for (int k = 0; k < count; ++k) {
      const int xs = k & 1;
      const int park = k % P;
      float accA[2][4], accB[2][4], gA[2][4], gB[2][4], gX[2][4];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) { accA[i][q4] = 0.f; accB[i][q4] = 0.f; gA[i][q4] = 0.f; gB[i][q4] = 0.f; gX[i][q4] = 0.f; }
      const int qA = k * NC + cA, qB = k * NC + cB;
      const int sA = qA % CY, sB = qB % CY;
      // every cost warp follows the x buffer (a warp without a chunk must not run ahead of the loader)
      role_wait(&bars->cx_full[xs], (k >> 1) & 1, lane, sleep_ns);
      if (j == 0) OT_TRACE(k, 2);
      if (hasA) {
        role_wait(&bars->cy_full[sA], (qA / CY) & 1, lane, sleep_ns);
        if (hasB) role_wait(&bars->cy_full[sB], (qB / CY) & 1, lane, sleep_ns);
        if (j == 0) OT_TRACE(k, 3);
        const uint32_t xbase = smem_u32(cx + (size_t)xs * chunk_bytes);
        const int xrow = (lane & 7) + ((lane >> 4) << 3);
        const uint32_t xa = xbase + (uint32_t)(xrow * RS) + ((lane >> 3) & 1) * 16;
        const uint32_t yaA = smem_u32(cy + (size_t)sA * chunk_bytes) + (uint32_t)(lrow * RS) + lcol * 2;
        const uint32_t yaB = smem_u32(cy + (size_t)(hasB ? sB : sA) * chunk_bytes) + (uint32_t)(lrow * RS) + lcol * 2;
        uint32_t f0[12], f1[12];                          // x (B operand), chunk A, chunk B fragments of one k step
        auto ldf = [&](uint32_t* f, int ks) {
          ldsm_x4(f, xa + ks * 32);
          ldsm_x4(f + 4, yaA + ks * 32);
          ldsm_x4(f + 8, yaB + ks * 32);
        };
        auto mm = [&](const uint32_t* f, auto with_x) {
          const uint32_t* bf = f;
          const uint32_t* fa = f + 4;
          const uint32_t* fb = f + 8;
          mma16816(accA[0], fa, bf[0], bf[1]);
          mma16816(accA[1], fa, bf[2], bf[3]);
          mma16816(accB[0], fb, bf[0], bf[1]);
          mma16816(accB[1], fb, bf[2], bf[3]);
          mma16816(gA[0], fa, fa[0], fa[2]);               // rows 0..7 of chunk * chunk^t
          mma16816(gA[1], fa, fa[1], fa[3]);               // rows 8..15
          mma16816(gB[0], fb, fb[0], fb[2]);
          mma16816(gB[1], fb, fb[1], fb[3]);
          if constexpr (decltype(with_x)::value) {         // x x^t: the B fragments re-read as an A fragment
            const uint32_t xf[4] = {bf[0], bf[2], bf[1], bf[3]};
            mma16816(gX[0], xf, bf[0], bf[1]);
            mma16816(gX[1], xf, bf[2], bf[3]);
          }
        };
        auto run = [&](auto with_x) {
          ldf(f0, 0);
#pragma unroll 1
          for (int ks = 0; ks + 2 < ksteps; ks += 2) {
            ldf(f1, ks + 1);
            mm(f0, with_x);
            ldf(f0, ks + 2);
            mm(f1, with_x);
          }
          ldf(f1, ksteps - 1);
          mm(f0, with_x);
          mm(f1, with_x);
        };
        if (j == 0) run(std::true_type{});
        else run(std::false_type{});
        // the rows are in registers: the buffers go back to the loader
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&bars->cy_empty[sA]);
          if (hasB) mbar_arrive(&bars->cy_empty[sB]);
        }
      }
      role_arrive(&bars->cx_empty[xs], lane);
      // the park's scratch is free once the gradient warps have taken W, ax, ay of its previous sample
      if (j == 0) OT_TRACE(k, 4);
      role_wait(&bars->scr_free[park], ((k / P) & 1) ^ 1, lane, sleep_ns);
      if (j == 0) OT_TRACE(k, 5);
      ParkScratch& sc = scr[park];
      if (hasA) {
        float* Sp = sc.S + (cA * kCH + g) * kSLd + 2 * (t & 1);
        const int swg = (g >> 1) & 3;                    // the same key for rows g and g + 8
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int ch = ((2 * i + (t >> 1)) ^ swg) * 4;   // columns 8 i + 2 t, + 1 live in chunk 2 i + t / 2
          *reinterpret_cast<float2*>(Sp + ch) = make_float2(accA[i][0], accA[i][1]);
          *reinterpret_cast<float2*>(Sp + 8 * kSLd + ch) = make_float2(accA[i][2], accA[i][3]);
        }
        if (diag) { sc.yn2[cA * kCH + g] = (g & 1) ? gA[0][1] : gA[0][0]; sc.yn2[cA * kCH + g + 8] = (g & 1) ? gA[1][3] : gA[1][2]; }
      }
      if (hasB) {
        float* Sp = sc.S + (cB * kCH + g) * kSLd + 2 * (t & 1);
        const int swg = (g >> 1) & 3;                    // the same key for rows g and g + 8
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int ch = ((2 * i + (t >> 1)) ^ swg) * 4;   // columns 8 i + 2 t, + 1 live in chunk 2 i + t / 2
          *reinterpret_cast<float2*>(Sp + ch) = make_float2(accB[i][0], accB[i][1]);
          *reinterpret_cast<float2*>(Sp + 8 * kSLd + ch) = make_float2(accB[i][2], accB[i][3]);
        }
        if (diag) { sc.yn2[cB * kCH + g] = (g & 1) ? gB[0][1] : gB[0][0]; sc.yn2[cB * kCH + g + 8] = (g & 1) ? gB[1][3] : gB[1][2]; }
      }
      if (j == 0 && diag) { sc.xn2[g] = (g & 1) ? gX[0][1] : gX[0][0]; sc.xn2[g + 8] = (g & 1) ? gX[1][3] : gX[1][2]; }
      role_arrive(&bars->s_ready[park], lane);
    }
  } else {
    // ===================================== IPOT warps ============================================
    const int park = warp - kFirstSolver;
    if (park >= P) return;
    ParkScratch& sc = scr[park];
    const float nib2 = -1.4426950408889634f / a.beta;
    for (int k = park; k < count; k += P) {
      const int64_t b = (int64_t)blockIdx.x + (int64_t)k * gridDim.x;
      // masks (global loads issued before the wait so that their latency hides behind the cost phase)
      bool xp_l, yp0, yp1;
      if (packed) {   // every row of a packed sample is a valid node
        int64_t xo, yo; int m_b, n_b;
        sample_geom(b, xo, yo, m_b, n_b);
        xp_l = lane >= m_b; yp0 = lane >= n_b; yp1 = lane + 32 >= n_b;
      } else {
        xp_l = lane >= a.M || is_pad(a.txt_mask, a.mask_kind, b * a.txt_ms + lane);
        yp0 = lane >= a.N || is_pad(a.img_mask, a.mask_kind, b * a.img_ms + lane);
        yp1 = lane + 32 >= a.N || is_pad(a.img_mask, a.mask_kind, b * a.img_ms + lane + 32);
      }
      const uint32_t xpad = __ballot_sync(0xffffffffu, xp_l) | 0xffff0000u;   // bit m: text node m is padding
      const uint32_t yv0 = __ballot_sync(0xffffffffu, !yp0), yv1 = __ballot_sync(0xffffffffu, !yp1);
      const float xlen = (float)(kMP - __popc(xpad & 0xffffu));
      const float ylen = (float)(__popc(yv0) + __popc(yv1));
      const bool empty = xlen == 0.f || ylen == 0.f;     // model_ot.py:62: the whole plan is masked
      const float yg0 = yp0 ? 1e4f : 0.f, yg1 = yp1 ? 1e4f : 0.f;
      const int c = lane & 15;                            // the column this lane owns for sigma / v
      const float xg_c = ((xpad >> c) & 1u) ? 1e4f : 0.f;
      // column c of the 16 rows held by this half-warp's lanes; the upper half walks the rows 8 ahead so
      // that the two halves hit disjoint banks (18 * 8 = 16 mod 32)
      const float* pcol = sc.P + (lane >> 4) * 16 * kPLd + c;
      const int ssw = (lane >> 1) & 3;                    // S tile swizzle key of rows lane and lane + 32
      const int prot = (lane >> 4) * 8;

      role_wait(&bars->s_ready[park], ((k / P) & 1), lane, sleep_ns);
      OT_TRACE(k, 9);
      // ---- kernel matrix A = exp(-C/beta), R = 1 on valid pairs --------------------------------
      float2 A0[8], A1[8], R0[8], R1[8];
      const float rx_c = 1.f / fmaxf(sqrtf(sc.xn2[c]), a.eps);
      if (lane < kMP) sc.v[lane] = rx_c;
      __syncwarp();
      {
        const float n0 = sc.yn2[lane], n1 = sc.yn2[lane + 32];
        const float ry0 = 1.f / fmaxf(sqrtf(n0), a.eps), ry1 = 1.f / fmaxf(sqrtf(n1), a.eps);
        const float4* s0 = reinterpret_cast<const float4*>(sc.S + lane * kSLd);
        const float4* s1 = reinterpret_cast<const float4*>(sc.S + (lane + 32) * kSLd);
        const float4* xn = reinterpret_cast<const float4*>(sc.v);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 u0 = s0[q ^ ssw], u1 = s1[q ^ ssw], x4 = xn[q];
          const float sv0[4] = {u0.x, u0.y, u0.z, u0.w}, sv1[4] = {u1.x, u1.y, u1.z, u1.w};
          const float xx[4] = {x4.x, x4.y, x4.z, x4.w};
          float a0[4], a1[4];
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const int m = 4 * q + jj;
            const float rx = xx[jj];
            const bool xv = !((xpad >> m) & 1u) && !empty;
            a0[jj] = (xv && !yp0) ? ex2((1.f - sv0[jj] * rx * ry0) * nib2) : 0.f;
            a1[jj] = (xv && !yp1) ? ex2((1.f - sv1[jj] * rx * ry1) * nib2) : 0.f;
          }
          A0[2 * q] = make_float2(a0[0], a0[1]); A0[2 * q + 1] = make_float2(a0[2], a0[3]);
          A1[2 * q] = make_float2(a1[0], a1[1]); A1[2 * q + 1] = make_float2(a1[2], a1[3]);
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          R0[q] = make_float2(A0[q].x != 0.f ? 1.f : 0.f, A0[q].y != 0.f ? 1.f : 0.f);
          R1[q] = make_float2(A1[q].x != 0.f ? 1.f : 0.f, A1[q].y != 0.f ? 1.f : 0.f);
        }
      }
      __syncwarp();      // sc.v (inverse norms) has been read by every lane
      float u0 = 1.f, u1 = 1.f;
      float v_c = 1.f;
      float sig_c = (xg_c == 0.f && !empty) ? 1.f / xlen : 0.f;
      if (lane < kMP) sc.w[lane] = v_c * sig_c;
      __syncwarp();

      OT_TRACE(k, 10);
      // R holds A (.) plan-factor at the top of every iteration.  One delta/sigma round; with EARLY the
      // multiply by A for the NEXT iteration is issued while this round's column sums travel through
      // shared memory.  EARLY is a compile-time flag: behind a runtime branch (one copy of the round in the
      // instruction stream) the multiply no longer overlaps the loads and the solves ran 20 % slower.
      float z0 = 1.f, z1 = 1.f;
      auto round = [&](auto early_tag) {
        constexpr bool EARLY = decltype(early_tag)::value;
        float2 w2[8];
        {
          const float4* wp = reinterpret_cast<const float4*>(sc.w);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 w4 = wp[q];
            w2[2 * q] = make_float2(w4.x, w4.y); w2[2 * q + 1] = make_float2(w4.z, w4.w);
          }
        }
        // row sums: in-thread over the 16 columns
        float2 pa = __fmul2_rn(R0[0], w2[0]), pb = __fmul2_rn(R0[1], w2[1]);
        float2 qa = __fmul2_rn(R1[0], w2[0]), qb = __fmul2_rn(R1[1], w2[1]);
#pragma unroll
        for (int q = 2; q < 8; q += 2) {
          pa = __ffma2_rn(R0[q], w2[q], pa); pb = __ffma2_rn(R0[q + 1], w2[q + 1], pb);
          qa = __ffma2_rn(R1[q], w2[q], qa); qb = __ffma2_rn(R1[q + 1], w2[q + 1], qb);
        }
        const float rs0 = (pa.x + pa.y) + (pb.x + pb.y), rs1 = (qa.x + qa.y) + (qb.x + qb.y);
        const float d0 = frcp(ylen * (u0 * rs0) + yg0), d1 = frcp(ylen * (u1 * rs1) + yg1);
        z0 = d0 * u0; z1 = d1 * u1;
        // column sums: partials of this lane's two rows -> transpose through shared memory
        const float2 zz0 = make_float2(z0, z0), zz1 = make_float2(z1, z1);
        float2* prow = reinterpret_cast<float2*>(sc.P + lane * kPLd);
#pragma unroll
        for (int q = 0; q < 8; ++q) prow[q] = __ffma2_rn(zz1, R1[q], __fmul2_rn(zz0, R0[q]));
        __syncwarp();
        if constexpr (EARLY) {
#pragma unroll
          for (int q = 0; q < 8; ++q) { R0[q] = __fmul2_rn(R0[q], A0[q]); R1[q] = __fmul2_rn(R1[q], A1[q]); }
        }
        float e0 = 0.f, e1 = 0.f, e2 = 0.f, e3 = 0.f;
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          e0 += pcol[((i + 0 + prot) & 15) * kPLd]; e1 += pcol[((i + 1 + prot) & 15) * kPLd];
          e2 += pcol[((i + 2 + prot) & 15) * kPLd]; e3 += pcol[((i + 3 + prot) & 15) * kPLd];
        }
        float cs = (e0 + e1) + (e2 + e3);
        cs += __shfl_xor_sync(0xffffffffu, cs, 16);
        sig_c = frcp(xlen * (v_c * cs) + xg_c);
      };
      auto publish_w = [&]() {
        if (lane < kMP) sc.w[lane] = v_c * sig_c;
        __syncwarp();
      };
      // one reference iteration (model_ot.py:55-61): k inner rounds on the same Q, then T = delta Q sigma
      auto iteration = [&](auto early_tag) {
#pragma unroll 1
        for (int kk = 0; kk + 1 < a.k; ++kk) { round(std::false_type{}); publish_w(); }
        round(early_tag);
        u0 = z0; u1 = z1;
        v_c *= sig_c;
      };
#pragma unroll
      for (int q = 0; q < 8; ++q) { R0[q] = __fmul2_rn(R0[q], A0[q]); R1[q] = __fmul2_rn(R1[q], A1[q]); }
      // (u, v) are folded back into R every `refold` iterations: A^refold stays far above the fp32 underflow (C <= 2)
      const int refold = max(1, min(8, (int)(14.f * a.beta)));
#pragma unroll 1
      for (int it = 0; it < a.iters;) {
        const int n = min(refold, a.iters - it);
#pragma unroll 1
        for (int i = 0; i + 1 < n; ++i) { iteration(std::true_type{}); publish_w(); }
        iteration(std::false_type{});
        it += n;
        if (it < a.iters) {     // fold the scalings back into R, together with the next multiply by A
          if (lane < kMP) sc.v[lane] = v_c;
          __syncwarp();
          const float4* vp = reinterpret_cast<const float4*>(sc.v);
          const float2 uu0 = make_float2(u0, u0), uu1 = make_float2(u1, u1);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 v4 = vp[q];
            const float2 va = make_float2(v4.x, v4.y), vb = make_float2(v4.z, v4.w);
            R0[2 * q] = __fmul2_rn(__fmul2_rn(__fmul2_rn(R0[2 * q], uu0), va), A0[2 * q]);
            R0[2 * q + 1] = __fmul2_rn(__fmul2_rn(__fmul2_rn(R0[2 * q + 1], uu0), vb), A0[2 * q + 1]);
            R1[2 * q] = __fmul2_rn(__fmul2_rn(__fmul2_rn(R1[2 * q], uu1), va), A1[2 * q]);
            R1[2 * q + 1] = __fmul2_rn(__fmul2_rn(__fmul2_rn(R1[2 * q + 1], uu1), vb), A1[2 * q + 1]);
          }
          u0 = u1 = 1.f;
          v_c = 1.f;
          publish_w();
        }
      }

      OT_TRACE(k, 11);
      // ---- distance, W, normalisation-backward coefficients (T = u R v) ------------------------
      if (lane < kMP) { sc.v[lane] = v_c; sc.w[lane] = rx_c; }
      __syncwarp();
      float dsum = 0.f, py0 = 0.f, py1 = 0.f;
      const float n0 = sc.yn2[lane], n1 = sc.yn2[lane + 32];
      const float ry0 = 1.f / fmaxf(sqrtf(n0), a.eps), ry1 = 1.f / fmaxf(sqrtf(n1), a.eps);
      uint32_t wp0[8], wp1[8];
      float* prow = sc.P + lane * kPLd;   // scalar stores below: once per sample
      {
        const float4* s0 = reinterpret_cast<const float4*>(sc.S + lane * kSLd);
        const float4* s1 = reinterpret_cast<const float4*>(sc.S + (lane + 32) * kSLd);
        const float4* xn = reinterpret_cast<const float4*>(sc.w);
        const float4* vp = reinterpret_cast<const float4*>(sc.v);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 u40 = s0[q ^ ssw], u41 = s1[q ^ ssw], x4 = xn[q], v4 = vp[q];
          const float sv0[4] = {u40.x, u40.y, u40.z, u40.w}, sv1[4] = {u41.x, u41.y, u41.z, u41.w};
          const float xx[4] = {x4.x, x4.y, x4.z, x4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
          const float r0[4] = {R0[2 * q].x, R0[2 * q].y, R0[2 * q + 1].x, R0[2 * q + 1].y};
          const float r1[4] = {R1[2 * q].x, R1[2 * q].y, R1[2 * q + 1].x, R1[2 * q + 1].y};
          float w0[4], w1[4];
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const int m = 4 * q + jj;
            const float rx = xx[jj];
            const bool xv = !((xpad >> m) & 1u) && !empty;
            const float sh0 = sv0[jj] * rx * ry0, sh1 = sv1[jj] * rx * ry1;
            const float t0 = (xv && !yp0) ? u0 * r0[jj] * vv[jj] : 0.f;     // model_ot.py:62 final mask
            const float t1 = (xv && !yp1) ? u1 * r1[jj] * vv[jj] : 0.f;
            dsum += (1.f - sh0) * t0 + (1.f - sh1) * t1;
            const float tg0 = a.scale * t0, tg1 = a.scale * t1;
            py0 += tg0 * sh0; py1 += tg1 * sh1;
            prow[m] = tg0 * sh0 + tg1 * sh1;
            w0[jj] = -(tg0 * rx * ry0);                                  // -W: no sign flip in the MMAs
            w1[jj] = -(tg1 * rx * ry1);
          }
          wp0[2 * q] = pack2(w0[0], w0[1]); wp0[2 * q + 1] = pack2(w0[2], w0[3]);
          wp1[2 * q] = pack2(w1[0], w1[1]); wp1[2 * q + 1] = pack2(w1[2], w1[3]);
        }
      }
      __syncwarp();
      {   // column sums of tg * s^ -> ax
        float e = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) e += pcol[((i + prot) & 15) * kPLd];
        e += __shfl_xor_sync(0xffffffffu, e, 16);
        const float xn2c = sc.xn2[c];
        const float rxc = rx_c;
        // |x| < eps: F.normalize divides by eps and the projection term has no gradient
        const float axc = (sqrtf(xn2c) >= a.eps) ? e * rxc * rxc : 0.f;
        __syncwarp();    // every lane has read S, xn2 and P before they are overwritten
        if (lane < kMP) sc.xn2[lane] = axc;
      }
      sc.yn2[lane] = (sqrtf(n0) >= a.eps) ? py0 * ry0 * ry0 : 0.f;
      sc.yn2[lane + 32] = (sqrtf(n1) >= a.eps) ? py1 * ry1 * ry1 : 0.f;
      {   // W tile (bf16, [kNR][kWLd]) over the S tile
        __nv_bfloat16* Wb = reinterpret_cast<__nv_bfloat16*>(sc.S);
        uint4* d0 = reinterpret_cast<uint4*>(Wb + lane * kWLd);
        uint4* d1 = reinterpret_cast<uint4*>(Wb + (lane + 32) * kWLd);
        d0[0] = make_uint4(wp0[0], wp0[1], wp0[2], wp0[3]); d0[1] = make_uint4(wp0[4], wp0[5], wp0[6], wp0[7]);
        d1[0] = make_uint4(wp1[0], wp1[1], wp1[2], wp1[3]); d1[1] = make_uint4(wp1[4], wp1[5], wp1[6], wp1[7]);
      }
      dsum = warp_sum(dsum);
      if (lane == 0) a.dist[b] = dsum;
      OT_TRACE(k, 12);
      role_arrive(&bars->w_ready[park], lane);
    }
  }
  } else {
  reg_inc<152>();
  {
    // ===================================== gradient warps ========================================
    // Each warp owns D/4 columns.  Per sample: W, ax, ay move from the park's scratch to the hand-over
    // buffer (so the scratch goes back to the cost warps at once), then per 16-row chunk
    //   dx += (-W)^t y   (accumulators in registers across the chunks)
    //   dy  = (-W) x + ay y   in place over the chunk,
    // 16 columns at a time with the next pair's fragments loaded before the current pair's MMAs.
    const int gw = warp - kGradWarp0;                  // 0..3
    const int lrow = lane & 15, lcol = (lane >> 4) * 8;
    const int dcols = a.D / kGradWarps;                // a multiple of 16
    const int dc0 = gw * dcols;
    const int npairs = dcols / 16;                     // <= 8 (D <= 512)
    constexpr int kMaxPairs = 8;
    int qs = 0, qu = 0;
    for (int k = 0; k < count; ++k) {
      const int park = k % P;
      const int xs = k & 1;
      ParkScratch& sc = scr[park];
      role_wait(&bars->w_ready[park], (k / P) & 1, lane, sleep_ns);
      if (gw == 0) OT_TRACE(k, 6);
      if (!grads) {
        role_arrive(&bars->scr_free[park], lane);
        continue;
      }

      // W, ax, ay move to the hand-over buffer so that the scratch goes back to the cost warps at once.  (Keeping
      // the fragments of all four chunks in registers instead needs the chunk loop unrolled: four times the code,
      // measured 8 % slower -- the kernel is sensitive to its instruction footprint.)
      grad_bar();                                      // every gradient warp is done with the previous sample's W
      {
        const int gtid = gw * 32 + lane;
        const uint4* src = reinterpret_cast<const uint4*>(sc.S);
        uint4* dst = reinterpret_cast<uint4*>(wbuf->W);
        for (int i = gtid; i < kNR * kWLd * 2 / 16; i += kGradWarps * 32) dst[i] = src[i];
        if (gtid < kNR) wbuf->ay[gtid] = sc.yn2[gtid];
        else if (gtid < kNR + kMP) wbuf->ax[gtid - kNR] = sc.xn2[gtid - kNR];
      }
      grad_bar();
      role_arrive(&bars->scr_free[park], lane);        // the next sample's cost tile may take the scratch over
      const float ax0 = wbuf->ax[g], ax1 = wbuf->ax[g + 8];
      const uint32_t Wu = smem_u32(wbuf->W);
      if (gw == 0) OT_TRACE(k, 25);
      role_wait(&bars->gx_full[xs], (k >> 1) & 1, lane, sleep_ns);
      if (gw == 0) OT_TRACE(k, 26);
      const uint32_t xb = smem_u32(gx + (size_t)xs * chunk_bytes) + (uint32_t)(lrow * RS) + (uint32_t)((dc0 + lcol) * 2);
      float dxa[2 * kMaxPairs][4];
#pragma unroll
      for (int p = 0; p < 2 * kMaxPairs; ++p) { dxa[p][0] = 0.f; dxa[p][1] = 0.f; dxa[p][2] = 0.f; dxa[p][3] = 0.f; }
      for (int c = 0; c < NC; ++c) {
        uint32_t wtc[4], wac[4];                        // A = (-W)^t of the chunk's rows, A = -W rows of the chunk
        ldsm_x4_t(wtc, Wu + (uint32_t)(((c * 16 + (lane & 7) + ((lane >> 4) << 3)) * kWLd + ((lane >> 3) & 1) * 8) * 2));
        ldsm_x4(wac, Wu + (uint32_t)(((c * 16 + lrow) * kWLd + lcol) * 2));
        const float ay0 = wbuf->ay[c * 16 + g], ay1 = wbuf->ay[c * 16 + g + 8];
        if (gw == 0) OT_TRACE(k, 13 + c);
        role_wait(&bars->gy_full[qs], qu, lane, sleep_ns);
        // (an empty chunk of a short packed sample is NOT skipped: its stale rows meet exact zeros in W, and skipping
        // measured 12 % slower -- 237 -> 265 us at c3 with ragged sets -- the roles' relative timing matters more than
        // the saved MMAs)
        if (gw == 0) OT_TRACE(k, 17 + c);
        const uint32_t yb = smem_u32(gy + (size_t)qs * chunk_bytes) + (uint32_t)(lrow * RS) + (uint32_t)((dc0 + lcol) * 2);
        uint32_t f0[12], f1[12];                        // y as B operand (transposed), x as B operand, y in C layout
        auto ldf = [&](uint32_t* f, int p) {
          const uint32_t coff = (uint32_t)(p * 32);
          ldsm_x4_t(f, yb + coff);
          ldsm_x4_t(f + 4, xb + coff);
          ldsm_x4(f + 8, yb + coff);
        };
        auto pair = [&](const uint32_t* f, float* d0, float* d1, int p) {
          mma16816(d0, wtc, f[0], f[1]);                 // dx += (-W)^t y : K = the 16 image rows of this chunk
          mma16816(d1, wtc, f[2], f[3]);
          float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};
          mma16816(c0, wac, f[4], f[5]);                 // dy = (-W) x + ay y
          mma16816(c1, wac, f[6], f[7]);
          const uint32_t o[4] = {pack2(fmaf(ay0, bf_lo(f[8]), c0[0]), fmaf(ay0, bf_hi(f[8]), c0[1])),
                                 pack2(fmaf(ay1, bf_lo(f[9]), c0[2]), fmaf(ay1, bf_hi(f[9]), c0[3])),
                                 pack2(fmaf(ay0, bf_lo(f[10]), c1[0]), fmaf(ay0, bf_hi(f[10]), c1[1])),
                                 pack2(fmaf(ay1, bf_lo(f[11]), c1[2]), fmaf(ay1, bf_hi(f[11]), c1[3]))};
          stsm_x4(yb + (uint32_t)(p * 32), o);           // in place: the pair's own 16 columns
        };
        ldf(f0, 0);
#pragma unroll
        for (int p = 0; p < kMaxPairs; p += 2) {
          if (p < npairs) {
            if (p + 1 < npairs) ldf(f1, p + 1);
            pair(f0, dxa[2 * p], dxa[2 * p + 1], p);
            if (p + 2 < npairs) ldf(f0, p + 2);
            if (p + 1 < npairs) pair(f1, dxa[2 * p + 2], dxa[2 * p + 3], p + 1);
          }
        }
        fence_proxy_async();       // generic writes of dy -> visible to the bulk stores
        if (gw == 0) OT_TRACE(k, 21 + c);
        role_arrive(&bars->gy_out[qs], lane);
        if (++qs == GY) { qs = 0; qu ^= 1; }
      }
      // dx = acc + ax x, in place over x
#pragma unroll
      for (int p = 0; p < kMaxPairs; ++p) {
        if (p < npairs) {
          const uint32_t coff = (uint32_t)(p * 32);
          uint32_t xw[4];
          ldsm_x4(xw, xb + coff);
          const uint32_t o[4] = {pack2(fmaf(ax0, bf_lo(xw[0]), dxa[2 * p][0]), fmaf(ax0, bf_hi(xw[0]), dxa[2 * p][1])),
                                 pack2(fmaf(ax1, bf_lo(xw[1]), dxa[2 * p][2]), fmaf(ax1, bf_hi(xw[1]), dxa[2 * p][3])),
                                 pack2(fmaf(ax0, bf_lo(xw[2]), dxa[2 * p + 1][0]), fmaf(ax0, bf_hi(xw[2]), dxa[2 * p + 1][1])),
                                 pack2(fmaf(ax1, bf_lo(xw[3]), dxa[2 * p + 1][2]), fmaf(ax1, bf_hi(xw[3]), dxa[2 * p + 1][3]))};
          stsm_x4(xb + coff, o);
        }
      }
      fence_proxy_async();
      if (gw == 0) OT_TRACE(k, 7);
      role_arrive(&bars->gx_out[xs], lane);
    }
  }
  }
}

}  // namespace

size_t ot_stream_smem_bytes(int D, int parks, int cy_depth, int gy_depth) {
  const size_t RS = (size_t)D * 2 + 16;
  return (size_t)(4 + cy_depth + gy_depth) * kCH * RS + RS + sizeof(WBuf) + (size_t)parks * sizeof(ParkScratch) + sizeof(Bars) + 128;
}

// ring depths and parks for a shape: the gradient ring gets 3 chunks, the cost ring a whole sample (up to 4),
// the rest of the shared memory goes to parks (concurrent IPOT solves)
bool ot_stream_plan(int M, int N, int D, int dtype, int* parks, int* cy_depth, int* gy_depth) {
  if (dtype != CE_BF16 || M < 1 || M > kMP || N < 1 || N > kNR) return false;
  if (D % 64 != 0 || D < 64 || D > 512) return false;
  const int NC = (N + kCH - 1) / kCH;
  // the cost ring holds a whole sample (a cost warp works on a chunk pair; with one-bit phases it must never
  // wait two uses ahead of a slot: cy >= NC - 1), the gradient ring three chunks
  int cy = NC < 3 ? 3 : NC, gy = 3;
  static const int cy_env = [] { const char* e = getenv("CE_OT_CY"); return e ? atoi(e) : 0; }();
  static const int gy_env = [] { const char* e = getenv("CE_OT_GY"); return e ? atoi(e) : 0; }();
  static const int p_env = [] { const char* e = getenv("CE_OT_PARKS"); return e ? atoi(e) : 0; }();
  if (cy_env >= 3 && cy_env >= NC - 1 && cy_env <= kMaxRing) cy = cy_env;
  if (gy_env >= 2 && gy_env <= kMaxRing) gy = gy_env;
  int p = kMaxParks;
  while (p > 0 && ot_stream_smem_bytes(D, p, cy, gy) > 232448) --p;
  if (p_env >= 1 && p_env < p) p = p_env;
  if (p < 2) return false;
  *parks = p; *cy_depth = cy; *gy_depth = gy;
  return true;
}

bool ot_stream_supported(int M, int N, int D, int dtype) {
  int p, cy, gy;
  return ot_stream_plan(M, N, D, dtype, &p, &cy, &gy);
}

int launch_ot_stream(OtFusedArgs a, cudaStream_t st) {
  if (!ot_stream_plan(a.M, a.N, a.D, CE_BF16, &a.slots, &a.cy_depth, &a.gy_depth))
    return fail(CE_ERR_SHAPE, "OT stream: unsupported shape M=%d N=%d D=%d", a.M, a.N, a.D);
  const size_t smem = ot_stream_smem_bytes(a.D, a.slots, a.cy_depth, a.gy_depth);
  const int grid = a.B < num_sms() ? a.B : num_sms();
  {
    const char* e = getenv("CE_OT_TRACE_PTR");   // debug: tools/ot_trace.py
    a.trace = e != nullptr ? reinterpret_cast<long long*>(strtoull(e, nullptr, 0)) : nullptr;
  }
  {
    static const int sl = [] { const char* e = getenv("CE_OT_SLEEP"); return e ? atoi(e) : 0; }();
    a.poll_mode = sl;
  }
  if (a.txt_off != nullptr) {
    CE_CUDA_TRY(cudaFuncSetAttribute(ot_stream_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ot_stream_kernel<true><<<grid, kThreads, smem, st>>>(a);
  } else {
    CE_CUDA_TRY(cudaFuncSetAttribute(ot_stream_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ot_stream_kernel<false><<<grid, kThreads, smem, st>>>(a);
  }
  CE_LAUNCH_CHECK();
  return CE_OK;
}

}  // namespace ce
