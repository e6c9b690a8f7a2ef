// OT graph-alignment loss as ONE persistent, shared-memory-resident kernel (bf16, small plans).
// Reference behaviour: src/clip-event/model_ot.py:8-84 (cost_matrix_cosine -> ipot -> trace) and
// src/clip-event/model_clip.py:679-715; gradient per SURVEY.md 8a-8.
//
// One CTA per SM, warp-specialised, `S` samples resident in shared memory at a time:
//   warp 0        IO     per-row bulk TMA copies (cp.async.bulk) of a sample's x [M,D] and y [N,D]
//                        into a padded slot (row stride D*2+16 B: conflict-free ldmatrix), and bulk
//                        stores of the gradients that were formed IN PLACE over x and y
//   warps 1..3    IPOT   one warp per slot: masks, C = 1 - S/(|x||y|), the IPOT iterations on a
//                        register-resident factorised plan (lane = image-node rows l and l+32, all
//                        text-node columns in registers; row sums are in-thread, column sums go
//                        through a padded shared-memory transpose), distance, W, ax, ay
//   warps 4..11   MMA    cost contraction S = y x^t (mma.sync bf16, K split over two warp sets, row
//                        norms from the tensor core as the diagonal of tile*tile^t), then
//                        dx = -W^t y + ax x (kept packed in registers), dy = -W x + ay y written over
//                        y, dx written over x
// x and y cross HBM once in each direction (algorithmic bytes 2 (M+N) D e per sample); the solver's
// latency (50 dependent iterations, ~8 us) is hidden by the other slots' loads, contractions and
// stores.  Roles hand over through mbarriers: full (TMA bytes) -> s_ready -> w_ready -> out_ready.
#include <type_traits>

#include "ot_fused.cuh"

namespace ce {
namespace {

constexpr int kThreads = 384;          // 12 warps
constexpr int kMmaWarps = 8;
constexpr int kMaxSlots = 3;
constexpr int kMP = 16;                // text nodes padded to one m16 / two n8 tiles
constexpr int kNR = 64;                // image-node rows covered by the four m16 tiles
constexpr int kSLd = 20;               // floats per row of the S hand-over tile (conflict-free LDS.128)
constexpr int kWLd = 24;               // bf16 per row of the W tile (48 B: conflict-free ldmatrix)
constexpr int kPLd = 18;               // floats per lane row of the column-sum transpose (conflict-free STS.64 / LDS.32)

struct SlotScratch {                   // per slot, lives after the slots in shared memory
  float S[kNR * kSLd];                 // raw dots (fp32); later W as bf16 [kNR][kWLd]
  float yn2[kNR];                      // |y|^2, later ay
  float xn2[kMP];                      // |x|^2, later ax
  float P[32 * kPLd];                  // IPOT column-sum transpose
  float w[kMP];                        // v * sigma broadcast
  float v[kMP];                        // v broadcast (refold / epilogue)
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
__device__ __forceinline__ float bf16r(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
// A fragment of diag(d) restricted to rows g (value vg) and g+8 (value vg8): see mma m16n8k16 layout
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
// One lane polls (with a back-off: nine warps x 32 lanes spinning on try_wait flood the shared-memory
// pipe the solver warps live on), the rest of the warp joins at the __syncwarp.
__device__ __forceinline__ void warp_wait(uint64_t* bar, uint32_t parity, int lane, int mode = 0) {
  if (mode == 2) { mbar_wait(bar, parity); return; }
  if (mode == 1) {
    if (lane == 0) { while (!mbar_try_wait(bar, parity)) __nanosleep(20); }
    __syncwarp();
    return;
  }
  if (lane == 0) {
    if (!mbar_try_wait(bar, parity)) {
      uint32_t spins = 0;
      uint64_t t0 = 0;
      while (!mbar_try_wait(bar, parity)) {
        __nanosleep(64);
        if ((++spins & 0x3fffu) == 0) {
          uint64_t now;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
          if (t0 == 0) t0 = now;
          else if (now - t0 > 4000000000ull) {   // 4 s: a pipeline bug traps instead of hanging the GPU
            printf("clip_event_b200: OT mbarrier wait timed out (block %d warp %d)\n", (int)blockIdx.x, (int)(threadIdx.x >> 5));
            __trap();
          }
        }
      }
    }
  }
  __syncwarp();
}
__device__ __forceinline__ void mma_bar() {
  __syncwarp();
  asm volatile("bar.sync 1, 256;" ::: "memory");
}

// Role timeline (tuning aid): with CE_OT_TRACE_PTR set to a device buffer of 64 x 32 int64, CTA 0 records
// clock64 at the hand-over points of the first 64 samples it processes: trace[k * 32 + event].
#define OT_TRACE(k, e)                                                                          \
  do {                                                                                          \
    if (a.trace != nullptr && blockIdx.x == 0 && (k) < 64 && lane == 0) a.trace[(k) * 32 + (e)] = clock64(); \
  } while (0)


// tcgen05.st / ld of 32 lanes x 32 columns (one 32-bit word per lane and column): the parking copies
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Producer/consumer hand-overs between the roles go through NAMED BARRIERS: the consumer blocks in hardware
// (bar.sync) until the producer's bar.arrive -- no polling.  Measured: lanes polling an mbarrier with try_wait
// slow the latency-bound solver warps of the same SM by up to 3x (the faster the poll, the worse).
//   id 1            the eight MMA warps among themselves
//   id 2 + park     cost tile ready       (MMA warps arrive, solver warp syncs)
//   id 5 + park     W, ax, ay ready       (solver warp arrives, MMA warps sync)
//   id 8            gradients in place    (MMA warps arrive, IO warp syncs)
//   id 9            gradient slot free    (IO warp arrives after the bulk stores have read it, MMA warps sync)
//   id 11 + cslot   cost slot parked      (MMA warps arrive, IO warp syncs and refills the slot)
// bar.* is an ALIGNED instruction: the warp must be converged, or each divergent group counts as an arrival.
constexpr int kHandoverThreads = 32 + 256;
__device__ __forceinline__ void named_arrive(int id) {
  __threadfence_block();
  __syncwarp();
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(kHandoverThreads) : "memory");
}
__device__ __forceinline__ void named_sync(int id) {
  __syncwarp();
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(kHandoverThreads) : "memory");
}

// Pipeline.  A CTA's samples k = 0 .. count-1 pass through
//   C(k)  load into COST slot k % 2, cost contraction S = y x^t (+ row norms), then the slot's bytes are PARKED in
//         tensor memory (park k % P) and the slot is refilled with sample k + 2
//   I(k)  solver warp k % P: IPOT on the cost tile in scratch[k % P]  (~35 k cycles: the long pole)
//   G(k)  sample k is un-parked into the GRADIENT slot, dx, dy are formed in place and bulk-stored
// The MMA warps run  C(0) .. C(P-1) | [C(P) early] G(0) park(P) | [C(P+1) early] G(1) park(P+1) | ...  where
// "early" means: the cost tile of the park's NEXT sample is computed BEFORE the solver finishes the current one
// and waits in registers, so that the solver restarts the moment G(k) has taken W, ax, ay out of the scratch.
__global__ void __launch_bounds__(kThreads, 1) ot_fused_kernel(const OtFusedArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int RS = a.D * 2 + 16;                       // slot row stride in bytes
  const int slot_bytes = (a.M + a.N) * RS;
  const int P = a.slots;                             // parks = solver warps in use (1..3)
  constexpr int kGSlot = 2;                          // slots 0, 1: cost jobs (alternating); slot 2: gradient jobs
  uint8_t* zero_row = smem + (size_t)kMaxSlots * slot_bytes;  // D*2 bytes of zeros (clamped rows read it)
  uint8_t* trash_row = zero_row + RS;                 // clamped rows write here
  SlotScratch* scr = reinterpret_cast<SlotScratch*>(trash_row + RS);
  uint64_t* bars = reinterpret_cast<uint64_t*>(scr + kMaxSlots);
  uint64_t* full = bars;                      // [2] per cost slot: TMA bytes landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kMaxSlots);
  volatile int* early_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  // samples of this CTA: b = blockIdx.x + k * gridDim.x
  const int count = (a.B - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  // ---- one-time: zero everything the tensor core may read before a load covered it -------------
  {
    const int total16 = (int)(reinterpret_cast<uint8_t*>(bars) - smem) / 16;
    uint4* p = reinterpret_cast<uint4*>(smem);
    for (int i = tid; i < total16; i += kThreads) p[i] = make_uint4(0u, 0u, 0u, 0u);
    if (tid == 0) {
      for (int s = 0; s < kMaxSlots; ++s) mbar_init(&full[s], 1);
      mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, 512);   // the whole tensor memory: three parked samples
    fence_proxy_async();     // generic zero-fill before the async-proxy loads into the same bytes
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  const uint32_t tmem_base = *tmem_slot;

  const int row_bytes = a.D * 2;
  if (warp == 0) {
    // ===================================== IO warp ==============================================
    auto issue_load = [&](int k) {
      const int slot = k & 1;
      const int64_t b = (int64_t)blockIdx.x + (int64_t)k * gridDim.x;
      uint8_t* sl = smem + (size_t)slot * slot_bytes;
      const uint8_t* xg = reinterpret_cast<const uint8_t*>(a.txt) + b * a.txt_bs * 2;
      const uint8_t* yg = reinterpret_cast<const uint8_t*>(a.img) + b * a.img_bs * 2;
      OT_TRACE(k, 0);
      if (lane == 0) mbar_expect_tx(&full[slot], (uint32_t)((a.M + a.N) * row_bytes));
      __syncwarp();
      if (a.dbg & 2) {   // debug: no loads (the consumer sees the arrival only)
        if (lane == 0) asm volatile("mbarrier.complete_tx.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&full[slot])), "r"((uint32_t)((a.M + a.N) * row_bytes)) : "memory");
        return;
      }
      for (int r = lane; r < a.M + a.N; r += 32) {
        const uint8_t* src = r < a.M ? xg + (int64_t)r * row_bytes : yg + (int64_t)(r - a.M) * row_bytes;
        bulk_load(sl + (size_t)r * RS, src, (uint32_t)row_bytes, &full[slot]);
      }
      __syncwarp();
    };
    for (int k = 0; k < 2 && k < count; ++k) issue_load(k);
    named_arrive(9);                               // the gradient slot starts out free
    // samples 2 .. P+1 follow the parks of the prologue
    for (int k = 2; k < P + 2 && k < count; ++k) {
      named_sync(11 + (k & 1));
      issue_load(k);
    }
    for (int k = 0; k < count; ++k) {
      named_sync(8);                               // gradients of sample k are in the gradient slot
      OT_TRACE(k, 1);
      if (a.dtxt != nullptr && !(a.dbg & 2)) {
        const int64_t b = (int64_t)blockIdx.x + (int64_t)k * gridDim.x;
        uint8_t* sl = smem + (size_t)kGSlot * slot_bytes;
        uint8_t* dxg = reinterpret_cast<uint8_t*>(a.dtxt) + b * a.txt_bs * 2;
        uint8_t* dyg = reinterpret_cast<uint8_t*>(a.dimg) + b * a.img_bs * 2;
        for (int r = lane; r < a.M + a.N; r += 32) {
          uint8_t* dst = r < a.M ? dxg + (int64_t)r * row_bytes : dyg + (int64_t)(r - a.M) * row_bytes;
          bulk_store(dst, sl + (size_t)r * RS, (uint32_t)row_bytes);
        }
        if (a.dslot0 != nullptr && lane == 31)   // the dropped whole-image slot's gradient is zero
          bulk_store(reinterpret_cast<uint8_t*>(a.dslot0) + b * a.img_bs * 2, zero_row, (uint32_t)row_bytes);
        tma_store_commit();
      }
      __syncwarp();
      tma_store_wait_read();                       // the gradient slot has been read out: hand it back first
      OT_TRACE(k, 2);
      if (k + 1 < count) named_arrive(9);
      // sample k+P was parked at the end of this gradient job: its cost slot takes sample k+P+2
      if (k + P + 2 < count) {
        named_sync(11 + ((k + P) & 1));
        issue_load(k + P + 2);
      }
    }
    tma_store_wait_all();
  } else if (warp < 4) {
    // ===================================== IPOT warps ===========================================
    const int park = warp - 1;
    if (park < P) {
      SlotScratch& sc = scr[park];
      const float nib2 = -1.4426950408889634f / a.beta;
      for (int k = park; k < count; k += P) {
        const int64_t b = (int64_t)blockIdx.x + (int64_t)k * gridDim.x;
        // masks (global loads issued before the wait so that their latency hides behind the cost phase)
        const bool xp_l = lane >= a.M || is_pad(a.txt_mask, a.mask_kind, b * a.txt_ms + lane);
        const bool yp0 = lane >= a.N || is_pad(a.img_mask, a.mask_kind, b * a.img_ms + lane);
        const bool yp1 = lane + 32 >= a.N || is_pad(a.img_mask, a.mask_kind, b * a.img_ms + lane + 32);
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
        const int prot = (lane >> 4) * 8;

        named_sync(2 + park);
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
            const float4 u0 = s0[q], u1 = s1[q], x4 = xn[q];
            const float sv0[4] = {u0.x, u0.y, u0.z, u0.w}, sv1[4] = {u1.x, u1.y, u1.z, u1.w};
            const float xx[4] = {x4.x, x4.y, x4.z, x4.w};
            float a0[4], a1[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int m = 4 * q + j;
              const float rx = xx[j];
              const bool xv = !((xpad >> m) & 1u) && !empty;
              a0[j] = (xv && !yp0) ? ex2((1.f - sv0[j] * rx * ry0) * nib2) : 0.f;
              a1[j] = (xv && !yp1) ? ex2((1.f - sv1[j] * rx * ry1) * nib2) : 0.f;
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
        // shared memory (branch-free, so the compiler can interleave it with the loads).
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
          for (int kk = 0; kk + 1 < a.k; ++kk) { round(std::false_type{}); publish_w(); }
          round(early_tag);
          u0 = z0; u1 = z1;
          v_c *= sig_c;
        };
#pragma unroll
        for (int q = 0; q < 8; ++q) { R0[q] = __fmul2_rn(R0[q], A0[q]); R1[q] = __fmul2_rn(R1[q], A1[q]); }
        // (u, v) are folded back into R every `refold` iterations: A^refold stays far above the fp32 underflow (C <= 2)
        const int refold = max(1, min(8, (int)(14.f * a.beta)));
        for (int it = 0; it < a.iters;) {
          const int n = min(refold, a.iters - it);
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
            const float4 u40 = s0[q], u41 = s1[q], x4 = xn[q], v4 = vp[q];
            const float sv0[4] = {u40.x, u40.y, u40.z, u40.w}, sv1[4] = {u41.x, u41.y, u41.z, u41.w};
            const float xx[4] = {x4.x, x4.y, x4.z, x4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
            const float r0[4] = {R0[2 * q].x, R0[2 * q].y, R0[2 * q + 1].x, R0[2 * q + 1].y};
            const float r1[4] = {R1[2 * q].x, R1[2 * q].y, R1[2 * q + 1].x, R1[2 * q + 1].y};
            float w0[4], w1[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int m = 4 * q + j;
              const float rx = xx[j];
              const bool xv = !((xpad >> m) & 1u) && !empty;
              const float sh0 = sv0[j] * rx * ry0, sh1 = sv1[j] * rx * ry1;
              const float t0 = (xv && !yp0) ? u0 * r0[j] * vv[j] : 0.f;     // model_ot.py:62 final mask
              const float t1 = (xv && !yp1) ? u1 * r1[j] * vv[j] : 0.f;
              dsum += (1.f - sh0) * t0 + (1.f - sh1) * t1;
              const float tg0 = a.scale * t0, tg1 = a.scale * t1;
              py0 += tg0 * sh0; py1 += tg1 * sh1;
              prow[m] = tg0 * sh0 + tg1 * sh1;
              w0[j] = -(tg0 * rx * ry0);                                  // -W: no sign flip in the MMAs
              w1[j] = -(tg1 * rx * ry1);
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
        named_arrive(5 + park);
      }
    }
  } else {
    // ===================================== MMA warps ============================================
    const int mw = warp - 4;
    const int nt = mw & 3, hf = mw >> 2;          // also: TMEM lane quadrant (warp % 4) and column half
    const int ksteps = a.D / 16;
    const int kh0 = hf * (ksteps / 2), kh1 = hf == 0 ? ksteps / 2 : ksteps;
    const int lrow = lane & 15, lcol = (lane >> 4) * 8;       // ldmatrix / stmatrix lane -> (row, column) of a 16x16 block
    const uint32_t zero_u = smem_u32(zero_row), trash_u = smem_u32(trash_row);
    // parking geometry: the slot's (M+N) rows x D*2 bytes are cut into units of 32 16-byte chunks; unit u
    // belongs to TMEM lane quadrant u % 4, group u / 4; eight groups fill one 32x32b.x32 store (32 columns)
    const int upr_shift = a.D == 512 ? 1 : 0;                 // units per row: D / 256 (1 or 2)
    const int units = (a.M + a.N) << upr_shift;
    const int groups = (units + 3) >> 2;
    const int batches = (groups + 7) >> 3;
    const int park_cols = batches * 32;
    const uint32_t tq = tmem_base + ((uint32_t)(nt * 32) << 16);

    auto park_io = [&](int slot, int park, bool to_tmem) {
      uint8_t* sl = smem + (size_t)slot * slot_bytes;
      for (int bt = hf; bt < batches; bt += 2) {
        uint32_t r[32];
        const uint32_t taddr = tq + (uint32_t)(park * park_cols + bt * 32);
        if (!to_tmem) {
          tmem_ld32(taddr, reinterpret_cast<float*>(r));
          tmem_ld_wait();
        }
#pragma unroll
        for (int g8 = 0; g8 < 8; ++g8) {
          const int u = ((bt * 8 + g8) << 2) + nt;
          const bool in = u < units;
          const int row = u >> upr_shift, jj = u & ((1 << upr_shift) - 1);
          uint4* p = reinterpret_cast<uint4*>(sl + (size_t)row * RS + (size_t)(jj * 32 + lane) * 16);
          if (to_tmem) {
            const uint4 v = in ? *p : make_uint4(0u, 0u, 0u, 0u);
            r[4 * g8] = v.x; r[4 * g8 + 1] = v.y; r[4 * g8 + 2] = v.z; r[4 * g8 + 3] = v.w;
          } else if (in) {
            *p = make_uint4(r[4 * g8], r[4 * g8 + 1], r[4 * g8 + 2], r[4 * g8 + 3]);
          }
        }
        if (to_tmem) tmem_st32(taddr, r);
      }
      if (to_tmem) tmem_st_wait();
    };

    uint32_t ph_full = 0;
    // cost tile of the NEXT sample of a park: computed while the solver still owns the scratch, kept in
    // registers until the gradient job of the current sample has taken W, ax, ay out of it
    float cacc[2][4], cyn_lo = 0.f, cyn_hi = 0.f, cxn_lo = 0.f, cxn_hi = 0.f;
    auto cost_compute = [&](int k) {
      const int slot = k & 1;
      const uint32_t sl = smem_u32(smem + (size_t)slot * slot_bytes);
      // A operand: image-node rows 16 nt .. +15; B operand: the 16 text-node rows (as two n8 tiles)
      const int yrow = nt * 16 + lrow;
      const uint32_t ya = (yrow < a.N ? sl + (uint32_t)((a.M + yrow) * RS) : zero_u) + lcol * 2;
      const int xrow = (lane & 7) + ((lane >> 4) << 3);
      const uint32_t xa = (xrow < a.M ? sl + (uint32_t)(xrow * RS) : zero_u) + ((lane >> 3) & 1) * 16;
      float acc[2][4];
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[j][q] = 0.f;
      // sums of squares on the FMA pipe (the legacy mma pipe is the scarcer one here): rows g / g+8 of this
      // warp's image tile, and -- warps with nt == 0 -- text rows g / g+8
      float2 ys0 = make_float2(0.f, 0.f), ys1 = ys0, xs0 = ys0, xs1 = ys0;
      warp_wait(&full[slot], (ph_full >> slot) & 1u, lane, a.poll_mode);
      ph_full ^= 1u << slot;
      if (mw == 0) OT_TRACE(k, 3);
#pragma unroll 4
      for (int ks = kh0; ks < ((a.dbg & 1) ? kh0 : kh1); ++ks) {
        uint32_t af[4], bf[4];
        ldsm_x4(af, ya + ks * 32);
        ldsm_x4(bf, xa + ks * 32);
        mma16816(acc[0], af, bf[0], bf[1]);
        mma16816(acc[1], af, bf[2], bf[3]);
        {
          const float2 p0 = make_float2(bf_lo(af[0]), bf_hi(af[0])), p2 = make_float2(bf_lo(af[2]), bf_hi(af[2]));
          const float2 p1 = make_float2(bf_lo(af[1]), bf_hi(af[1])), p3 = make_float2(bf_lo(af[3]), bf_hi(af[3]));
          ys0 = __ffma2_rn(p0, p0, ys0); ys0 = __ffma2_rn(p2, p2, ys0);
          ys1 = __ffma2_rn(p1, p1, ys1); ys1 = __ffma2_rn(p3, p3, ys1);
        }
        if (nt == 0) {
          const float2 q0 = make_float2(bf_lo(bf[0]), bf_hi(bf[0])), q1 = make_float2(bf_lo(bf[1]), bf_hi(bf[1]));
          const float2 q2 = make_float2(bf_lo(bf[2]), bf_hi(bf[2])), q3 = make_float2(bf_lo(bf[3]), bf_hi(bf[3]));
          xs0 = __ffma2_rn(q0, q0, xs0); xs0 = __ffma2_rn(q1, q1, xs0);
          xs1 = __ffma2_rn(q2, q2, xs1); xs1 = __ffma2_rn(q3, q3, xs1);
        }
      }
      float yn_lo = ys0.x + ys0.y, yn_hi = ys1.x + ys1.y, xn_lo = xs0.x + xs0.y, xn_hi = xs1.x + xs1.y;
#pragma unroll
      for (int o = 1; o <= 2; o <<= 1) {
        yn_lo += __shfl_xor_sync(0xffffffffu, yn_lo, o); yn_hi += __shfl_xor_sync(0xffffffffu, yn_hi, o);
        xn_lo += __shfl_xor_sync(0xffffffffu, xn_lo, o); xn_hi += __shfl_xor_sync(0xffffffffu, xn_hi, o);
      }
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int q = 0; q < 4; ++q) cacc[j][q] = acc[j][q];
      cyn_lo = yn_lo; cyn_hi = yn_hi; cxn_lo = xn_lo; cxn_hi = xn_hi;
      if (mw == 0) OT_TRACE(k, 4);
    };
    // hand the tile over to the solver warp: the first K half stores, the second adds
    auto cost_writeback = [&](int k) {
      const int park = k % P;
      SlotScratch& sc = scr[park];
      float* Sp = sc.S + (nt * 16 + g) * kSLd + 2 * t;
      if (hf == 0) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          *reinterpret_cast<float2*>(Sp + 8 * j) = make_float2(cacc[j][0], cacc[j][1]);
          *reinterpret_cast<float2*>(Sp + 8 * kSLd + 8 * j) = make_float2(cacc[j][2], cacc[j][3]);
        }
        if (t == 0) {
          sc.yn2[nt * 16 + g] = cyn_lo; sc.yn2[nt * 16 + g + 8] = cyn_hi;
          if (nt == 0) { sc.xn2[g] = cxn_lo; sc.xn2[g + 8] = cxn_hi; }
        }
      }
      mma_bar();
      if (hf == 1) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          atomicAdd(Sp + 8 * j, cacc[j][0]); atomicAdd(Sp + 8 * j + 1, cacc[j][1]);
          atomicAdd(Sp + 8 * kSLd + 8 * j, cacc[j][2]); atomicAdd(Sp + 8 * kSLd + 8 * j + 1, cacc[j][3]);
        }
        if (t == 0) {
          atomicAdd(&sc.yn2[nt * 16 + g], cyn_lo); atomicAdd(&sc.yn2[nt * 16 + g + 8], cyn_hi);
          if (nt == 0) { atomicAdd(&sc.xn2[g], cxn_lo); atomicAdd(&sc.xn2[g + 8], cxn_hi); }
        }
      }
      named_arrive(2 + park);
    };
    // the sample's bytes move to tensor memory for the duration of the solve; its cost slot is free for sample k + 2
    auto cost_park = [&](int k) {
      park_io(k & 1, k % P, true);
      if (k + 2 < count) named_arrive(11 + (k & 1));
    };

    auto grad_job = [&](int k, int k_next) {      // k_next: sample whose cost tile waits in registers, or -1
      const int park = k % P;
      SlotScratch& sc = scr[park];
      const uint32_t sl = smem_u32(smem + (size_t)kGSlot * slot_bytes);
      named_sync(5 + park);
      if (mw == 0) OT_TRACE(k, 5);
      const bool grads = a.dtxt != nullptr && !(a.dbg & 1);
      // everything the contractions need from the scratch goes to registers first ...
      const uint32_t Wu = smem_u32(sc.S);
      uint32_t wt[4][4];                          // A = (-W)^t, one fragment per 16 image rows
      uint32_t wa[4];                             // A = -W, rows of this warp's image tile
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        ldsm_x4_t(wt[ks], Wu + (uint32_t)(((ks * 16 + (lane & 7) + ((lane >> 4) << 3)) * kWLd + ((lane >> 3) & 1) * 8) * 2));
      ldsm_x4(wa, Wu + (uint32_t)(((nt * 16 + lrow) * kWLd + lcol) * 2));
      const float ax0 = sc.xn2[g], ax1 = sc.xn2[g + 8];
      const float ay0 = sc.yn2[nt * 16 + g], ay1 = sc.yn2[nt * 16 + g + 8];
      // ... so that the next sample's cost tile can take the scratch over and its solve starts at once
      if (k_next >= 0) {
        mma_bar();
        cost_writeback(k_next);
      }
      named_sync(9);                              // the gradient slot is free (the previous sample's stores have read it)
      if (!grads) {    // forward only: nothing to contract
        named_arrive(8);
        return;
      }
      park_io(kGSlot, park, false);
      mma_bar();      // the slot holds x, y again
      // ---- dx = (-W)^t y + ax x : this warp's D/8 columns, all 64 image rows as K -----------------
      const int dcols = a.D / 8;                  // columns per warp (a multiple of 16 is required)
      const int dc0 = mw * dcols;
      uint32_t yb[4];                             // per K block: smem address of this lane's y row
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const int r = ks * 16 + lrow;
        yb[ks] = (r < a.N ? sl + (uint32_t)((a.M + r) * RS) : zero_u) + lcol * 2;
      }
      // C-fragment addresses of the rows this lane finishes with the element-wise term: rows g and g + 8
      const uint32_t xc0 = (g < a.M ? sl + (uint32_t)(g * RS) : zero_u) + t * 4;
      const uint32_t xc1 = (g + 8 < a.M ? sl + (uint32_t)((g + 8) * RS) : zero_u) + t * 4;
      constexpr int kMaxPairs = 4;                // D <= 512
      uint32_t dxp[kMaxPairs][4];
      const int npairs = dcols / 16;
#pragma unroll
      for (int p = 0; p < kMaxPairs; ++p) {
        if (p < npairs) {
          const uint32_t coff = (uint32_t)((dc0 + p * 16) * 2);
          float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            uint32_t bf[4];
            ldsm_x4_t(bf, yb[ks] + coff);
            mma16816(c0, wt[ks], bf[0], bf[1]);
            mma16816(c1, wt[ks], bf[2], bf[3]);
          }
          const uint32_t x00 = lds32(xc0 + coff), x01 = lds32(xc0 + coff + 16), x10 = lds32(xc1 + coff), x11 = lds32(xc1 + coff + 16);
          dxp[p][0] = pack2(fmaf(ax0, bf_lo(x00), c0[0]), fmaf(ax0, bf_hi(x00), c0[1]));
          dxp[p][1] = pack2(fmaf(ax1, bf_lo(x10), c0[2]), fmaf(ax1, bf_hi(x10), c0[3]));
          dxp[p][2] = pack2(fmaf(ax0, bf_lo(x01), c1[0]), fmaf(ax0, bf_hi(x01), c1[1]));
          dxp[p][3] = pack2(fmaf(ax1, bf_lo(x11), c1[2]), fmaf(ax1, bf_hi(x11), c1[3]));
        }
      }
      if (mw == 0) OT_TRACE(k, 6);
      mma_bar();      // every warp has finished reading y for dx
      // ---- dy = (-W) x + ay y : image rows 16 nt .. +15, half of the columns, in place over y ------
      {
        const uint32_t xb = (lrow < a.M ? sl + (uint32_t)(lrow * RS) : zero_u) + lcol * 2;
        const int yrow = nt * 16 + lrow;
        const uint32_t ywr = (yrow < a.N ? sl + (uint32_t)((a.M + yrow) * RS) : trash_u) + lcol * 2;
        const int r0 = nt * 16 + g, r1 = r0 + 8;
        const uint32_t yc0 = (r0 < a.N ? sl + (uint32_t)((a.M + r0) * RS) : zero_u) + t * 4;
        const uint32_t yc1 = (r1 < a.N ? sl + (uint32_t)((a.M + r1) * RS) : zero_u) + t * 4;
        const int hcols = a.D / 2;
#pragma unroll 2
        for (int p = 0; p < hcols / 16; ++p) {
          const uint32_t coff = (uint32_t)((hf * hcols + p * 16) * 2);
          float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};
          uint32_t bx[4];
          ldsm_x4_t(bx, xb + coff);
          const uint32_t y00 = lds32(yc0 + coff), y01 = lds32(yc0 + coff + 16), y10 = lds32(yc1 + coff), y11 = lds32(yc1 + coff + 16);
          mma16816(c0, wa, bx[0], bx[1]);
          mma16816(c1, wa, bx[2], bx[3]);
          const uint32_t o[4] = {pack2(fmaf(ay0, bf_lo(y00), c0[0]), fmaf(ay0, bf_hi(y00), c0[1])),
                                 pack2(fmaf(ay1, bf_lo(y10), c0[2]), fmaf(ay1, bf_hi(y10), c0[3])),
                                 pack2(fmaf(ay0, bf_lo(y01), c1[0]), fmaf(ay0, bf_hi(y01), c1[1])),
                                 pack2(fmaf(ay1, bf_lo(y11), c1[2]), fmaf(ay1, bf_hi(y11), c1[3]))};
          __syncwarp();          // every lane has read its y words of this pair before the tile is overwritten
          stsm_x4(ywr + coff, o);
        }
      }
      if (mw == 0) OT_TRACE(k, 7);
      mma_bar();      // every warp has finished reading x for dy
      {
        const uint32_t xwr = (lrow < a.M ? sl + (uint32_t)(lrow * RS) : trash_u) + lcol * 2;
#pragma unroll
        for (int p = 0; p < kMaxPairs; ++p)
          if (p < npairs) stsm_x4(xwr + (uint32_t)((dc0 + p * 16) * 2), dxp[p]);
      }
      fence_proxy_async();     // generic writes of the gradients -> visible to the bulk stores
      if (mw == 0) OT_TRACE(k, 8);
      named_arrive(8);
    };

    for (int k = 0; k < P && k < count; ++k) {
      cost_compute(k);
      mma_bar();
      cost_writeback(k);
      cost_park(k);
    }
    for (int k = 0; k < count; ++k) {
      const bool more = k + P < count;
      // The next sample of this park: if its bytes have already landed (they normally have: the load was issued
      // two sample periods ago), its cost tile is computed NOW; otherwise after the gradient job.  One warp decides.
      bool early = false;
      if (more) {
        const int cslot = (k + P) & 1;
        if (mw == 0 && lane == 0) *early_flag = mbar_try_wait(&full[cslot], (ph_full >> cslot) & 1u) ? 1 : 0;
        mma_bar();
        early = *early_flag != 0;
      }
      if (early) cost_compute(k + P);
      grad_job(k, early ? k + P : -1);
      if (more && !early) {
        cost_compute(k + P);
        mma_bar();
        cost_writeback(k + P);
      }
      if (more) cost_park(k + P);             // park k % P is free again: sample k was un-parked above
    }
  }
  // tensor memory goes back before the CTA leaves
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

}  // namespace

size_t ot_fused_smem_bytes(int M, int N, int D, int slots) {
  (void)slots;    // three staging slots and three scratch blocks whatever the number of parks
  const size_t RS = (size_t)D * 2 + 16;
  return (size_t)kMaxSlots * (M + N) * RS + 2 * RS + (size_t)kMaxSlots * sizeof(SlotScratch) +
         kMaxSlots * sizeof(uint64_t) + 16;
}

// parks: how many samples fit in the 512 columns of tensor memory next to each other
int ot_fused_slots(int M, int N, int D) {
  if (D != 256 && D != 512) return 0;
  if (ot_fused_smem_bytes(M, N, D, kMaxSlots) > 232448) return 0;
  const int units = (M + N) * (D / 256);
  const int cols = ((units + 3) / 4 + 7) / 8 * 32;
  const int parks = 512 / cols;
  return parks > kMaxSlots ? kMaxSlots : parks;
}

bool ot_fused_supported(int M, int N, int D, int dtype) {
  if (dtype != CE_BF16 || M < 1 || M > kMP || N < 1 || N > kNR) return false;
  return ot_fused_slots(M, N, D) >= 2;
}

int launch_ot_fused(OtFusedArgs a, cudaStream_t st) {
  a.slots = ot_fused_slots(a.M, a.N, a.D);
  if (a.slots < 1) return fail(CE_ERR_SHAPE, "OT fused: sample does not fit in shared memory");
  const size_t smem = ot_fused_smem_bytes(a.M, a.N, a.D, a.slots);
  CE_CUDA_TRY(cudaFuncSetAttribute(ot_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = a.B < num_sms() ? a.B : num_sms();
  {
    const char* e = getenv("CE_OT_TRACE_PTR");   // debug: tools/ot_trace.py
    a.trace = e != nullptr ? reinterpret_cast<long long*>(strtoull(e, nullptr, 0)) : nullptr;
    const char* pm = getenv("CE_OT_POLL");
    a.poll_mode = pm != nullptr ? atoi(pm) : 0;
    const char* dbg = getenv("CE_OT_DBG");
    a.dbg = dbg != nullptr ? atoi(dbg) : 0;
  }
  ot_fused_kernel<<<grid, kThreads, smem, st>>>(a);
  CE_LAUNCH_CHECK();
  return CE_OK;
}

}  // namespace ce
