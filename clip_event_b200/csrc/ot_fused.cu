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
    if (lane == 0) { while (!mbar_try_wait(bar, parity)) { } }
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
__device__ __forceinline__ void mma_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// Role timeline (tuning aid): with CE_OT_TRACE_PTR set to a device buffer of 64 x 32 int64, CTA 0 records
// clock64 at the hand-over points of the first 64 samples it processes: trace[k * 32 + event].
#define OT_TRACE(k, e)                                                                          \
  do {                                                                                          \
    if (a.trace != nullptr && blockIdx.x == 0 && (k) < 64 && lane == 0) a.trace[(k) * 32 + (e)] = clock64(); \
  } while (0)


__global__ void __launch_bounds__(kThreads, 1) ot_fused_kernel(const OtFusedArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int RS = a.D * 2 + 16;                       // slot row stride in bytes
  const int slot_bytes = (a.M + a.N) * RS;
  const int S = a.slots;
  uint8_t* zero_row = smem + (size_t)S * slot_bytes;  // D*2 bytes of zeros (clamped rows read it)
  uint8_t* trash_row = zero_row + RS;                 // clamped rows write here
  SlotScratch* scr = reinterpret_cast<SlotScratch*>(trash_row + RS);
  uint64_t* bars = reinterpret_cast<uint64_t*>(scr + S);
  uint64_t* full = bars;                  // [S] TMA bytes landed
  uint64_t* s_ready = bars + kMaxSlots;   // [S] cost tile + norms in scratch
  uint64_t* w_ready = bars + 2 * kMaxSlots;   // [S] W, ax, ay in scratch
  uint64_t* out_ready = bars + 3 * kMaxSlots; // [S] gradients in place

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
      for (int s = 0; s < kMaxSlots; ++s) {
        mbar_init(&full[s], 1);
        mbar_init(&s_ready[s], kMmaWarps);
        mbar_init(&w_ready[s], 1);
        mbar_init(&out_ready[s], kMmaWarps);
      }
      mbar_fence_init();
    }
    fence_proxy_async();     // generic zero-fill before the async-proxy loads into the same bytes
    __syncthreads();
  }

  const int row_bytes = a.D * 2;
  if (warp == 0) {
    // ===================================== IO warp ==============================================
    auto issue_load = [&](int k) {
      const int slot = k % S;
      const int64_t b = (int64_t)blockIdx.x + (int64_t)k * gridDim.x;
      uint8_t* sl = smem + (size_t)slot * slot_bytes;
      const uint8_t* xg = reinterpret_cast<const uint8_t*>(a.txt) + b * a.txt_bs * 2;
      const uint8_t* yg = reinterpret_cast<const uint8_t*>(a.img) + b * a.img_bs * 2;
      OT_TRACE(k, 0);
      if (lane == 0) mbar_expect_tx(&full[slot], (uint32_t)((a.M + a.N) * row_bytes));
      __syncwarp();
      for (int r = lane; r < a.M + a.N; r += 32) {
        const uint8_t* src = r < a.M ? xg + (int64_t)r * row_bytes : yg + (int64_t)(r - a.M) * row_bytes;
        bulk_load(sl + (size_t)r * RS, src, (uint32_t)row_bytes, &full[slot]);
      }
    };
    for (int k = 0; k < S && k < count; ++k) issue_load(k);
    for (int k = 0; k < count; ++k) {
      const int slot = k % S;
      const uint32_t ph = (uint32_t)(k / S) & 1u;
      warp_wait(&out_ready[slot], ph, lane, a.poll_mode);
      OT_TRACE(k, 1);
      if (a.dtxt != nullptr) {
        const int64_t b = (int64_t)blockIdx.x + (int64_t)k * gridDim.x;
        uint8_t* sl = smem + (size_t)slot * slot_bytes;
        uint8_t* dxg = reinterpret_cast<uint8_t*>(a.dtxt) + b * a.txt_bs * 2;
        uint8_t* dyg = reinterpret_cast<uint8_t*>(a.dimg) + b * a.img_bs * 2;
        for (int r = lane; r < a.M + a.N; r += 32) {
          uint8_t* dst = r < a.M ? dxg + (int64_t)r * row_bytes : dyg + (int64_t)(r - a.M) * row_bytes;
          bulk_store(dst, sl + (size_t)r * RS, (uint32_t)row_bytes);
        }
        if (a.dslot0 != nullptr && lane == 31)   // the dropped whole-image slot's gradient is zero
          bulk_store(reinterpret_cast<uint8_t*>(a.dslot0) + b * a.img_bs * 2, zero_row, (uint32_t)row_bytes);
        tma_store_commit();
        tma_store_wait_read();     // the slot's bytes have been read out: it may be refilled
      }
      __syncwarp();
      OT_TRACE(k, 2);
      if (k + S < count) issue_load(k + S);
    }
    tma_store_wait_all();
  } else if (warp < 4) {
    // ===================================== IPOT warps ===========================================
    const int slot = warp - 1;
    if (slot < S) {
      SlotScratch& sc = scr[slot];
      const float nib = -1.f / a.beta;
      for (int k = slot; k < count; k += S) {
        const uint32_t ph = (uint32_t)(k / S) & 1u;
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

        warp_wait(&s_ready[slot], ph, lane, a.poll_mode);
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
              a0[j] = (xv && !yp0) ? expf((1.f - sv0[j] * rx * ry0) * nib) : 0.f;
              a1[j] = (xv && !yp1) ? expf((1.f - sv1[j] * rx * ry1) * nib) : 0.f;
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

        // R holds A*T-factor at the top of every iteration: the multiply for the NEXT iteration is issued
        // while this iteration's column sums travel through shared memory
#pragma unroll
        for (int q = 0; q < 8; ++q) { R0[q] = __fmul2_rn(R0[q], A0[q]); R1[q] = __fmul2_rn(R1[q], A1[q]); }
        const int refold = max(1, min(8, (int)(14.f * a.beta)));   // A^refold stays far above the fp32 underflow (C <= 2)
        int since_fold = 0;
        for (int it = 0; it < a.iters; ++it) {
          float z0 = u0, z1 = u1;
          const bool last = it + 1 == a.iters;
          const bool fold = (since_fold + 1 == refold) && !last;
          for (int kk = 0; kk < a.k; ++kk) {
            // row sums: in-thread over the 16 columns
            float2 w2[8];
            {
              const float4* wp = reinterpret_cast<const float4*>(sc.w);
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float4 w4 = wp[q];
                w2[2 * q] = make_float2(w4.x, w4.y); w2[2 * q + 1] = make_float2(w4.z, w4.w);
              }
            }
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
            if (kk + 1 == a.k && !fold && !last) {   // next iteration's multiply by A rides under the transpose
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
            if (kk + 1 < a.k) {
              if (lane < kMP) sc.w[lane] = v_c * sig_c;
              __syncwarp();
            }
          }
          u0 = z0; u1 = z1;
          v_c *= sig_c;
          ++since_fold;
          if (fold) {     // fold the scalings back into R (together with the next multiply by A)
            since_fold = 0;
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
          }
          if (lane < kMP) sc.w[lane] = v_c * sig_c;
          __syncwarp();
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
        __syncwarp();
        OT_TRACE(k, 12);
        if (lane == 0) mbar_arrive(&w_ready[slot]);
      }
    }
  } else {
    // ===================================== MMA warps ============================================
    const int mw = warp - 4;
    const int nt = mw & 3, hf = mw >> 2;
    const int ksteps = a.D / 16;
    const int kh0 = hf * (ksteps / 2), kh1 = hf == 0 ? ksteps / 2 : ksteps;
    const int lrow = lane & 15, lcol = (lane >> 4) * 8;       // ldmatrix / stmatrix lane -> (row, column) of a 16x16 block
    const uint32_t zero_u = smem_u32(zero_row), trash_u = smem_u32(trash_row);

    auto cost_phase = [&](int k) {
      const int slot = k % S;
      const uint32_t ph = (uint32_t)(k / S) & 1u;
      SlotScratch& sc = scr[slot];
      const uint32_t sl = smem_u32(smem + (size_t)slot * slot_bytes);
      // A operand: image-node rows 16 nt .. +15; B operand: the 16 text-node rows (as two n8 tiles)
      const int yrow = nt * 16 + lrow;
      const uint32_t ya = (yrow < a.N ? sl + (uint32_t)((a.M + yrow) * RS) : zero_u) + lcol * 2;
      const int xrow = (lane & 7) + ((lane >> 4) << 3);
      const uint32_t xa = (xrow < a.M ? sl + (uint32_t)(xrow * RS) : zero_u) + ((lane >> 3) & 1) * 16;
      float acc[2][4], yd[2][4], xd[2][4];
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[j][q] = yd[j][q] = xd[j][q] = 0.f;
      warp_wait(&full[slot], ph, lane, a.poll_mode);
      if (mw == 0) OT_TRACE(k, 3);
#pragma unroll 4
      for (int ks = kh0; ks < kh1; ++ks) {
        uint32_t af[4], bf[4];
        ldsm_x4(af, ya + ks * 32);
        ldsm_x4(bf, xa + ks * 32);
        mma16816(acc[0], af, bf[0], bf[1]);
        mma16816(acc[1], af, bf[2], bf[3]);
        mma16816(yd[0], af, af[0], af[2]);     // diagonal of tile * tile^t = row sums of squares
        mma16816(yd[1], af, af[1], af[3]);
        if (nt == 0) {
          const uint32_t xaf[4] = {bf[0], bf[2], bf[1], bf[3]};
          mma16816(xd[0], xaf, xaf[0], xaf[2]);
          mma16816(xd[1], xaf, xaf[1], xaf[3]);
        }
      }
      // hand the tile over: the first K half stores, the second adds
      float* Sp = sc.S + (nt * 16 + g) * kSLd + 2 * t;
      const bool diag_holder = t == (g >> 1);
      const float yn_lo = (g & 1) ? yd[0][1] : yd[0][0], yn_hi = (g & 1) ? yd[1][3] : yd[1][2];
      const float xn_lo = (g & 1) ? xd[0][1] : xd[0][0], xn_hi = (g & 1) ? xd[1][3] : xd[1][2];
      if (hf == 0) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          *reinterpret_cast<float2*>(Sp + 8 * j) = make_float2(acc[j][0], acc[j][1]);
          *reinterpret_cast<float2*>(Sp + 8 * kSLd + 8 * j) = make_float2(acc[j][2], acc[j][3]);
        }
        if (diag_holder) {
          sc.yn2[nt * 16 + g] = yn_lo; sc.yn2[nt * 16 + g + 8] = yn_hi;
          if (nt == 0) { sc.xn2[g] = xn_lo; sc.xn2[g + 8] = xn_hi; }
        }
      }
      mma_bar();
      if (hf == 1) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          atomicAdd(Sp + 8 * j, acc[j][0]); atomicAdd(Sp + 8 * j + 1, acc[j][1]);
          atomicAdd(Sp + 8 * kSLd + 8 * j, acc[j][2]); atomicAdd(Sp + 8 * kSLd + 8 * j + 1, acc[j][3]);
        }
        if (diag_holder) {
          atomicAdd(&sc.yn2[nt * 16 + g], yn_lo); atomicAdd(&sc.yn2[nt * 16 + g + 8], yn_hi);
          if (nt == 0) { atomicAdd(&sc.xn2[g], xn_lo); atomicAdd(&sc.xn2[g + 8], xn_hi); }
        }
      }
      __syncwarp();
      if (mw == 0) OT_TRACE(k, 4);
      if (lane == 0) mbar_arrive(&s_ready[slot]);
    };

    auto grad_phase = [&](int k) {
      const int slot = k % S;
      const uint32_t ph = (uint32_t)(k / S) & 1u;
      SlotScratch& sc = scr[slot];
      const uint32_t sl = smem_u32(smem + (size_t)slot * slot_bytes);
      warp_wait(&w_ready[slot], ph, lane, a.poll_mode);
      if (mw == 0) OT_TRACE(k, 5);
      if (a.dtxt == nullptr) {    // forward only: nothing to contract
        __syncwarp();
        if (lane == 0) mbar_arrive(&out_ready[slot]);
        return;
      }
      const uint32_t Wu = smem_u32(sc.S);
      // ---- dx = (-W)^t y + ax x : this warp's D/8 columns, all 64 image rows as K -----------------
      const int dcols = a.D / 8;                  // columns per warp (a multiple of 16 is required)
      const int dc0 = mw * dcols;
      uint32_t wt[4][4];                          // A = (-W)^t, one fragment per 16 image rows
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        ldsm_x4_t(wt[ks], Wu + (uint32_t)(((ks * 16 + (lane & 7) + ((lane >> 4) << 3)) * kWLd + ((lane >> 3) & 1) * 8) * 2));
      uint32_t axh[4], axl[4];
      {
        const float a0 = sc.xn2[g], a1 = sc.xn2[g + 8];
        const float h0 = bf16r(a0), h1 = bf16r(a1);
        diag_frag(axh, h0, h1, g, t);
        diag_frag(axl, a0 - h0, a1 - h1, g, t);
      }
      uint32_t yb[4];                             // per K block: smem address of this lane's y row
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const int r = ks * 16 + lrow;
        yb[ks] = (r < a.N ? sl + (uint32_t)((a.M + r) * RS) : zero_u) + lcol * 2;
      }
      const uint32_t xb = (lrow < a.M ? sl + (uint32_t)(lrow * RS) : zero_u) + lcol * 2;
      constexpr int kMaxPairs = 6;                // D <= 768
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
          uint32_t bx[4];
          ldsm_x4_t(bx, xb + coff);
          mma16816(c0, axh, bx[0], bx[1]); mma16816(c0, axl, bx[0], bx[1]);
          mma16816(c1, axh, bx[2], bx[3]); mma16816(c1, axl, bx[2], bx[3]);
          dxp[p][0] = pack2(c0[0], c0[1]); dxp[p][1] = pack2(c0[2], c0[3]);
          dxp[p][2] = pack2(c1[0], c1[1]); dxp[p][3] = pack2(c1[2], c1[3]);
        }
      }
      if (mw == 0) OT_TRACE(k, 6);
      mma_bar();      // every warp has finished reading y for dx
      // ---- dy = (-W) x + ay y : image rows 16 nt .. +15, half of the columns, in place over y ------
      {
        uint32_t wa[4];
        ldsm_x4(wa, Wu + (uint32_t)(((nt * 16 + lrow) * kWLd + lcol) * 2));
        uint32_t ayh[4], ayl[4];
        const float a0 = sc.yn2[nt * 16 + g], a1 = sc.yn2[nt * 16 + g + 8];
        const float h0 = bf16r(a0), h1 = bf16r(a1);
        diag_frag(ayh, h0, h1, g, t);
        diag_frag(ayl, a0 - h0, a1 - h1, g, t);
        const int yrow = nt * 16 + lrow;
        const bool yin = yrow < a.N;
        const uint32_t yrd = (yin ? sl + (uint32_t)((a.M + yrow) * RS) : zero_u) + lcol * 2;
        const uint32_t ywr = (yin ? sl + (uint32_t)((a.M + yrow) * RS) : trash_u) + lcol * 2;
        const int hcols = a.D / 2;
#pragma unroll 2
        for (int p = 0; p < hcols / 16; ++p) {
          const uint32_t coff = (uint32_t)((hf * hcols + p * 16) * 2);
          float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};
          uint32_t bx[4], by[4];
          ldsm_x4_t(bx, xb + coff);
          ldsm_x4_t(by, yrd + coff);
          mma16816(c0, wa, bx[0], bx[1]);
          mma16816(c1, wa, bx[2], bx[3]);
          mma16816(c0, ayh, by[0], by[1]); mma16816(c0, ayl, by[0], by[1]);
          mma16816(c1, ayh, by[2], by[3]); mma16816(c1, ayl, by[2], by[3]);
          const uint32_t o[4] = {pack2(c0[0], c0[1]), pack2(c0[2], c0[3]), pack2(c1[0], c1[1]), pack2(c1[2], c1[3])};
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
      __syncwarp();
      if (mw == 0) OT_TRACE(k, 8);
      if (lane == 0) mbar_arrive(&out_ready[slot]);
    };

    // static schedule: the slots run one third of a cycle apart in steady state
    for (int k = 0; k < S && k < count; ++k) cost_phase(k);
    for (int k = 0; k < count; ++k) {
      grad_phase(k);
      if (k + S < count) cost_phase(k + S);
    }
  }
}

}  // namespace

size_t ot_fused_smem_bytes(int M, int N, int D, int slots) {
  const size_t RS = (size_t)D * 2 + 16;
  return (size_t)slots * (M + N) * RS + 2 * RS + (size_t)slots * sizeof(SlotScratch) + 4 * kMaxSlots * sizeof(uint64_t) + 16;
}

int ot_fused_slots(int M, int N, int D) {
  for (int s = kMaxSlots; s >= 1; --s)
    if (ot_fused_smem_bytes(M, N, D, s) <= 232448) return s;
  return 0;
}

bool ot_fused_supported(int M, int N, int D, int dtype) {
  if (dtype != CE_BF16 || M < 1 || M > kMP || N < 1 || N > kNR) return false;
  if (D % 128 != 0 || D > 768) return false;     // D/8 columns per warp in 16-column pairs
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
  }
  ot_fused_kernel<<<grid, kThreads, smem, st>>>(a);
  CE_LAUNCH_CHECK();
  return CE_OK;
}

}  // namespace ce
