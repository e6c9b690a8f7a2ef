// OT graph-alignment loss (IPOT) for sm_100a: cost contraction, shared/register-resident solver,
// gradient contraction.  Reference behaviour: src/clip-event/model_ot.py:8-84 and
// src/clip-event/model_clip.py:679-715 (see include/clip_event_b200.h for the boundary).
//
// Three launches per call, all batched over samples (no cross-sample exchange):
//   ot_cost_kernel  S[b,n,m] = <y_n, x_m>, |x_m|^2, |y_n|^2      reads x, y once   (HBM-bound)
//   ot_ipot_kernel  C = 1 - S/(|x||y|); T = IPOT(C); dist; W, ax, ay   (on-chip, ALU-bound)
//   ot_grad_kernel  dx = -W y + ax*x ; dy = -W^t x + ay*y          reads x, y, writes dx, dy
// where W[m,n] = scale * T[n,m] / (|x_m||y_n|) and ax, ay carry the normalisation backward
// (d x^ = -dC y^,  dx = (dx^ - x^ (x^.dx^)) / |x|,  x^.dx^ = -sum_n dC[m,n] S^[m,n]).
// The contractions run on the tensor cores through warp-level mma (tf32, 3-way split in fp32
// mode so the cost matrix keeps fp32 accuracy: IPOT amplifies cost error by iters/beta).
#include <stdlib.h>

#include <algorithm>

#include "ce_common.cuh"
#include "ot_fused.cuh"

namespace ce {
namespace {


struct OtArgs {
  const void* txt;
  const void* img;
  int64_t txt_bs, img_bs;  // sample strides in elements
  int B, M, N, D;
  int tiles_per_cta;       // 16-row tiles of image nodes handled by one CTA
  float* S;                // [B, N, MP]  raw dots, later W
  float* nx2;              // [B, MP]     |x|^2, later ax
  float* ny2;              // [B, Nld]    |y|^2, later ay
  int Nld;
  void* dtxt;
  void* dimg;
  float* dx_acc;           // [B, M, D] fp32, only when a sample is split over several CTAs
  int nsplit;
};

template <int NSPLIT>
__device__ __forceinline__ void split_tf32(float v, uint32_t& hi, uint32_t& lo) {
  hi = f2tf32(v);
  if constexpr (NSPLIT == 3) lo = f2tf32(v - __uint_as_float(hi));
  else lo = 0;
}

// acc += A*B with the configured number of tf32 products (small cross terms first).
template <int NSPLIT>
__device__ __forceinline__ void mma_split(float* acc, const uint32_t* ahi, const uint32_t* alo,
                                          const uint32_t* bhi, const uint32_t* blo) {
  if constexpr (NSPLIT == 3) {
    mma_tf32(acc, alo, bhi);
    mma_tf32(acc, ahi, blo);
  }
  mma_tf32(acc, ahi, bhi);
}

__device__ __forceinline__ void mma_bf16(float* c, const uint32_t* a, const uint32_t* b) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, "
      "{%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// B fragment (k16 x n8) of a row-major [k][n] bf16 tile: two transposed 8x8 loads.
__device__ __forceinline__ void ldsm_x2_trans(uint32_t* r, const void* row_ptr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];"
               : "=r"(r[0]), "=r"(r[1])
               : "r"(smem_u32(row_ptr)));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t* r, const void* row_ptr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(row_ptr)));
}
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

// A fragment (m16 x k16, bf16) of diag(v0 at rows 0-7, v1 at rows 8-15): thread (g, t) holds the
// diagonal element of row g in a0 and of row g+8 in a3 when 2t or 2t+1 equals g.
__device__ __forceinline__ void diag_frag(uint32_t* af, float v_g, float v_g8, int g, int t) {
  const float z = 0.f;
  af[0] = pack_bf16(2 * t == g ? v_g : z, 2 * t + 1 == g ? v_g : z);
  af[1] = 0u;
  af[2] = 0u;
  af[3] = pack_bf16(2 * t == g ? v_g8 : z, 2 * t + 1 == g ? v_g8 : z);
}
__device__ __forceinline__ float bf16_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

// ---- cp.async ring: every stage holds one 128-byte column slab of [x rows | y rows] ----------
__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool valid) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src),
               "r"(valid ? 16 : 0)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int kSlabBytes = 128;   // bytes of one row staged per step (32 fp32 / 64 bf16 columns)
constexpr int kStages = 2;

// Stage slab `c` of the MP text rows followed by `rows` image rows; RS = smem row stride in bytes.
// A thread always copies the same 16-byte column piece (nthreads is a multiple of 8), so the loop
// only walks rows: no divisions, one bounds test per row.
template <int RS>
__device__ __forceinline__ void issue_slab(uint8_t* stage, const uint8_t* xg, const uint8_t* yg,
                                           int MP, int M, int rows, int rows_valid, int row_bytes,
                                           int c, int nthreads) {
  const int piece = (threadIdx.x & 7) * 16;
  const int col = c * kSlabBytes + piece;
  const bool col_ok = col < row_bytes;
  const int rstep = nthreads >> 3;
  const int r0 = threadIdx.x >> 3;
  {
    const uint8_t* src = xg + (int64_t)r0 * row_bytes + col;
    uint8_t* dst = stage + r0 * RS + piece;
    for (int r = r0; r < MP; r += rstep, src += (int64_t)rstep * row_bytes, dst += rstep * RS) {
      const bool valid = col_ok && r < M;
      cp_async16(dst, valid ? src : xg, valid);
    }
  }
  {
    const uint8_t* src = yg + (int64_t)r0 * row_bytes + col;
    uint8_t* dst = stage + (MP + r0) * RS + piece;
    for (int r = r0; r < rows; r += rstep, src += (int64_t)rstep * row_bytes, dst += rstep * RS) {
      const bool valid = col_ok && r < rows_valid;
      cp_async16(dst, valid ? src : yg, valid);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Kernel A: S = y x^t (raw), row sums of squares.
// grid (B, nsplit); CTA handles image-node tiles [ns*tiles_per_cta, ...); warp w owns tiles
// w, w+NWARPS, ...; text nodes (padded to MP) sit on the mma N axis.  Inputs stream through a
// 3-stage cp.async ring in their own dtype: bf16 feeds m16n8k16 bf16 mma directly, fp32 feeds
// the 3xTF32 path.
// ------------------------------------------------------------------------------------------
constexpr int kRsA = kSlabBytes + 16;   // 36 words: conflict-free fragment reads along K

template <int DT, int MP, int NWARPS, int TPW>
__global__ void __launch_bounds__(NWARPS * 32) ot_cost_kernel(OtArgs a) {
  constexpr bool F32 = DT == CE_F32;
  constexpr int NT = NWARPS * 32;
  constexpr int NJ = MP / 8;
  constexpr int RSW = kRsA / 4;               // row stride in 32-bit words
  constexpr int ESZ = F32 ? 4 : 2;
  extern __shared__ __align__(16) uint8_t smem_a[];
  const int b = blockIdx.x, ns = blockIdx.y;
  const int ntiles = (a.N + 15) / 16;
  const int tile0 = ns * a.tiles_per_cta;
  const int my_tiles = min(a.tiles_per_cta, ntiles - tile0);
  const int row0 = tile0 * 16;
  const int rows = my_tiles * 16;
  const int rows_valid = min(rows, a.N - row0);
  const int stage_bytes = (MP + rows) * kRsA;
  const int row_bytes = a.D * ESZ;
  const uint8_t* xg = reinterpret_cast<const uint8_t*>(a.txt) + (int64_t)b * a.txt_bs * ESZ;
  const uint8_t* yg = reinterpret_cast<const uint8_t*>(a.img) + ((int64_t)b * a.img_bs + (int64_t)row0 * a.D) * ESZ;
  const int w = warp_id(), lane = lane_id(), g = lane >> 2, t = lane & 3;
  const int nslab = (row_bytes + kSlabBytes - 1) / kSlabBytes;

  float acc[TPW][NJ][4];
  float yss[TPW][2];
  float xss[NJ];
  float yd[F32 ? 1 : TPW][2][4];        // bf16: diagonal blocks of tile * tile^t (sums of squares)
  float xd[F32 ? 1 : MP / 16][2][4];
#pragma unroll
  for (int i = 0; i < (F32 ? 1 : TPW); ++i)
#pragma unroll
    for (int c = 0; c < 4; ++c) yd[i][0][c] = yd[i][1][c] = 0.f;
#pragma unroll
  for (int q = 0; q < (F32 ? 1 : MP / 16); ++q)
#pragma unroll
    for (int c = 0; c < 4; ++c) xd[q][0][c] = xd[q][1][c] = 0.f;
#pragma unroll
  for (int i = 0; i < TPW; ++i) {
    yss[i][0] = yss[i][1] = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[i][j][c] = 0.f;
  }
#pragma unroll
  for (int j = 0; j < NJ; ++j) xss[j] = 0.f;

#pragma unroll
  for (int s = 0; s < kStages - 1; ++s) {
    if (s < nslab) issue_slab<kRsA>(smem_a + s * stage_bytes, xg, yg, MP, a.M, rows, rows_valid, row_bytes, s, NT);
    cp_async_commit();
  }
  for (int c = 0; c < nslab; ++c) {
    {
      const int cn = c + kStages - 1;
      if (cn < nslab) issue_slab<kRsA>(smem_a + (cn % kStages) * stage_bytes, xg, yg, MP, a.M, rows, rows_valid, row_bytes, cn, NT);
      cp_async_commit();
    }
    cp_async_wait<kStages - 1>();
    __syncthreads();
    const uint32_t* xs = reinterpret_cast<const uint32_t*>(smem_a + (c % kStages) * stage_bytes);
    const uint32_t* ys = xs + MP * RSW;
    if constexpr (F32) {
      // fp32 mode.  The tensor core truncates its fp32 accumulator after every instruction, a bias
      // that grows with the accumulator's magnitude and the chain length (measured: ~4e-6
      // relative over D = 512) and that IPOT then amplifies by iters/beta.  So each 32-column
      // slab is accumulated from zero (12 instructions, small partial sums) and folded into the
      // running sum with an ordinary round-to-nearest add.
#pragma unroll
      for (int i = 0; i < TPW; ++i) {
        int tile = w + i * NWARPS;
        if (tile < my_tiles) {
          float tacc[NJ][4];
#pragma unroll
          for (int j = 0; j < NJ; ++j)
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) tacc[j][cc] = 0.f;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint32_t* yr = ys + (tile * 16 + g) * RSW + ks * 8 + t;
            float av[4] = {__uint_as_float(yr[0]), __uint_as_float(yr[8 * RSW]), __uint_as_float(yr[4]),
                           __uint_as_float(yr[8 * RSW + 4])};
            yss[i][0] += av[0] * av[0] + av[2] * av[2];
            yss[i][1] += av[1] * av[1] + av[3] * av[3];
            uint32_t ahi[4], alo[4];
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) split_tf32<3>(av[cc], ahi[cc], alo[cc]);
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
              float b0 = __uint_as_float(xs[(8 * j + g) * RSW + ks * 8 + t]);
              float b1 = __uint_as_float(xs[(8 * j + g) * RSW + ks * 8 + t + 4]);
              if (i == 0) xss[j] += b0 * b0 + b1 * b1;
              uint32_t bhi[2], blo[2];
              split_tf32<3>(b0, bhi[0], blo[0]);
              split_tf32<3>(b1, bhi[1], blo[1]);
              mma_split<3>(tacc[j], ahi, alo, bhi, blo);
            }
          }
#pragma unroll
          for (int j = 0; j < NJ; ++j)
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) acc[i][j][cc] += tacc[j][cc];
        }
      }
    } else {
      // bf16 mode: products of bf16 inputs are exact in the fp32 accumulator.  The row sums of
      // squares also come from the tensor core: the diagonal of tile * tile^t, whose B fragments
      // are the tile's own A-fragment registers ({a0,a2} = rows 0-7, {a1,a3} = rows 8-15) -- this
      // replaces 16 unpack+FMA instructions per fragment on the (saturated) FMA/ALU pipes.
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {          // 4 x k16 per 64-column slab
        uint32_t bf[NJ][2];
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          bf[j][0] = xs[(8 * j + g) * RSW + ks * 8 + t];
          bf[j][1] = xs[(8 * j + g) * RSW + ks * 8 + t + 4];
        }
        if (w == 0) {
#pragma unroll
          for (int q = 0; q < MP / 16; ++q) {
            const uint32_t xa[4] = {bf[2 * q][0], bf[2 * q + 1][0], bf[2 * q][1], bf[2 * q + 1][1]};
            const uint32_t b_lo[2] = {xa[0], xa[2]}, b_hi[2] = {xa[1], xa[3]};
            mma_bf16(xd[q][0], xa, b_lo);
            mma_bf16(xd[q][1], xa, b_hi);
          }
        }
#pragma unroll
        for (int i = 0; i < TPW; ++i) {
          int tile = w + i * NWARPS;
          if (tile < my_tiles) {
            const uint32_t* yr = ys + (tile * 16 + g) * RSW + ks * 8 + t;
            uint32_t af[4] = {yr[0], yr[8 * RSW], yr[4], yr[8 * RSW + 4]};
            const uint32_t b_lo[2] = {af[0], af[2]}, b_hi[2] = {af[1], af[3]};
            mma_bf16(yd[i][0], af, b_lo);
            mma_bf16(yd[i][1], af, b_hi);
#pragma unroll
            for (int j = 0; j < NJ; ++j) mma_bf16(acc[i][j], af, bf[j]);
          }
        }
      }
    }
    __syncthreads();
  }

  float* Sg = a.S + ((int64_t)b * a.N + row0) * MP;
#pragma unroll
  for (int i = 0; i < TPW; ++i) {
    int tile = w + i * NWARPS;
    if (tile < my_tiles) {
      int r = tile * 16 + g;
      if constexpr (F32) {
        float s0 = yss[i][0], s1 = yss[i][1];
        s0 += __shfl_xor_sync(0xffffffffu, s0, 1);
        s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
        s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
        s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
        if (t == 0) {
          if (r < rows_valid) a.ny2[(int64_t)b * a.Nld + row0 + r] = s0;
          if (r + 8 < rows_valid) a.ny2[(int64_t)b * a.Nld + row0 + r + 8] = s1;
        }
      } else if (t == (g >> 1)) {   // this thread holds the diagonal elements (g, g) and (g+8, g+8)
        const float s0 = (g & 1) ? yd[i][0][1] : yd[i][0][0];
        const float s1 = (g & 1) ? yd[i][1][3] : yd[i][1][2];
        if (r < rows_valid) a.ny2[(int64_t)b * a.Nld + row0 + r] = s0;
        if (r + 8 < rows_valid) a.ny2[(int64_t)b * a.Nld + row0 + r + 8] = s1;
      }
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        int m = 8 * j + 2 * t;
        if (r < rows_valid)
          *reinterpret_cast<float2*>(Sg + (int64_t)r * MP + m) = make_float2(acc[i][j][0], acc[i][j][1]);
        if (r + 8 < rows_valid)
          *reinterpret_cast<float2*>(Sg + (int64_t)(r + 8) * MP + m) =
              make_float2(acc[i][j][2], acc[i][j][3]);
      }
    }
  }
  if (ns == 0 && w == 0) {
    if constexpr (F32) {
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        float v = xss[j];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        if (t == 0) a.nx2[(int64_t)b * MP + 8 * j + g] = v;
      }
    } else if (t == (g >> 1)) {
#pragma unroll
      for (int q = 0; q < MP / 16; ++q) {
        a.nx2[(int64_t)b * MP + 16 * q + g] = (g & 1) ? xd[q][0][1] : xd[q][0][0];
        a.nx2[(int64_t)b * MP + 16 * q + g + 8] = (g & 1) ? xd[q][1][3] : xd[q][1][2];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Kernel B: IPOT on one sample per CTA, plan and kernel matrix resident in registers.
// Thread layout: TC = MP/4 threads across a row (4 consecutive text nodes each), 256/TC image
// rows per pass, RPT passes.  Row sums: in-thread + log2(TC) shuffles.  Column sums: in-thread
// over RPT rows, shuffles over the rows of a warp, one smem exchange across warps.
// ------------------------------------------------------------------------------------------
struct IpotArgs {
  float* S;         // in: raw dots [B,N,MP]; out: W
  float* nx2;       // in: |x|^2 [B,MP];  out: ax
  float* ny2;       // in: |y|^2 [B,Nld]; out: ay
  const void* txt_mask;
  const void* img_mask;
  int64_t txt_ms, img_ms;
  int mask_kind;
  int B, M, N, Nld;
  float beta, eps, scale;
  int iters, k;
  float* dist;      // [B]
};

__device__ __forceinline__ bool node_is_pad(const void* mask, int kind, int64_t idx) {
  if (kind == CE_MASK_NUM_I64) return reinterpret_cast<const int64_t*>(mask)[idx] == 0;
  return reinterpret_cast<const uint8_t*>(mask)[idx] != 0;
}

// The plan is carried in factorised form T = diag(u) R diag(v): with Q = A*T the reference
// recurrence (model_ot.py:55-61)
//     delta = 1/(y_len * Q sigma + y_guard);  sigma' = 1/(x_len * Q^t delta + x_guard);  T' = delta*Q*sigma'
// becomes R1 = A*R (1 mul / element), rowdot = R1 (v*sigma) (1 fma), coldot = R1^t (delta*u) (1 fma),
// u' = delta*u, v' = v*sigma'  -- 3 element-wise operations per iteration instead of 5, issued as
// packed f32x2 instructions.  Every kRefold iterations u and v are folded back into R so that R
// (which shrinks like A^t) stays far from the fp32 underflow range.
constexpr int kRefold = 4;

__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
// 1/x in one MUFU (<= 1 ulp); the full-precision division's slow path doubled the loop's length
__device__ __forceinline__ float frcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// One launch geometry for both solvers:
//   G    threads cooperating on one sample (32: one warp, no block barrier; 256: one CTA)
//   TC   = MP/4 threads across a row, each owning 4 consecutive text nodes (two float2)
//   RP   = G/TC image rows per pass, RPT passes
// SCHEME 0: every thread sums the warps' column partials itself (1 barrier / iteration)
// SCHEME 1: warp 0 sums them and publishes sigma (2 barriers / iteration, fewer instructions)
// ASM: keep the kernel matrix A in (dynamic) shared memory instead of registers -- for plans too large
// for 2 x RPT x 4 registers per thread (64 x 577).
// ridx / nvalid (optional): compacted list of the sample's valid image rows -- slot k of the thread layout then
// works on row ridx[k]; padded rows are not visited at all (their W rows and ay are zero-filled at the end).
template <int MP, int RPT, int G, int SCHEME = 0, bool ASM = false, bool RAGGED = false>
__device__ __forceinline__ void ipot_body(const IpotArgs& a, const int* ridx, int nvalid) {
  extern __shared__ __align__(16) float s_A[];     // [RPT*RP][MP] when ASM
  constexpr int TC = MP / 4;
  constexpr int RP = G / TC;
  constexpr int BT = G < 256 ? 256 : G;   // threads per CTA
  constexpr int SPC = BT / G;             // samples per CTA
  constexpr int NW = G / 32;          // warps per sample
  static_assert(TC <= 32 && RP >= 1, "layout");
  __shared__ __align__(16) float s_red[2][NW][MP];   // column partials (G > 32 only)
  __shared__ __align__(16) float s_sig[2][MP];
  __shared__ float s_cnt[SPC][2];
  __shared__ float s_red2[NW];
  const int tid = threadIdx.x, lane = tid & 31;
  const int sub = tid / G;                       // sample slot inside the CTA
  const int gt = tid % G;                        // thread inside the sample group
  const int w = gt >> 5;                         // warp inside the sample group
  const int b = blockIdx.x * SPC + sub;
  const bool live = b < a.B;
  const int bb = live ? b : 0;
  const int tc = gt % TC, tr = gt / TC;
  const int m0 = tc * 4;
  auto group_sync = [&]() { if constexpr (G == 32) __syncwarp(); else __syncthreads(); };
  auto row_of = [&](int i) {                     // image row of this thread's pass i (a.N: none)
    const int k = tr + i * RP;
    if constexpr (!RAGGED) return k;
    else return k < nvalid ? ridx[k] : a.N;
  };

  // ---- masks, lengths, inverse norms -------------------------------------------------------
  float rx[4], xg[4];
  float xcount = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int m = m0 + j;
    bool pad = m >= a.M || node_is_pad(a.txt_mask, a.mask_kind, (int64_t)bb * a.txt_ms + m);
    rx[j] = 1.f / fmaxf(sqrtf(a.nx2[(int64_t)bb * MP + m]), a.eps);
    xg[j] = pad ? 1e4f : 0.f;
    if (!pad && tr == 0) xcount += 1.f;
  }
  uint32_t ypad = 0;                                 // bit i: image row of pass i is padding
  float ycount = 0.f;
#pragma unroll
  for (int i = 0; i < RPT; ++i) {
    int n = row_of(i);
    bool pad = n >= a.N || node_is_pad(a.img_mask, a.mask_kind, (int64_t)bb * a.img_ms + n);
    ypad |= (pad ? 1u : 0u) << i;
    if (tc == 0 && !pad) ycount += 1.f;
  }
  auto row_inv_norm = [&](int i) {
    int n = row_of(i);
    float n2 = n < a.N ? a.ny2[(int64_t)bb * a.Nld + n] : 1.f;
    return 1.f / fmaxf(sqrtf(n2), a.eps);
  };
  if (tid < SPC * 2) (&s_cnt[0][0])[tid] = 0.f;
  __syncthreads();
  xcount = warp_sum(xcount);
  ycount = warp_sum(ycount);
  if (lane == 0) { atomicAdd(&s_cnt[sub][0], xcount); atomicAdd(&s_cnt[sub][1], ycount); }
  __syncthreads();
  const float xlen = s_cnt[sub][0], ylen = s_cnt[sub][1];
  float* Sg = a.S + (int64_t)bb * a.N * MP;
  const bool empty = (xlen == 0.f || ylen == 0.f);   // model_ot.py:62: whole plan masked -> 0

  // ---- kernel matrix A = exp(-C/beta), R = 1 on valid pairs ----------------------------------
  float2 A[ASM ? 1 : RPT][2], R[RPT][2];
  float u[RPT];
  constexpr int NG = (RPT + TC - 1) / TC;   // groups of TC rows; lane tc owns row gi*TC + tc of group gi
  float uown[NG], ygown[NG], zown[NG];
#pragma unroll
  for (int gi = 0; gi < NG; ++gi) {
    uown[gi] = 1.f;
    zown[gi] = 1.f;
    const int i = gi * TC + tc;
    ygown[gi] = (i >= RPT || ((ypad >> i) & 1u)) ? 1e4f : 0.f;
  }
  const float nib = -1.f / a.beta;
#pragma unroll
  for (int i = 0; i < RPT; ++i) {
    int n = row_of(i);
    float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n < a.N) s4 = *reinterpret_cast<const float4*>(Sg + (int64_t)n * MP + m0);
    const float sv[4] = {s4.x, s4.y, s4.z, s4.w};
    const float ryi = row_inv_norm(i);
    float av[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      bool valid = !((ypad >> i) & 1u) && xg[j] == 0.f && !empty;
      av[j] = valid ? expf((1.f - sv[j] * rx[j] * ryi) * nib) : 0.f;
    }
    if constexpr (ASM) {
      *reinterpret_cast<float4*>(s_A + (size_t)(tr + i * RP) * MP + m0) = make_float4(av[0], av[1], av[2], av[3]);
    } else {
      A[i][0] = f2(av[0], av[1]); A[i][1] = f2(av[2], av[3]);
    }
    R[i][0] = f2(av[0] != 0.f ? 1.f : 0.f, av[1] != 0.f ? 1.f : 0.f);
    R[i][1] = f2(av[2] != 0.f ? 1.f : 0.f, av[3] != 0.f ? 1.f : 0.f);
    u[i] = 1.f;
  }
  // exp() can only return 0 for a valid pair if C/beta > 87, impossible for C <= 2, beta >= 0.03
  float v[4], sig[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) { v[j] = 1.f; sig[j] = (xg[j] == 0.f && !empty) ? 1.f / xlen : 0.f; }

  // ---- iterations ------------------------------------------------------------------------------
  int flip = 0;
  for (int it = 0; it < a.iters; ++it) {
#pragma unroll
    for (int i = 0; i < RPT; ++i) {           // R1 = A * R
      if constexpr (ASM) {
        const float4 a4 = *reinterpret_cast<const float4*>(s_A + (size_t)(tr + i * RP) * MP + m0);
        R[i][0] = __fmul2_rn(R[i][0], f2(a4.x, a4.y));
        R[i][1] = __fmul2_rn(R[i][1], f2(a4.z, a4.w));
      } else {
        R[i][0] = __fmul2_rn(R[i][0], A[i][0]);
        R[i][1] = __fmul2_rn(R[i][1], A[i][1]);
      }
    }
    float z[RPT];
    for (int kk = 0; kk < a.k; ++kk) {
      const float2 w0 = f2(v[0] * sig[0], v[1] * sig[1]), w1 = f2(v[2] * sig[2], v[3] * sig[3]);
      float2 cs0 = f2(0.f, 0.f), cs1 = f2(0.f, 0.f);
      float rsv[RPT];
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        float2 p = __ffma2_rn(R[i][1], w1, __fmul2_rn(R[i][0], w0));
        rsv[i] = p.x + p.y;
      }
      // Row sums over the TC lanes of a row.  Rows are taken TC at a time: a butterfly that halves
      // the number of live values per step leaves lane tc with the full sum of row (i0 + tc) for
      // TC-1 shuffles per TC rows (instead of TC*log2(TC)); that lane alone computes delta, which is
      // then handed back to the row's lanes with one shuffle per row.
#pragma unroll
      for (int gi = 0; gi < NG; ++gi) {
        constexpr int kNone = 0;
        (void)kNone;
        const int i0 = gi * TC;
        float vals[TC];
#pragma unroll
        for (int q = 0; q < TC; ++q) vals[q] = (i0 + q < RPT) ? rsv[(i0 + q < RPT) ? i0 + q : 0] : 0.f;
#pragma unroll
        for (int half = TC / 2; half >= 1; half >>= 1) {
          const bool upper = (tc & half) != 0;
#pragma unroll
          for (int q = 0; q < half; ++q) {
            float keep = upper ? vals[q + half] : vals[q];
            float send = upper ? vals[q] : vals[q + half];
            vals[q] = keep + __shfl_xor_sync(0xffffffffu, send, half);
          }
        }
        // lane tc now holds the full sum of row i0 + tc; it alone computes that row's delta
        const float uo = uown[gi];
        const float d = frcp(ylen * (uo * vals[0]) + ygown[gi]);
        const float zo = d * uo;
        zown[gi] = zo;
#pragma unroll
        for (int q = 0; q < TC; ++q) {
          if (i0 + q < RPT) {
            float zq = __shfl_sync(0xffffffffu, zo, (lane & ~(TC - 1)) | q);
            z[(i0 + q < RPT) ? i0 + q : 0] = zq;
            const float2 zz = f2(zq, zq);
            cs0 = __ffma2_rn(zz, R[(i0 + q < RPT) ? i0 + q : 0][0], cs0);
            cs1 = __ffma2_rn(zz, R[(i0 + q < RPT) ? i0 + q : 0][1], cs1);
          }
        }
      }
      float cs[4] = {cs0.x, cs0.y, cs1.x, cs1.y};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int o = TC; o < 32; o <<= 1) cs[j] += __shfl_xor_sync(0xffffffffu, cs[j], o);
      }
      if constexpr (G > 32 && SCHEME == 0) {
        if (lane < TC) *reinterpret_cast<float4*>(&s_red[flip][w][m0]) = make_float4(cs[0], cs[1], cs[2], cs[3]);
        __syncthreads();
        float4 t4 = *reinterpret_cast<const float4*>(&s_red[flip][0][m0]);
        float2 ta = f2(t4.x, t4.y), tb = f2(t4.z, t4.w);
#pragma unroll
        for (int ww = 1; ww < NW; ++ww) {
          float4 q4 = *reinterpret_cast<const float4*>(&s_red[flip][ww][m0]);
          ta = __fadd2_rn(ta, f2(q4.x, q4.y));
          tb = __fadd2_rn(tb, f2(q4.z, q4.w));
        }
        cs[0] = ta.x; cs[1] = ta.y; cs[2] = tb.x; cs[3] = tb.y;
        flip ^= 1;
      }
      if constexpr (G > 32 && SCHEME == 1) {
        // v is identical in every thread of a column group, so warp 0 can finish sigma alone
        if (lane < TC) *reinterpret_cast<float4*>(&s_red[flip][w][m0]) = make_float4(cs[0], cs[1], cs[2], cs[3]);
        __syncthreads();
        if (gt < MP) {
          float tsum = 0.f;
#pragma unroll
          for (int ww = 0; ww < NW; ++ww) tsum += s_red[flip][ww][gt];
          s_sig[flip][gt] = tsum;
        }
        __syncthreads();
        float4 t4 = *reinterpret_cast<const float4*>(&s_sig[flip][m0]);
        cs[0] = t4.x; cs[1] = t4.y; cs[2] = t4.z; cs[3] = t4.w;
        flip ^= 1;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) sig[j] = frcp(xlen * (v[j] * cs[j]) + xg[j]);
    }
#pragma unroll
    for (int i = 0; i < RPT; ++i) u[i] = z[i];
#pragma unroll
    for (int gi = 0; gi < NG; ++gi) uown[gi] = zown[gi];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] *= sig[j];
    if ((it % kRefold) == kRefold - 1) {     // fold the scalings back into R
      const float2 v0 = f2(v[0], v[1]), v1 = f2(v[2], v[3]);
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        const float2 uu = f2(u[i], u[i]);
        R[i][0] = __fmul2_rn(__fmul2_rn(R[i][0], uu), v0);
        R[i][1] = __fmul2_rn(__fmul2_rn(R[i][1], uu), v1);
        u[i] = 1.f;
      }
#pragma unroll
      for (int gi = 0; gi < NG; ++gi) uown[gi] = 1.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = 1.f;   // sigma itself is unchanged by the fold
    }
  }

  // ---- distance, W, normalisation-backward coefficients (T = u R v) --------------------------
  // |y|^2 is re-read here (A is dead, registers are free) and must be read by every thread of a
  // row before the row's tc == 0 thread overwrites it with ay.
  float ryv[RPT];
#pragma unroll
  for (int i = 0; i < RPT; ++i) ryv[i] = row_inv_norm(i);
  group_sync();
  float dsum = 0.f;
  float px[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < RPT; ++i) {
    int n = row_of(i);
    float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n < a.N) s4 = *reinterpret_cast<const float4*>(Sg + (int64_t)n * MP + m0);
    const float sv[4] = {s4.x, s4.y, s4.z, s4.w};
    const float rr[4] = {R[i][0].x, R[i][0].y, R[i][1].x, R[i][1].y};
    const float ryi = ryv[i];
    float py = 0.f, wv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      bool valid = !((ypad >> i) & 1u) && xg[j] == 0.f && !empty;
      float shat = sv[j] * rx[j] * ryi;
      float tt = valid ? u[i] * rr[j] * v[j] : 0.f;   // model_ot.py:62 final mask
      dsum += (1.f - shat) * tt;
      float tg = a.scale * tt;
      px[j] += tg * shat;
      py += tg * shat;
      wv[j] = tg * rx[j] * ryi;
    }
#pragma unroll
    for (int o = 1; o < TC; o <<= 1) py += __shfl_xor_sync(0xffffffffu, py, o);
    if (n < a.N && live) {
      *reinterpret_cast<float4*>(Sg + (int64_t)n * MP + m0) = make_float4(wv[0], wv[1], wv[2], wv[3]);
      if (tc == 0) {
        float n2 = a.ny2[(int64_t)b * a.Nld + n];
        // |y| < eps: F.normalize divides by eps and the projection term has no gradient
        a.ny2[(int64_t)b * a.Nld + n] = (sqrtf(n2) >= a.eps) ? py * ryi * ryi : 0.f;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
#pragma unroll
    for (int o = TC; o < 32; o <<= 1) px[j] += __shfl_xor_sync(0xffffffffu, px[j], o);
  }
  dsum = warp_sum(dsum);
  if constexpr (G > 32) {
    __syncthreads();
    if (lane < TC) *reinterpret_cast<float4*>(&s_red[0][w][m0]) = make_float4(px[0], px[1], px[2], px[3]);
    if (lane == 0) s_red2[w] = dsum;
    __syncthreads();
    if (tid < MP) {
      float t = 0.f;
#pragma unroll
      for (int ww = 0; ww < NW; ++ww) t += s_red[0][ww][tid];
      float n2 = a.nx2[(int64_t)b * MP + tid];
      float r = 1.f / fmaxf(sqrtf(n2), a.eps);
      a.nx2[(int64_t)b * MP + tid] = (sqrtf(n2) >= a.eps) ? t * r * r : 0.f;
    }
    if (tid == 0) {
      float t = 0.f;
#pragma unroll
      for (int ww = 0; ww < NW; ++ww) t += s_red2[ww];
      a.dist[b] = t;
    }
  } else {
    if (live) {
      if (tr == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float n2 = a.nx2[(int64_t)b * MP + m0 + j];
          a.nx2[(int64_t)b * MP + m0 + j] = (sqrtf(n2) >= a.eps) ? px[j] * rx[j] * rx[j] : 0.f;
        }
      }
      if (lane == 0) a.dist[b] = dsum;
    }
  }
  if (RAGGED && live) {   // padded rows were not visited: their plan rows and ay are zero
    for (int n = gt; n < a.N; n += G) {
      if (node_is_pad(a.img_mask, a.mask_kind, (int64_t)b * a.img_ms + n)) {
        float4* wrow = reinterpret_cast<float4*>(Sg + (int64_t)n * MP);
#pragma unroll
        for (int j = 0; j < MP / 4; ++j) wrow[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        a.ny2[(int64_t)b * a.Nld + n] = 0.f;
      }
    }
  }
}

template <int MP, int RPT, int G, int MINB = 1, int SCHEME = 0, bool ASM = false>
__global__ void __launch_bounds__((G < 256 ? 256 : G), MINB) ot_ipot_kernel(const __grid_constant__ IpotArgs a) {
  ipot_body<MP, RPT, G, SCHEME, ASM>(a, nullptr, 0);
}

// Ragged node sets (the reference pads them, dataset_voa.py:566-577): one CTA per sample compacts the valid image
// rows and runs the solver body sized for THEM -- a sample with 60 of 257 rows valid takes the 2-pass body, not the
// 9-pass one.  Same arithmetic per valid row; padded rows cost nothing.
template <int MP>
__global__ void __launch_bounds__(256, 2) ot_ipot_ragged_kernel(const __grid_constant__ IpotArgs a) {
  constexpr int RP = 256 / (MP / 4);
  constexpr int kSlots = 9 * RP;
  __shared__ int s_idx[kSlots];
  __shared__ int s_wcnt[(kSlots + 255) / 256][8];
  const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
  const int b = blockIdx.x;
  constexpr int NH = (kSlots + 255) / 256;
  int pos[NH];
  bool val[NH];
#pragma unroll
  for (int h = 0; h < NH; ++h) {
    const int n = h * 256 + tid;
    val[h] = n < a.N && !node_is_pad(a.img_mask, a.mask_kind, (int64_t)b * a.img_ms + n);
    const uint32_t bal = __ballot_sync(0xffffffffu, val[h]);
    pos[h] = __popc(bal & ((1u << lane) - 1u));
    if (lane == 0) s_wcnt[h][wrp] = __popc(bal);
  }
  __syncthreads();
  int nvalid = 0;
#pragma unroll
  for (int h = 0; h < NH; ++h) {
    int before = nvalid;
    for (int ww = 0; ww < 8; ++ww) {
      const int cnt = s_wcnt[h][ww];
      if (ww < wrp) before += cnt;
      nvalid += cnt;
    }
    if (val[h]) s_idx[before + pos[h]] = h * 256 + tid;
  }
  __syncthreads();
  const int np = (nvalid + RP - 1) / RP;
  if (np <= 1) ipot_body<MP, 1, 256, 0, false, true>(a, s_idx, nvalid);
  else if (np <= 2) ipot_body<MP, 2, 256, 0, false, true>(a, s_idx, nvalid);
  else if (np <= 4) ipot_body<MP, 4, 256, 0, false, true>(a, s_idx, nvalid);
  else if (np <= 7) ipot_body<MP, 7, 256, 0, false, true>(a, s_idx, nvalid);
  else ipot_body<MP, 9, 256, 0, false, true>(a, s_idx, nvalid);
}


// Fallback solver for shapes whose plan does not fit the register-resident kernel: same maths,
// kernel matrix and plan live in a global scratch (L2-resident), one CTA per sample.
struct IpotBigArgs {
  IpotArgs a;
  float* scratch;  // [B, 2, N, MP]
  int MP;
};
__global__ void __launch_bounds__(256) ot_ipot_big_kernel(IpotBigArgs p) {
  const IpotArgs& a = p.a;
  const int MP = p.MP, b = blockIdx.x, tid = threadIdx.x;
  extern __shared__ float sm[];
  float* s_rx = sm;                 // MP
  float* s_sigma = s_rx + MP;       // MP
  float* s_xg = s_sigma + MP;       // MP
  float* s_col = s_xg + MP;         // MP
  float* s_ry = s_col + MP;         // N
  float* s_yg = s_ry + a.N;         // N
  float* s_delta = s_yg + a.N;      // N
  __shared__ float s_cnt[2];
  __shared__ float s_dsum;
  if (tid < 2) s_cnt[tid] = 0.f;
  if (tid == 0) s_dsum = 0.f;
  __syncthreads();
  for (int m = tid; m < MP; m += 256) {
    bool pad = m >= a.M || node_is_pad(a.txt_mask, a.mask_kind, (int64_t)b * a.txt_ms + m);
    s_rx[m] = 1.f / fmaxf(sqrtf(a.nx2[(int64_t)b * MP + m]), a.eps);
    s_xg[m] = pad ? 1e4f : 0.f;
    if (!pad) atomicAdd(&s_cnt[0], 1.f);
  }
  for (int n = tid; n < a.N; n += 256) {
    bool pad = node_is_pad(a.img_mask, a.mask_kind, (int64_t)b * a.img_ms + n);
    s_ry[n] = 1.f / fmaxf(sqrtf(a.ny2[(int64_t)b * a.Nld + n]), a.eps);
    s_yg[n] = pad ? 1e4f : 0.f;
    if (!pad) atomicAdd(&s_cnt[1], 1.f);
  }
  __syncthreads();
  const float xlen = s_cnt[0], ylen = s_cnt[1];
  float* Sg = a.S + (int64_t)b * a.N * MP;
  float* Ag = p.scratch + (int64_t)b * 2 * a.N * MP;
  float* Tg = Ag + (int64_t)a.N * MP;
  const int E = a.N * MP;
  if (xlen == 0.f || ylen == 0.f) {
    for (int e = tid; e < E; e += 256) Sg[e] = 0.f;
    for (int n = tid; n < a.N; n += 256) a.ny2[(int64_t)b * a.Nld + n] = 0.f;
    for (int m = tid; m < MP; m += 256) a.nx2[(int64_t)b * MP + m] = 0.f;
    if (tid == 0) a.dist[b] = 0.f;
    return;
  }
  const float nib = -1.f / a.beta;
  for (int e = tid; e < E; e += 256) {
    int n = e / MP, m = e % MP;
    bool valid = s_yg[n] == 0.f && s_xg[m] == 0.f;
    float c = 1.f - Sg[e] * s_rx[m] * s_ry[n];
    Ag[e] = valid ? expf(c * nib) : 0.f;
    Tg[e] = valid ? 1.f : 0.f;
  }
  for (int m = tid; m < MP; m += 256) s_sigma[m] = s_xg[m] == 0.f ? 1.f / xlen : 0.f;
  __syncthreads();
  const int w = tid >> 5, lane = tid & 31;
  for (int it = 0; it < a.iters; ++it) {
    for (int e = tid; e < E; e += 256) Tg[e] *= Ag[e];
    __syncthreads();
    for (int kk = 0; kk < a.k; ++kk) {
      for (int n = w; n < a.N; n += 8) {
        float rs = 0.f;
        for (int m = lane; m < MP; m += 32) rs += Tg[n * MP + m] * s_sigma[m];
        rs = warp_sum(rs);
        if (lane == 0) s_delta[n] = 1.f / (ylen * rs + s_yg[n]);
      }
      __syncthreads();
      for (int m = w; m < MP; m += 8) {
        float cs = 0.f;
        for (int n = lane; n < a.N; n += 32) cs += s_delta[n] * Tg[n * MP + m];
        cs = warp_sum(cs);
        if (lane == 0) s_sigma[m] = 1.f / (xlen * cs + s_xg[m]);
      }
      __syncthreads();
    }
    for (int e = tid; e < E; e += 256) {
      int n = e / MP, m = e % MP;
      Tg[e] = s_delta[n] * Tg[e] * s_sigma[m];
    }
    __syncthreads();
  }
  for (int m = tid; m < MP; m += 256) s_col[m] = 0.f;
  for (int n = tid; n < a.N; n += 256) s_delta[n] = 0.f;
  __syncthreads();
  float dsum = 0.f;
  for (int e = tid; e < E; e += 256) {
    int n = e / MP, m = e % MP;
    bool valid = s_yg[n] == 0.f && s_xg[m] == 0.f;
    float shat = Sg[e] * s_rx[m] * s_ry[n];
    float tt = valid ? Tg[e] : 0.f;
    dsum += (1.f - shat) * tt;
    float tg = a.scale * tt;
    atomicAdd(&s_col[m], tg * shat);
    atomicAdd(&s_delta[n], tg * shat);
    Sg[e] = tg * s_rx[m] * s_ry[n];
  }
  dsum = warp_sum(dsum);
  if (lane == 0) atomicAdd(&s_dsum, dsum);
  __syncthreads();
  for (int m = tid; m < MP; m += 256) {
    float n2 = a.nx2[(int64_t)b * MP + m];
    a.nx2[(int64_t)b * MP + m] = (sqrtf(n2) >= a.eps) ? s_col[m] * s_rx[m] * s_rx[m] : 0.f;
  }
  for (int n = tid; n < a.N; n += 256) {
    float n2 = a.ny2[(int64_t)b * a.Nld + n];
    a.ny2[(int64_t)b * a.Nld + n] = (sqrtf(n2) >= a.eps) ? s_delta[n] * s_ry[n] * s_ry[n] : 0.f;
  }
  if (tid == 0) a.dist[b] = s_dsum;
}

// ------------------------------------------------------------------------------------------
// Kernel C: dy = -W^t x + ay*y (rows of this CTA), dx = -W y + ax*x (partial over its rows).
// Same cp.async ring as kernel A.  fp32: 3xTF32 on fp32 slabs.  bf16: bf16 mma; W is kept in
// smem as bf16 in both orientations, the slabs' "across K" operands come through ldmatrix.trans.
// ------------------------------------------------------------------------------------------
template <int DT>
__device__ __forceinline__ void store2(typename In<DT>::type* p, float v0, float v1) {
  if constexpr (DT == CE_F32) {
    *reinterpret_cast<float2*>(p) = make_float2(v0, v1);
  } else {
    *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(v0, v1);
  }
}

template <int DT>
struct GradCfg {
  static constexpr bool F32 = DT == CE_F32;
  static constexpr int RS = kSlabBytes + (F32 ? 32 : 16);   // fp32: 40 words (== 8 mod 32); bf16: 36 words
  static constexpr int DCOLS = kSlabBytes / (F32 ? 4 : 2);  // columns per slab: 32 / 64
  // W storage: fp32 [rows][MP+4] floats, or bf16 [rows][MP+8] (both orientations of the bf16 mma
  // A operand are read from it: plain loads for W^t, ldmatrix.trans for W)
  __host__ __device__ static size_t w_bytes(int MP, int rows) {
    return F32 ? sizeof(float) * (size_t)rows * (MP + 4) : 2 * (size_t)rows * (MP + 8);
  }
  __host__ __device__ static size_t smem_bytes(int MP, int rows) {
    return (size_t)kStages * (MP + rows) * RS + w_bytes(MP, rows) + sizeof(float) * (MP + rows) + 64;
  }
};

template <int DT, int MP, int NWARPS, int TPW>
__global__ void __launch_bounds__(NWARPS * 32) ot_grad_kernel(OtArgs a) {
  using Cfg = GradCfg<DT>;
  constexpr bool F32 = Cfg::F32;
  constexpr int NT = NWARPS * 32;
  constexpr int RS = Cfg::RS, RSW = RS / 4;
  constexpr int DC = Cfg::DCOLS;
  constexpr int ESZ = F32 ? 4 : 2;
  constexpr int NOUT = (MP / 16) * (DC / 8);             // dx output tiles per slab
  constexpr int OPW = (NOUT + NWARPS - 1) / NWARPS;
  using T = typename In<DT>::type;
  extern __shared__ __align__(16) uint8_t smem_g[];
  const int b = blockIdx.x, ns = blockIdx.y;
  const int ntiles = (a.N + 15) / 16;
  const int tile0 = ns * a.tiles_per_cta;
  const int my_tiles = min(a.tiles_per_cta, ntiles - tile0);
  const int row0 = tile0 * 16;
  const int rows = my_tiles * 16;
  const int rows_valid = min(rows, a.N - row0);
  const int stage_bytes = (MP + rows) * RS;
  const int row_bytes = a.D * ESZ;
  uint8_t* wbase = smem_g + kStages * stage_bytes;
  float* axs = reinterpret_cast<float*>(wbase + Cfg::w_bytes(MP, rows));
  float* ays = axs + MP;
  const uint8_t* xg = reinterpret_cast<const uint8_t*>(a.txt) + (int64_t)b * a.txt_bs * ESZ;
  const uint8_t* yg = reinterpret_cast<const uint8_t*>(a.img) + ((int64_t)b * a.img_bs + (int64_t)row0 * a.D) * ESZ;
  T* dxg = reinterpret_cast<T*>(a.dtxt) + (int64_t)b * a.txt_bs;
  T* dyg = reinterpret_cast<T*>(a.dimg) + (int64_t)b * a.img_bs + (int64_t)row0 * a.D;
  const int w = warp_id(), lane = lane_id(), g = lane >> 2, t = lane & 3;
  const int nslab = (row_bytes + kSlabBytes - 1) / kSlabBytes;

#pragma unroll
  for (int s = 0; s < kStages - 1; ++s) {
    if (s < nslab) issue_slab<RS>(smem_g + s * stage_bytes, xg, yg, MP, a.M, rows, rows_valid, row_bytes, s, NT);
    cp_async_commit();
  }

  // W, ax, ay of this CTA's rows (overlaps the first slabs' flight)
  constexpr int LDW = MP + 4;                 // fp32 W row stride (floats)
  constexpr int LDWB = MP + 8;                // bf16 W[n][m] row stride (elements)
  float* Ws = reinterpret_cast<float*>(wbase);
  __nv_bfloat16* Wb = reinterpret_cast<__nv_bfloat16*>(wbase);
  {
    const float* Wg = a.S + ((int64_t)b * a.N + row0) * MP;
    for (int idx = threadIdx.x; idx < rows * (MP / 4); idx += NT) {
      int r = idx / (MP / 4), c = (idx % (MP / 4)) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < rows_valid) v = *reinterpret_cast<const float4*>(Wg + (int64_t)r * MP + c);
      if constexpr (F32) {
        *reinterpret_cast<float4*>(Ws + r * LDW + c) = v;
      } else {
        uint2 pk = make_uint2(pack_bf16(-v.x, -v.y), pack_bf16(-v.z, -v.w));   // -W: no sign flip later
        *reinterpret_cast<uint2*>(Wb + r * LDWB + c) = pk;
      }
    }
    for (int m = threadIdx.x; m < MP; m += NT) axs[m] = a.nx2[(int64_t)b * MP + m];
    for (int r = threadIdx.x; r < rows; r += NT)
      ays[r] = r < rows_valid ? a.ny2[(int64_t)b * a.Nld + row0 + r] : 0.f;
  }

  for (int c = 0; c < nslab; ++c) {
    {
      const int cn = c + kStages - 1;
      if (cn < nslab) issue_slab<RS>(smem_g + (cn % kStages) * stage_bytes, xg, yg, MP, a.M, rows, rows_valid, row_bytes, cn, NT);
      cp_async_commit();
    }
    cp_async_wait<kStages - 1>();
    __syncthreads();
    const uint8_t* stage = smem_g + (c % kStages) * stage_bytes;
    const uint32_t* xs = reinterpret_cast<const uint32_t*>(stage);
    const uint32_t* ys = xs + MP * RSW;
    const int d0 = c * DC;

    // ---- dy tiles: rows of this warp, K = text nodes -------------------------------------
#pragma unroll
    for (int i = 0; i < TPW; ++i) {
      int tile = w + i * NWARPS;
      if (tile < my_tiles) {
        float acc[DC / 8][4];
#pragma unroll
        for (int j = 0; j < DC / 8; ++j)
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) acc[j][cc] = 0.f;
        if constexpr (F32) {
#pragma unroll
          for (int ks = 0; ks < MP / 8; ++ks) {
            const float* wr = Ws + (tile * 16 + g) * LDW + ks * 8 + t;
            uint32_t ahi[4], alo[4];
            split_tf32<3>(wr[0], ahi[0], alo[0]);
            split_tf32<3>(wr[8 * LDW], ahi[1], alo[1]);
            split_tf32<3>(wr[4], ahi[2], alo[2]);
            split_tf32<3>(wr[8 * LDW + 4], ahi[3], alo[3]);
#pragma unroll
            for (int j = 0; j < DC / 8; ++j) {
              uint32_t bhi[2], blo[2];
              split_tf32<3>(__uint_as_float(xs[(ks * 8 + t) * RSW + 8 * j + g]), bhi[0], blo[0]);
              split_tf32<3>(__uint_as_float(xs[(ks * 8 + t + 4) * RSW + 8 * j + g]), bhi[1], blo[1]);
              mma_split<3>(acc[j], ahi, alo, bhi, blo);
            }
          }
        } else {
          // acc = (-W^t) x + diag(ay) y : the elementwise term rides on the tensor core as two more
          // k-steps (ay split into bf16 hi + lo), so the epilogue below is only pack + store
#pragma unroll
          for (int ks = 0; ks < MP / 16; ++ks) {
            const uint32_t* wr = reinterpret_cast<const uint32_t*>(Wb + (tile * 16 + g) * LDWB + ks * 16) + t;
            uint32_t af[4] = {wr[0], wr[8 * (LDWB / 2)], wr[4], wr[8 * (LDWB / 2) + 4]};
#pragma unroll
            for (int j = 0; j < DC / 8; ++j) {
              uint32_t bf[2];
              ldsm_x2_trans(bf, stage + (ks * 16 + (lane & 15)) * RS + j * 16);
              mma_bf16(acc[j], af, bf);
            }
          }
          const float ay0 = ays[tile * 16 + g], ay1 = ays[tile * 16 + g + 8];
          const float ay0h = bf16_round(ay0), ay1h = bf16_round(ay1);
          uint32_t dh[4], dl[4];
          diag_frag(dh, ay0h, ay1h, g, t);
          diag_frag(dl, ay0 - ay0h, ay1 - ay1h, g, t);
          const uint8_t* ytile = stage + (MP + tile * 16 + (lane & 15)) * RS;
#pragma unroll
          for (int j = 0; j < DC / 8; ++j) {
            uint32_t bf[2];
            ldsm_x2_trans(bf, ytile + j * 16);
            mma_bf16(acc[j], dh, bf);
            mma_bf16(acc[j], dl, bf);
          }
        }
        int r = tile * 16 + g;
        if constexpr (!F32) {
          T* p0 = dyg + (int64_t)r * a.D + d0 + 2 * t;
          T* p1 = p0 + (int64_t)8 * a.D;
          if (tile * 16 + 16 <= rows_valid && d0 + DC <= a.D) {     // interior tile: no per-store tests
#pragma unroll
            for (int j = 0; j < DC / 8; ++j) {
              store2<DT>(p0 + 8 * j, acc[j][0], acc[j][1]);
              store2<DT>(p1 + 8 * j, acc[j][2], acc[j][3]);
            }
          } else {
#pragma unroll
            for (int j = 0; j < DC / 8; ++j) {
              if (d0 + 8 * j + 2 * t < a.D) {
                if (r < rows_valid) store2<DT>(p0 + 8 * j, acc[j][0], acc[j][1]);
                if (r + 8 < rows_valid) store2<DT>(p1 + 8 * j, acc[j][2], acc[j][3]);
              }
            }
          }
        } else {
#pragma unroll
        for (int j = 0; j < DC / 8; ++j) {
          int d = 8 * j + 2 * t;
          if (d0 + d < a.D) {
            float y00 = __uint_as_float(ys[r * RSW + d]), y01 = __uint_as_float(ys[r * RSW + d + 1]);
            float y10 = __uint_as_float(ys[(r + 8) * RSW + d]), y11 = __uint_as_float(ys[(r + 8) * RSW + d + 1]);
            if (r < rows_valid) {
              float ay = ays[r];
              store2<DT>(dyg + (int64_t)r * a.D + d0 + d, ay * y00 - acc[j][0], ay * y01 - acc[j][1]);
            }
            if (r + 8 < rows_valid) {
              float ay = ays[r + 8];
              store2<DT>(dyg + (int64_t)(r + 8) * a.D + d0 + d, ay * y10 - acc[j][2], ay * y11 - acc[j][3]);
            }
          }
        }
        }
      }
    }

    // ---- dx tiles: outputs split over warps, K = this CTA's image rows --------------------
#pragma unroll
    for (int o = 0; o < OPW; ++o) {
      int ot = w + o * NWARPS;
      if (ot < NOUT) {
        int mi = ot / (DC / 8), j = ot % (DC / 8);
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        if constexpr (F32) {
          for (int ks = 0; ks < rows / 8; ++ks) {
            const float* wr = Ws + (ks * 8 + t) * LDW + 16 * mi + g;
            uint32_t ahi[4], alo[4];
            split_tf32<3>(wr[0], ahi[0], alo[0]);
            split_tf32<3>(wr[8], ahi[1], alo[1]);
            split_tf32<3>(wr[4 * LDW], ahi[2], alo[2]);
            split_tf32<3>(wr[4 * LDW + 8], ahi[3], alo[3]);
            uint32_t bhi[2], blo[2];
            split_tf32<3>(__uint_as_float(ys[(ks * 8 + t) * RSW + 8 * j + g]), bhi[0], blo[0]);
            split_tf32<3>(__uint_as_float(ys[(ks * 8 + t + 4) * RSW + 8 * j + g]), bhi[1], blo[1]);
            mma_split<3>(acc, ahi, alo, bhi, blo);
          }
        } else {
          const uint8_t* ysb = stage + MP * RS;
          // A = W[m][n] out of Wb[n][m]: four transposed 8x8 blocks (m lo/hi x k lo/hi)
          const int blk = lane >> 3, br = lane & 7;
          const __nv_bfloat16* wrow = Wb + (br + (blk >> 1) * 8) * LDWB + 16 * mi + (blk & 1) * 8;
          for (int ks = 0; ks < rows / 16; ++ks) {
            uint32_t af[4];
            ldsm_x4_trans(af, wrow + ks * 16 * LDWB);
            uint32_t bf[2];
            ldsm_x2_trans(bf, ysb + (ks * 16 + (lane & 15)) * RS + j * 16);
            mma_bf16(acc, af, bf);
          }
        }
        int m = 16 * mi + g, d = 8 * j + 2 * t;
        if (d0 + d < a.D) {
          float x00, x01, x10, x11;
          if constexpr (F32) {
            x00 = __uint_as_float(xs[m * RSW + d]); x01 = __uint_as_float(xs[m * RSW + d + 1]);
            x10 = __uint_as_float(xs[(m + 8) * RSW + d]); x11 = __uint_as_float(xs[(m + 8) * RSW + d + 1]);
          } else {
            uint32_t u0 = xs[m * RSW + d / 2], u1 = xs[(m + 8) * RSW + d / 2];
            x00 = bf_lo(u0); x01 = bf_hi(u0); x10 = bf_lo(u1); x11 = bf_hi(u1);
          }
          constexpr float sgn = F32 ? -1.f : 1.f;   // the bf16 copy of W already carries the minus sign
          if (a.nsplit == 1) {
            if (m < a.M)
              store2<DT>(dxg + (int64_t)m * a.D + d0 + d, fmaf(axs[m], x00, sgn * acc[0]), fmaf(axs[m], x01, sgn * acc[1]));
            if (m + 8 < a.M)
              store2<DT>(dxg + (int64_t)(m + 8) * a.D + d0 + d, fmaf(axs[m + 8], x10, sgn * acc[2]),
                         fmaf(axs[m + 8], x11, sgn * acc[3]));
          } else {
            float* dacc = a.dx_acc + ((int64_t)b * a.M) * a.D + d0 + d;
            if (m < a.M) {
              atomicAdd(dacc + (int64_t)m * a.D, sgn * acc[0]);
              atomicAdd(dacc + (int64_t)m * a.D + 1, sgn * acc[1]);
            }
            if (m + 8 < a.M) {
              atomicAdd(dacc + (int64_t)(m + 8) * a.D, sgn * acc[2]);
              atomicAdd(dacc + (int64_t)(m + 8) * a.D + 1, sgn * acc[3]);
            }
          }
        }
      }
    }
    __syncthreads();
  }
}

// dx = dx_acc + ax * x for samples that were split over several CTAs.
template <int DT>
__global__ void ot_dx_finish_kernel(OtArgs a, int MP) {
  using T = typename In<DT>::type;
  int64_t total = (int64_t)a.B * a.M * a.D;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int d = (int)(i % a.D);
    int m = (int)((i / a.D) % a.M);
    int b = (int)(i / ((int64_t)a.D * a.M));
    const T* xg = reinterpret_cast<const T*>(a.txt) + (int64_t)b * a.txt_bs;
    T* dxg = reinterpret_cast<T*>(a.dtxt) + (int64_t)b * a.txt_bs;
    float v = a.dx_acc[i] + a.nx2[(int64_t)b * MP + m] * In<DT>::ld(xg + (int64_t)m * a.D + d);
    In<DT>::st(dxg + (int64_t)m * a.D + d, v);
  }
}

// loss = scale * sum_b dist[b] (left to right like the reference's Python sum, model_clip.py:707),
// and zero-fill of the dropped whole-image slot's gradient.
template <int DT>
__global__ void ot_tail_kernel(const float* dist, int B, float scale, float* loss, void* slot0,
                               int64_t bs, int D) {
  using T = typename In<DT>::type;
  if (blockIdx.x == 0 && threadIdx.x < 32 && loss != nullptr) {
    float s = 0.f;
    for (int b = threadIdx.x; b < B; b += 32) s += dist[b];
    s = warp_sum(s);
    if (threadIdx.x == 0) *loss = s * scale;
  }
  if (slot0 != nullptr) {   // rows are 16-byte aligned multiples of 16 bytes (checked by the caller)
    const int ppr = D * (int)sizeof(T) / 16;          // 16-byte pieces per row
    const int64_t total = (int64_t)B * ppr;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
      const int b = (int)(i / ppr), pc = (int)(i - (int64_t)b * ppr);
      reinterpret_cast<uint4*>(reinterpret_cast<T*>(slot0) + (int64_t)b * bs)[pc] = make_uint4(0u, 0u, 0u, 0u);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Building blocks with the reference's granularity (plain SIMT; not on the fused path)
// ------------------------------------------------------------------------------------------
template <int DT>
__global__ void ot_cost_matrix_kernel(const void* x, const void* y, int B, int M, int N, int D,
                                      float eps, float* cost) {
  // one warp per (b, m, n) triple
  using T = typename In<DT>::type;
  int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  int64_t total = (int64_t)B * M * N;
  if (gw >= total) return;
  int n = (int)(gw % N), m = (int)((gw / N) % M), b = (int)(gw / ((int64_t)N * M));
  const T* xr = reinterpret_cast<const T*>(x) + ((int64_t)b * M + m) * D;
  const T* yr = reinterpret_cast<const T*>(y) + ((int64_t)b * N + n) * D;
  float dot = 0.f, xx = 0.f, yy = 0.f;
  for (int d = lane; d < D; d += 32) {
    float xv = In<DT>::ld(xr + d), yv = In<DT>::ld(yr + d);
    dot += xv * yv; xx += xv * xv; yy += yv * yv;
  }
  dot = warp_sum(dot); xx = warp_sum(xx); yy = warp_sum(yy);
  if (lane == 0) cost[gw] = 1.f - dot / (fmaxf(sqrtf(xx), eps) * fmaxf(sqrtf(yy), eps));
}

__global__ void ot_trace_kernel(const float* x, int B, int n, float* out) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float s = 0.f;
  for (int i = 0; i < n; ++i) s += x[((int64_t)b * n + i) * n + i];
  out[b] = s;
}

// ipot() with the reference's signature: C [B,M,N] (already masked) -> T [B,N,M].
__global__ void __launch_bounds__(256) ot_ipot_plain_kernel(const float* cost, const uint8_t* xpad,
                                                            const uint8_t* ypad, int B, int M, int N,
                                                            float beta, int iters, int k,
                                                            float* plan) {
  extern __shared__ float sm[];
  float* sigma = sm;          // M
  float* delta = sigma + M;   // N
  float* xg = delta + N;      // M
  float* yg = xg + M;         // N
  __shared__ float cnt[2];
  const int b = blockIdx.x, tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  if (tid < 2) cnt[tid] = 0.f;
  __syncthreads();
  for (int m = tid; m < M; m += 256) { bool p = xpad[(int64_t)b * M + m]; xg[m] = p ? 1e4f : 0.f; if (!p) atomicAdd(&cnt[0], 1.f); }
  for (int n = tid; n < N; n += 256) { bool p = ypad[(int64_t)b * N + n]; yg[n] = p ? 1e4f : 0.f; if (!p) atomicAdd(&cnt[1], 1.f); }
  __syncthreads();
  const float xlen = cnt[0], ylen = cnt[1];
  const float* C = cost + (int64_t)b * M * N;
  float* T = plan + (int64_t)b * N * M;   // [N, M]; holds Q during an iteration
  for (int e = tid; e < N * M; e += 256) {
    int n = e / M, m = e % M;
    T[e] = (xg[m] == 0.f && yg[n] == 0.f) ? 1.f : 0.f;
  }
  for (int m = tid; m < M; m += 256) sigma[m] = xg[m] == 0.f ? 1.f / xlen : 0.f;
  __syncthreads();
  for (int it = 0; it < iters; ++it) {
    for (int e = tid; e < N * M; e += 256) {
      int n = e / M, m = e % M;
      float a = (xg[m] == 0.f && yg[n] == 0.f) ? expf(-C[(int64_t)m * N + n] / beta) : 0.f;
      T[e] *= a;
    }
    __syncthreads();
    for (int kk = 0; kk < k; ++kk) {
      for (int n = w; n < N; n += 8) {
        float rs = 0.f;
        for (int m = lane; m < M; m += 32) rs += T[n * M + m] * sigma[m];
        rs = warp_sum(rs);
        if (lane == 0) delta[n] = 1.f / (ylen * rs + yg[n]);
      }
      __syncthreads();
      for (int m = w; m < M; m += 8) {
        float cs = 0.f;
        for (int n = lane; n < N; n += 32) cs += delta[n] * T[n * M + m];
        cs = warp_sum(cs);
        if (lane == 0) sigma[m] = 1.f / (xlen * cs + xg[m]);
      }
      __syncthreads();
    }
    for (int e = tid; e < N * M; e += 256) {
      int n = e / M, m = e % M;
      T[e] = delta[n] * T[e] * sigma[m];
    }
    __syncthreads();
  }
  for (int e = tid; e < N * M; e += 256) {
    int n = e / M, m = e % M;
    if (!(xg[m] == 0.f && yg[n] == 0.f)) T[e] = 0.f;
  }
}

template <int DT>
__global__ void scale_inplace_kernel(void* x, int64_t rows, int64_t row_len, int64_t row_stride,
                                     const float* g, const float* g_same) {
  using T = typename In<DT>::type;
  float s = *g;
  // gradients that were formed for EQUAL upstream gradients of two losses: anything else must not
  // pass silently
  if (g_same != nullptr && *g_same != s) s = __int_as_float(0x7fc00000);
  if (s == 1.f) return;
  int64_t total = rows * row_len;
  T* p = reinterpret_cast<T*>(x);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / row_len, c = i % row_len;
    T* q = p + r * row_stride + c;
    In<DT>::st(q, In<DT>::ld(q) * s);
  }
}

// ------------------------------------------------------------------------------------------
// host dispatch
// ------------------------------------------------------------------------------------------
struct OtPlan {
  int MP, nwarps, tpw, tiles_per_cta, nsplit, rpt, rows;
};

int make_plan(int M, int N, OtPlan* p) {
  if (M < 1 || M > 64) return fail(CE_ERR_SHAPE, "OT: text nodes M=%d outside 1..64", M);
  if (N < 1 || N > 1024) return fail(CE_ERR_SHAPE, "OT: image nodes N=%d outside 1..1024", N);
  p->MP = M <= 16 ? 16 : (M <= 32 ? 32 : 64);
  int ntiles = (N + 15) / 16;
  if (ntiles <= 4) { p->nwarps = 4; p->tpw = 1; }
  else { p->nwarps = 8; p->tpw = p->MP == 64 ? 2 : 3; }
  int cap = p->nwarps * p->tpw;
  {   // tuning aid: CE_OT_CAP limits the 16-row tiles per CTA (more, smaller CTAs per sample)
    static const int cap_env = [] { const char* e = getenv("CE_OT_CAP"); return e != nullptr ? atoi(e) : 0; }();
    if (cap_env > 0 && cap_env < cap) cap = cap_env;
  }
  p->nsplit = (ntiles + cap - 1) / cap;
  p->tiles_per_cta = (ntiles + p->nsplit - 1) / p->nsplit;
  int rows = p->tiles_per_cta * 16;
  p->rows = rows;
  int rp = 256 / (p->MP / 4);
  p->rpt = (N + rp - 1) / rp;
  return CE_OK;
}

template <int DT, int MP, int NW, int TPW>
int launch_cost_grad(bool grad, const OtArgs& a, const OtPlan& p, cudaStream_t st) {
  dim3 grid(a.B, p.nsplit);
  if (!grad) {
    const size_t smem = (size_t)kStages * (MP + p.rows) * kRsA;
    auto kern = ot_cost_kernel<DT, MP, NW, TPW>;
    CE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, NW * 32, smem, st>>>(a);
  } else {
    const size_t smem = GradCfg<DT>::smem_bytes(MP, p.rows);
    if (smem > 227 * 1024) return fail(CE_ERR_SHAPE, "OT: gradient kernel needs %zu B of shared memory", smem);
    auto kern = ot_grad_kernel<DT, MP, NW, TPW>;
    CE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, NW * 32, smem, st>>>(a);
  }
  CE_LAUNCH_CHECK();
  return CE_OK;
}

template <int DT, int MP>
int dispatch_cfg(bool grad, const OtArgs& a, const OtPlan& p, cudaStream_t st) {
  if (p.nwarps == 4) return launch_cost_grad<DT, MP, 4, 1>(grad, a, p, st);
  return launch_cost_grad<DT, MP, 8, (MP == 64 ? 2 : 3)>(grad, a, p, st);
}

template <int DT>
int dispatch_mp(bool grad, const OtArgs& a, const OtPlan& p, cudaStream_t st) {
  switch (p.MP) {
    case 16: return dispatch_cfg<DT, 16>(grad, a, p, st);
    case 32: return dispatch_cfg<DT, 32>(grad, a, p, st);
    default: return dispatch_cfg<DT, 64>(grad, a, p, st);
  }
}

template <int MP>
int launch_ipot_mp(const IpotArgs& a, int N, cudaStream_t st, bool* handled) {
  *handled = true;
  // one warp per sample while the plan fits ~40 registers per lane, else 256 or 512 threads
  constexpr int TC = MP / 4;
  const int rptw = (N + 32 / TC - 1) / (32 / TC);
  const int rpt256 = (N + 256 / TC - 1) / (256 / TC);
  const int rpt512 = (N + 512 / TC - 1) / (512 / TC);
#define CE_IPOT_CASE(COND, R, G, ...) \
  if (COND <= R) { ot_ipot_kernel<MP, R, G, ##__VA_ARGS__><<<(G == 32 ? (a.B + 7) / 8 : a.B), (G < 256 ? 256 : G), 0, st>>>(a); return CE_OK; }
  CE_IPOT_CASE(rptw, 2, 32, 2) CE_IPOT_CASE(rptw, 4, 32, 2) CE_IPOT_CASE(rptw, 7, 32, 2) CE_IPOT_CASE(rptw, 10, 32)
  {   // CE_OT_RAGGED=0 (tuning aid): every sample runs the body sized for N
    static const bool ragged_on = [] { const char* e = getenv("CE_OT_RAGGED"); return e == nullptr || atoi(e) != 0; }();
    if (ragged_on && rpt256 >= 2 && rpt256 <= 9) { ot_ipot_ragged_kernel<MP><<<a.B, 256, 0, st>>>(a); return CE_OK; }
  }
  CE_IPOT_CASE(rpt256, 1, 256) CE_IPOT_CASE(rpt256, 2, 256) CE_IPOT_CASE(rpt256, 4, 256)
  // (measured at c4 and dropped: 512 threads per sample, and the two-barrier SCHEME 1 at either size)
  CE_IPOT_CASE(rpt256, 7, 256, 2, 0) CE_IPOT_CASE(rpt256, 9, 256, 2, 0) CE_IPOT_CASE(rpt256, 13, 256)
  CE_IPOT_CASE(rpt512, 10, 512) CE_IPOT_CASE(rpt512, 13, 512)
  if (rpt512 <= 20) {   // kernel matrix in shared memory (64 x 577 and the like)
    const size_t sm = sizeof(float) * (size_t)20 * (512 / TC) * MP;
    auto kern = ot_ipot_kernel<MP, 20, 512, 1, 0, true>;
    CE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    kern<<<a.B, 512, sm, st>>>(a);
    return CE_OK;
  }
#undef CE_IPOT_CASE
  *handled = false;
  return CE_OK;
}

}  // namespace
}  // namespace ce

using namespace ce;

extern "C" size_t ce_ot_workspace_bytes(int B, int M, int N, int D) {
  OtPlan p;
  if (make_plan(M, N, &p) != CE_OK) return 0;
  int Nld = (N + 3) / 4 * 4;
  size_t bytes = 0;
  bytes += align_up(sizeof(float) * (size_t)B * N * p.MP, 256);       // S / W
  bytes += align_up(sizeof(float) * (size_t)B * p.MP, 256);           // nx2 / ax
  bytes += align_up(sizeof(float) * (size_t)B * Nld, 256);            // ny2 / ay
  bytes += align_up(sizeof(float) * (size_t)B * 2 * N * p.MP, 256);   // big-shape solver scratch
  if (p.nsplit > 1 || ot_wide_nsplit(M, N, D) > 1) bytes += align_up(sizeof(float) * (size_t)B * M * D, 256);  // dx accumulators
  return bytes + 1024;
}

extern "C" int ce_ot_fwd_bwd(const void* txt, int64_t txt_bstride, const void* img,
                             int64_t img_bstride, const void* txt_mask, int64_t txt_mstride,
                             const void* img_mask, int64_t img_mstride, int mask_kind, int B, int M,
                             int N, int D, int dtype, float beta, int iters, int k,
                             float loss_scale, float* dist, float* loss, void* dtxt, void* dimg,
                             void* dimg_slot0, void* workspace, size_t workspace_bytes,
                             ce_stream_t stream) {
  CE_TRY(check_device());
  if (dtype != CE_F32 && dtype != CE_BF16) return fail(CE_ERR_DTYPE, "OT: unknown dtype %d", dtype);
  if (mask_kind != CE_MASK_NUM_I64 && mask_kind != CE_MASK_PAD_U8)
    return fail(CE_ERR_ARG, "OT: unknown mask_kind %d", mask_kind);
  if (B < 0 || D < 8 || D % 8 != 0) return fail(CE_ERR_SHAPE, "OT: need B >= 0 and D a multiple of 8 (B=%d D=%d)", B, D);
  if (iters < 0 || k < 1) return fail(CE_ERR_ARG, "OT: need iters >= 0 and k >= 1");
  if (!(beta > 0.f)) return fail(CE_ERR_ARG, "OT: beta must be positive");
  if ((dtxt == nullptr) != (dimg == nullptr)) return fail(CE_ERR_ARG, "OT: pass both gradient buffers or neither");
  const size_t esz = dtype == CE_F32 ? 4 : 2;
  if (((uintptr_t)txt | (uintptr_t)img | (uintptr_t)dtxt | (uintptr_t)dimg) & 15)
    return fail(CE_ERR_ALIGN, "OT: embedding pointers must be 16-byte aligned");
  if ((txt_bstride * esz) % 16 || (img_bstride * esz) % 16)
    return fail(CE_ERR_ALIGN, "OT: sample strides must keep 16-byte alignment");
  OtPlan p;
  CE_TRY(make_plan(M, N, &p));
  if (workspace_bytes < ce_ot_workspace_bytes(B, M, N, D))
    return fail(CE_ERR_WORKSPACE, "OT: workspace too small (%zu < %zu)", workspace_bytes, ce_ot_workspace_bytes(B, M, N, D));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (B == 0) {
    if (loss) CE_CUDA_TRY(cudaMemsetAsync(loss, 0, sizeof(float), st));
    return CE_OK;
  }
  // Small plans in bf16: ONE persistent kernel keeps x, y resident in shared memory from the load to
  // the gradient store (csrc/ot_fused.cu).  CE_OT_FUSED=0 (debug switch) forces the three-kernel path.
  static const bool fused_ok = [] { const char* e = getenv("CE_OT_FUSED"); return e == nullptr || atoi(e) != 0; }();
  // CE_OT_STREAM=0 selects the shared-memory-resident predecessor (csrc/ot_fused.cu) where both apply.
  static const bool stream_ok = [] { const char* e = getenv("CE_OT_STREAM"); return e == nullptr || atoi(e) != 0; }();
  const bool use_stream = fused_ok && stream_ok && ot_stream_supported(M, N, D, dtype);
  if (use_stream || (fused_ok && ot_fused_supported(M, N, D, dtype))) {
    OtFusedArgs fa{};
    fa.txt = txt; fa.img = img; fa.txt_bs = txt_bstride; fa.img_bs = img_bstride;
    fa.txt_mask = txt_mask; fa.img_mask = img_mask; fa.txt_ms = txt_mstride; fa.img_ms = img_mstride;
    fa.mask_kind = mask_kind; fa.B = B; fa.M = M; fa.N = N; fa.D = D;
    fa.beta = beta; fa.eps = 1e-5f; fa.scale = loss_scale; fa.iters = iters; fa.k = k;
    fa.dist = dist; fa.dtxt = dtxt; fa.dimg = dimg; fa.dslot0 = dtxt != nullptr ? dimg_slot0 : nullptr;
    if (use_stream) CE_TRY(launch_ot_stream(fa, st));
    else CE_TRY(launch_ot_fused(fa, st));
    ot_tail_kernel<CE_BF16><<<1, 256, 0, st>>>(dist, B, loss_scale, loss, nullptr, img_bstride, D);
    CE_LAUNCH_CHECK();
    return CE_OK;
  }
  const int Nld = (N + 3) / 4 * 4;
  Carver cv(workspace);
  OtArgs a{};
  a.txt = txt; a.img = img; a.txt_bs = txt_bstride; a.img_bs = img_bstride;
  a.B = B; a.M = M; a.N = N; a.D = D; a.tiles_per_cta = p.tiles_per_cta;
  a.S = cv.take<float>((size_t)B * N * p.MP);
  a.nx2 = cv.take<float>((size_t)B * p.MP);
  a.ny2 = cv.take<float>((size_t)B * Nld);
  float* scratch = cv.take<float>((size_t)B * 2 * N * p.MP);
  a.Nld = Nld; a.dtxt = dtxt; a.dimg = dimg; a.nsplit = p.nsplit;
  // bf16 gradients of these plans: the TMA-fed kernel of csrc/ot_wide.cu (CE_OT_WIDE=0: the mma.sync predecessor)
  const bool wide_grad = dtxt != nullptr && ot_wide_supported(M, N, D, dtype);
  const int grad_nsplit = wide_grad ? ot_wide_nsplit(M, N, D) : p.nsplit;
  a.dx_acc = (p.nsplit > 1 || grad_nsplit > 1) ? cv.take<float>((size_t)B * M * D) : nullptr;

  if (ot_wide_cost_supported(M, N, D, dtype)) {   // bf16: TMA-fed persistent kernel (csrc/ot_wide.cu)
    OtWideCostArgs wc{};
    wc.txt = txt; wc.img = img; wc.txt_bs = txt_bstride; wc.img_bs = img_bstride;
    wc.B = B; wc.M = M; wc.N = N; wc.D = D; wc.S = a.S; wc.nx2 = a.nx2; wc.ny2 = a.ny2; wc.Nld = Nld;
    CE_TRY(launch_ot_wide_cost(wc, st));
  } else if (dtype == CE_F32) CE_TRY(dispatch_mp<CE_F32>(false, a, p, st));
  else CE_TRY(dispatch_mp<CE_BF16>(false, a, p, st));

  IpotArgs ia{};
  ia.S = a.S; ia.nx2 = a.nx2; ia.ny2 = a.ny2; ia.txt_mask = txt_mask; ia.img_mask = img_mask;
  ia.txt_ms = txt_mstride; ia.img_ms = img_mstride; ia.mask_kind = mask_kind;
  ia.B = B; ia.M = M; ia.N = N; ia.Nld = Nld; ia.beta = beta; ia.eps = 1e-5f; ia.scale = loss_scale;
  ia.iters = iters; ia.k = k; ia.dist = dist;
  bool handled = false;
  switch (p.MP) {
    case 16: CE_TRY(launch_ipot_mp<16>(ia, N, st, &handled)); break;
    case 32: CE_TRY(launch_ipot_mp<32>(ia, N, st, &handled)); break;
    default: CE_TRY(launch_ipot_mp<64>(ia, N, st, &handled)); break;
  }
  if (!handled) {
    IpotBigArgs ba{ia, scratch, p.MP};
    size_t sm = sizeof(float) * (4 * (size_t)p.MP + 3 * (size_t)N);
    ot_ipot_big_kernel<<<B, 256, sm, st>>>(ba);
  }
  CE_LAUNCH_CHECK();

  if (dtxt != nullptr) {
    if (grad_nsplit > 1) CE_CUDA_TRY(cudaMemsetAsync(a.dx_acc, 0, sizeof(float) * (size_t)B * M * D, st));
    if (wide_grad) {
      OtWideGradArgs wg{};
      wg.txt = txt; wg.img = img; wg.txt_bs = txt_bstride; wg.img_bs = img_bstride;
      wg.B = B; wg.M = M; wg.N = N; wg.D = D; wg.W = a.S; wg.ax = a.nx2; wg.ay = a.ny2; wg.Nld = Nld;
      wg.dtxt = dtxt; wg.dimg = dimg; wg.dx_acc = a.dx_acc;
      CE_TRY(launch_ot_wide_grad(wg, st));
    } else if (dtype == CE_F32) CE_TRY(dispatch_mp<CE_F32>(true, a, p, st));
    else CE_TRY(dispatch_mp<CE_BF16>(true, a, p, st));
    if (grad_nsplit > 1) {
      int blocks = (int)std::min<int64_t>(((int64_t)B * M * D + 255) / 256, 148 * 8);
      if (dtype == CE_F32) ot_dx_finish_kernel<CE_F32><<<blocks, 256, 0, st>>>(a, p.MP);
      else ot_dx_finish_kernel<CE_BF16><<<blocks, 256, 0, st>>>(a, p.MP);
      CE_LAUNCH_CHECK();
    }
  }
  {
    void* slot = dtxt != nullptr ? dimg_slot0 : nullptr;
    int blocks = slot ? (int)std::min<int64_t>(((int64_t)B * D * (int64_t)esz / 16 + 255) / 256, 148 * 4) : 1;
    if (dtype == CE_F32) ot_tail_kernel<CE_F32><<<blocks, 256, 0, st>>>(dist, B, loss_scale, loss, slot, img_bstride, D);
    else ot_tail_kernel<CE_BF16><<<blocks, 256, 0, st>>>(dist, B, loss_scale, loss, slot, img_bstride, D);
    CE_LAUNCH_CHECK();
  }
  return CE_OK;
}

// Packed (variable-length) node sets: SURVEY.md 8f-3 -- the padded slots of model_clip.py:531-552 /
// dataset_voa.py:532-544,566-577 never reach the kernel.  Served by the streaming kernel only.
extern "C" int ce_ot_fwd_bwd_packed(const void* txt_rows, const int32_t* txt_off, const void* img_rows,
                                    const int32_t* img_off, int B, int max_m, int max_n, int D, int dtype,
                                    float beta, int iters, int k, float loss_scale, float* dist, float* loss,
                                    void* dtxt_rows, void* dimg_rows, ce_stream_t stream) {
  CE_TRY(check_device());
  if (dtype != CE_BF16) return fail(CE_ERR_DTYPE, "packed OT: bf16 only (pad the node sets for fp32)");
  if (B < 0 || iters < 0 || k < 1 || !(beta > 0.f)) return fail(CE_ERR_ARG, "packed OT: bad B / iters / k / beta");
  if (!ot_stream_supported(max_m, max_n, D, dtype))
    return fail(CE_ERR_SHAPE, "packed OT: needs max_m <= 16, max_n <= 64 and D a multiple of 64 up to 512 (max_m=%d max_n=%d D=%d); pad the node sets instead", max_m, max_n, D);
  if ((dtxt_rows == nullptr) != (dimg_rows == nullptr)) return fail(CE_ERR_ARG, "packed OT: pass both gradient buffers or neither");
  if (((uintptr_t)txt_rows | (uintptr_t)img_rows | (uintptr_t)dtxt_rows | (uintptr_t)dimg_rows) & 15)
    return fail(CE_ERR_ALIGN, "packed OT: row matrices must be 16-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (B == 0) {
    if (loss) CE_CUDA_TRY(cudaMemsetAsync(loss, 0, sizeof(float), st));
    return CE_OK;
  }
  OtFusedArgs fa{};
  fa.txt = txt_rows; fa.img = img_rows; fa.txt_off = txt_off; fa.img_off = img_off;
  fa.B = B; fa.M = max_m; fa.N = max_n; fa.D = D;
  fa.beta = beta; fa.eps = 1e-5f; fa.scale = loss_scale; fa.iters = iters; fa.k = k;
  fa.dist = dist; fa.dtxt = dtxt_rows; fa.dimg = dimg_rows; fa.dslot0 = nullptr;
  CE_TRY(launch_ot_stream(fa, st));
  ot_tail_kernel<CE_BF16><<<1, 256, 0, st>>>(dist, B, loss_scale, loss, nullptr, 0, D);
  CE_LAUNCH_CHECK();
  return CE_OK;
}

extern "C" int ce_ot_cost_matrix(const void* x, const void* y, int B, int M, int N, int D,
                                 int dtype, float eps, float* cost, ce_stream_t stream) {
  CE_TRY(check_device());
  if (dtype != CE_F32 && dtype != CE_BF16) return fail(CE_ERR_DTYPE, "cost_matrix: unknown dtype %d", dtype);
  if (B < 0 || M < 0 || N < 0 || D < 1) return fail(CE_ERR_SHAPE, "cost_matrix: bad shape");
  int64_t warps = (int64_t)B * M * N;
  if (warps == 0) return CE_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int64_t blocks = (warps * 32 + 255) / 256;
  if (blocks > 0x7fffffff) return fail(CE_ERR_SHAPE, "cost_matrix: too large");
  if (dtype == CE_F32) ot_cost_matrix_kernel<CE_F32><<<(int)blocks, 256, 0, st>>>(x, y, B, M, N, D, eps, cost);
  else ot_cost_matrix_kernel<CE_BF16><<<(int)blocks, 256, 0, st>>>(x, y, B, M, N, D, eps, cost);
  CE_LAUNCH_CHECK();
  return CE_OK;
}

extern "C" int ce_ot_ipot(const float* cost, const uint8_t* x_pad, const uint8_t* y_pad, int B, int M,
                          int N, float beta, int iters, int k, float* plan, ce_stream_t stream) {
  CE_TRY(check_device());
  if (B < 0 || M < 1 || N < 1) return fail(CE_ERR_SHAPE, "ipot: bad shape");
  if (iters < 0 || k < 1 || !(beta > 0.f)) return fail(CE_ERR_ARG, "ipot: need iters >= 0, k >= 1, beta > 0");
  if (B == 0) return CE_OK;
  size_t sm = sizeof(float) * 2 * ((size_t)M + N);
  if (sm > 200 * 1024) return fail(CE_ERR_SHAPE, "ipot: M+N too large");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  CE_CUDA_TRY(cudaFuncSetAttribute(ot_ipot_plain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  ot_ipot_plain_kernel<<<B, 256, sm, st>>>(cost, x_pad, y_pad, B, M, N, beta, iters, k, plan);
  CE_LAUNCH_CHECK();
  return CE_OK;
}

extern "C" int ce_ot_trace(const float* x, int B, int n, float* out, ce_stream_t stream) {
  CE_TRY(check_device());
  if (B < 0 || n < 0) return fail(CE_ERR_SHAPE, "trace: bad shape");
  if (B == 0) return CE_OK;
  ot_trace_kernel<<<(B + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, B, n, out);
  CE_LAUNCH_CHECK();
  return CE_OK;
}

extern "C" int ce_scale_inplace(void* x, int64_t rows, int64_t row_len, int64_t row_stride,
                                int dtype, const float* g, ce_stream_t stream) {
  CE_TRY(check_device());
  if (dtype != CE_F32 && dtype != CE_BF16) return fail(CE_ERR_DTYPE, "scale: unknown dtype %d", dtype);
  int64_t total = rows * row_len;
  if (total <= 0) return CE_OK;
  int blocks = (int)std::min<int64_t>((total + 255) / 256, 148 * 8);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == CE_F32) scale_inplace_kernel<CE_F32><<<blocks, 256, 0, st>>>(x, rows, row_len, row_stride, g, nullptr);
  else scale_inplace_kernel<CE_BF16><<<blocks, 256, 0, st>>>(x, rows, row_len, row_stride, g, nullptr);
  CE_LAUNCH_CHECK();
  return CE_OK;
}

extern "C" int ce_scale_inplace_same(void* x, int64_t n, int dtype, const float* g, const float* g_same,
                                     ce_stream_t stream) {
  CE_TRY(check_device());
  if (dtype != CE_F32 && dtype != CE_BF16) return fail(CE_ERR_DTYPE, "scale: unknown dtype %d", dtype);
  if (n <= 0) return CE_OK;
  int blocks = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == CE_F32) scale_inplace_kernel<CE_F32><<<blocks, 256, 0, st>>>(x, 1, n, n, g, g_same);
  else scale_inplace_kernel<CE_BF16><<<blocks, 256, 0, st>>>(x, 1, n, n, g, g_same);
  CE_LAUNCH_CHECK();
  return CE_OK;
}
