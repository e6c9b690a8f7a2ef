// Alternative IPOT iteration layout: lane = (row group r = lane >> 2, column group q = lane & 3) holds rows
// r + 8 j (j = 0..7) x columns 4 q .. 4 q + 3 of the 64 x 16 plan; both reductions are transposing shuffle
// butterflies (19 shuffles per iteration, no shared memory).  Register order is permuted per lane so that
// "keep the low half, send the high half" needs no selects.  Compare with ipot_loop.cu (same arithmetic).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float frcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
#define SHX(v, m) __shfl_xor_sync(0xffffffffu, (v), (m))

__global__ void k(long long* out, float* sink, int iters) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float2 A[8][2], R[8][2];
  for (int i = 0; i < 8; ++i)
    for (int h = 0; h < 2; ++h) {
      A[i][h] = make_float2(0.9f - 0.01f * i - 0.001f * lane, 0.8f + 0.01f * i + 0.02f * h);
      R[i][h] = make_float2(1.f, 1.f);
    }
  const float xlen = 16.f, ylen = 64.f;
  float u0 = 1.f, u1 = 1.f;        // plan row factors of the two rows this lane owns
  float v_c = 1.f;                 // plan column factor of the column this lane owns
  float w[4] = {1.f / 16, 1.f / 16, 1.f / 16, 1.f / 16};   // v * sigma of this lane's 4 columns (register order)
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { R[i][0] = __fmul2_rn(R[i][0], A[i][0]); R[i][1] = __fmul2_rn(R[i][1], A[i][1]); }
    // ---- row sums: 4-term dots in-thread, then a transposing reduce over the quad (lane bits 0, 1)
    const float2 w01 = make_float2(w[0], w[1]), w23 = make_float2(w[2], w[3]);
    float s[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { const float2 p = __ffma2_rn(R[i][1], w23, __fmul2_rn(R[i][0], w01)); s[i] = p.x + p.y; }
#pragma unroll
    for (int i = 0; i < 4; ++i) s[i] += SHX(s[i + 4], 1);
    s[0] += SHX(s[2], 2); s[1] += SHX(s[3], 2);
    const float d0 = frcp(ylen * (u0 * s[0])), d1 = frcp(ylen * (u1 * s[1]));
    float z[8];
    z[0] = d0 * u0; z[1] = d1 * u1;
    z[2] = SHX(z[0], 2); z[3] = SHX(z[1], 2);
#pragma unroll
    for (int i = 0; i < 4; ++i) z[i + 4] = SHX(z[i], 1);
    // ---- column sums: 8-term dots in-thread, transposing reduce over lane bits 2, 3, 4
    float2 c01 = __fmul2_rn(make_float2(z[0], z[0]), R[0][0]), c23 = __fmul2_rn(make_float2(z[0], z[0]), R[0][1]);
    float2 e01 = __fmul2_rn(make_float2(z[1], z[1]), R[1][0]), e23 = __fmul2_rn(make_float2(z[1], z[1]), R[1][1]);
#pragma unroll
    for (int i = 2; i < 8; i += 2) {
      c01 = __ffma2_rn(make_float2(z[i], z[i]), R[i][0], c01); c23 = __ffma2_rn(make_float2(z[i], z[i]), R[i][1], c23);
      e01 = __ffma2_rn(make_float2(z[i + 1], z[i + 1]), R[i + 1][0], e01); e23 = __ffma2_rn(make_float2(z[i + 1], z[i + 1]), R[i + 1][1], e23);
    }
    float c0 = c01.x + e01.x, c1 = c01.y + e01.y, c2 = c23.x + e23.x, c3 = c23.y + e23.y;
    c0 += SHX(c2, 4); c1 += SHX(c3, 4);
    c0 += SHX(c1, 8);
    c0 += SHX(c0, 16);
    const float sig = frcp(xlen * (v_c * c0));
    u0 = z[0]; u1 = z[1];
    v_c *= sig;
    if ((it & 7) == 7) {   // refold: u of all 8 rows = z (already gathered), v of the 4 columns by an all-gather
      float vv[4];
      vv[0] = v_c; vv[1] = SHX(vv[0], 8); vv[2] = SHX(vv[0], 4); vv[3] = SHX(vv[1], 4);
      const float2 v01 = make_float2(vv[0], vv[1]), v23 = make_float2(vv[2], vv[3]);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float2 zz = make_float2(z[i], z[i]);
        R[i][0] = __fmul2_rn(__fmul2_rn(R[i][0], zz), v01); R[i][1] = __fmul2_rn(__fmul2_rn(R[i][1], zz), v23);
      }
      u0 = u1 = 1.f; v_c = 1.f;
    }
    w[0] = v_c * sig;
    w[1] = SHX(w[0], 8);
    w[2] = SHX(w[0], 4); w[3] = SHX(w[1], 4);
  }
  long long t1 = clock64();
  if (lane == 0) out[blockIdx.x * 8 + wid] = t1 - t0;
  float acc = u0 + u1 + v_c;
  for (int i = 0; i < 8; ++i) acc += R[i][0].x + R[i][1].y;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main() {
  long long* out; float* sink;
  cudaMalloc(&out, 148 * 8 * 8); cudaMalloc(&sink, 148 * 256 * 4);
  for (int nw : {1, 3, 4, 8}) {
    for (int grid : {1, 148}) {
      k<<<grid, 32 * nw>>>(out, sink, 50);
      k<<<grid, 32 * nw>>>(out, sink, 50);
      cudaDeviceSynchronize();
      long long h[8];
      cudaMemcpy(h, out, 64, cudaMemcpyDeviceToHost);
      printf("quad layout: warps/CTA %d grid %3d: %.0f cycles / iteration (warp 0)\n", nw, grid, (double)h[0] / 50);
    }
  }
  return 0;
}
