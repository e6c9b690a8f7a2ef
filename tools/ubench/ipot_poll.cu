// Does a set of idle warps waiting on an mbarrier slow down latency-bound solver warps on the same SM?
// 12 warps: warps 1..3 run the IPOT iteration loop, the other nine wait on a barrier in `mode`:
//   0 nothing (exit)   1 try_wait spin, all lanes   2 lane 0 try_wait + nanosleep(64)   3 lane 0 nanosleep(1000) + flag poll
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>
#ifndef BIG
#define BIG 0
#endif
constexpr int kPLd = 18, kMP = 16;
__device__ __forceinline__ float frcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
struct Scr { float P[32 * kPLd]; float w[16]; float v[16]; };
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__global__ void __launch_bounds__(384, 1) k(long long* out, float* sink, int iters, int mode) {
  extern __shared__ __align__(128) unsigned char dyn[];
  // same placement as in ot_fused_kernel: scratch blocks behind 208 000 bytes of sample slots, 7872 bytes apart
  auto scr_at = [&](int i) -> Scr& { return *reinterpret_cast<Scr*>(dyn + (BIG ? 208000 + 5440 : 0) + (size_t)i * (BIG ? 7872 : sizeof(Scr))); };
  __shared__ uint64_t bar;
  __shared__ volatile int flag;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(3)); flag = 0; }
  __syncthreads();
  if (wid == 0 || wid >= 4) {
    if (mode == 0) return;
    if (mode == 1) { while (!try_wait(&bar, 0)) { } }
    else if (mode == 2) { if (lane == 0) { while (!try_wait(&bar, 0)) __nanosleep(64); } __syncwarp(); }
    else if (mode == 3) { if (lane == 0) { while (flag < 3) __nanosleep(1000); } __syncwarp(); }
    else if (mode == 4) { if (lane == 0) { while (!try_wait(&bar, 0)) __nanosleep(2000); } __syncwarp(); }
    return;
  }
  Scr& sc = scr_at(wid - 1);
  const int c = lane & 15;
  const float* pcol = sc.P + (lane >> 4) * 16 * kPLd + c;
  const int prot = (lane >> 4) * 8;
  float2 A0[8], A1[8], R0[8], R1[8];
  for (int q = 0; q < 8; ++q) {
    A0[q] = make_float2(0.9f - 0.01f * q - 0.001f * lane, 0.8f + 0.01f * q); A1[q] = make_float2(0.7f + 0.002f * lane, 0.95f - 0.01f * q);
    R0[q] = make_float2(1.f, 1.f); R1[q] = make_float2(1.f, 1.f);
  }
  const float xlen = 16.f, ylen = 64.f;
  float u0 = 1.f, u1 = 1.f, v_c = 1.f, sig_c = 1.f / xlen;
  if (lane < kMP) sc.w[lane] = v_c * sig_c;
  __syncwarp();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int q = 0; q < 8; ++q) { R0[q] = __fmul2_rn(R0[q], A0[q]); R1[q] = __fmul2_rn(R1[q], A1[q]); }
    float2 w2[8];
    {
      const float4* wp = reinterpret_cast<const float4*>(sc.w);
#pragma unroll
      for (int q = 0; q < 4; ++q) { const float4 w4 = wp[q]; w2[2 * q] = make_float2(w4.x, w4.y); w2[2 * q + 1] = make_float2(w4.z, w4.w); }
    }
    float2 pa = __fmul2_rn(R0[0], w2[0]), pb = __fmul2_rn(R0[1], w2[1]);
    float2 qa = __fmul2_rn(R1[0], w2[0]), qb = __fmul2_rn(R1[1], w2[1]);
#pragma unroll
    for (int q = 2; q < 8; q += 2) {
      pa = __ffma2_rn(R0[q], w2[q], pa); pb = __ffma2_rn(R0[q + 1], w2[q + 1], pb);
      qa = __ffma2_rn(R1[q], w2[q], qa); qb = __ffma2_rn(R1[q + 1], w2[q + 1], qb);
    }
    const float rs0 = (pa.x + pa.y) + (pb.x + pb.y), rs1 = (qa.x + qa.y) + (qb.x + qb.y);
    const float d0 = frcp(ylen * (u0 * rs0)), d1 = frcp(ylen * (u1 * rs1));
    const float z0 = d0 * u0, z1 = d1 * u1;
    const float2 zz0 = make_float2(z0, z0), zz1 = make_float2(z1, z1);
    float2* prow = reinterpret_cast<float2*>(sc.P + lane * kPLd);
#pragma unroll
    for (int q = 0; q < 8; ++q) prow[q] = __ffma2_rn(zz1, R1[q], __fmul2_rn(zz0, R0[q]));
    __syncwarp();
    float e0 = 0.f, e1 = 0.f, e2 = 0.f, e3 = 0.f;
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
      e0 += pcol[((i + 0 + prot) & 15) * kPLd]; e1 += pcol[((i + 1 + prot) & 15) * kPLd];
      e2 += pcol[((i + 2 + prot) & 15) * kPLd]; e3 += pcol[((i + 3 + prot) & 15) * kPLd];
    }
    float cs = (e0 + e1) + (e2 + e3);
    cs += __shfl_xor_sync(0xffffffffu, cs, 16);
    sig_c = frcp(xlen * (v_c * cs));
    u0 = z0; u1 = z1;
    v_c *= sig_c;
    if ((it & 3) == 3) {
      if (lane < kMP) sc.v[lane] = v_c;
      __syncwarp();
      const float4* vp = reinterpret_cast<const float4*>(sc.v);
      const float2 uu0 = make_float2(u0, u0), uu1 = make_float2(u1, u1);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 v4 = vp[q];
        const float2 va = make_float2(v4.x, v4.y), vb = make_float2(v4.z, v4.w);
        R0[2 * q] = __fmul2_rn(__fmul2_rn(R0[2 * q], uu0), va); R0[2 * q + 1] = __fmul2_rn(__fmul2_rn(R0[2 * q + 1], uu0), vb);
        R1[2 * q] = __fmul2_rn(__fmul2_rn(R1[2 * q], uu1), va); R1[2 * q + 1] = __fmul2_rn(__fmul2_rn(R1[2 * q + 1], uu1), vb);
      }
      u0 = u1 = 1.f; v_c = 1.f;
    }
    if (lane < kMP) sc.w[lane] = v_c * sig_c;
    __syncwarp();
  }
  long long t1 = clock64();
  if (lane == 0) out[blockIdx.x * 4 + wid] = t1 - t0;
  float s = u0 + u1 + v_c;
  for (int q = 0; q < 8; ++q) s += R0[q].x + R1[q].y;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
  __syncwarp();
  if (lane == 0) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar)) : "memory"); atomicAdd((int*)&flag, 1); }
}
#ifndef BIG
#define BIG 0
#endif
int main() {
  size_t dynb = BIG ? 231616 : 3 * sizeof(Scr);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dynb);
  long long* out; float* sink;
  cudaMalloc(&out, 148 * 4 * 8); cudaMalloc(&sink, 148 * 384 * 4);
  for (int mode = 0; mode < 5; ++mode) {
    k<<<148, 384, dynb>>>(out, sink, 50, mode);
    k<<<148, 384, dynb>>>(out, sink, 50, mode);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[4];
    cudaMemcpy(h, out, 32, cudaMemcpyDeviceToHost);
    printf("mode %d: %.0f %.0f %.0f cycles / iteration (%s)\n", mode, h[1] / 50.0, h[2] / 50.0, h[3] / 50.0, cudaGetErrorString(e));
  }
  return 0;
}
