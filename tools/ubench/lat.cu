// Latency micro-benchmarks for the primitives of the IPOT solver loop (one warp, dependent chains).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/lat tools/ubench/lat.cu && /tmp/lat
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

#define N 256
__device__ __forceinline__ float frcp(float x) { float y; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

__global__ void k(long long* out, float* sink, int mode) {
  __shared__ float sm[32 * 20 + 64];
  const int lane = threadIdx.x;
  float x = 1.0f + lane * 1e-3f, y = 0.999f;
  float2 a2 = make_float2(x, x), b2 = make_float2(y, y);
  sm[lane] = x;
  __syncwarp();
  long long t0 = clock64();
  if (mode == 0) { for (int i = 0; i < N; ++i) x = fmaf(x, y, 0.5f); }
  else if (mode == 1) { for (int i = 0; i < N; ++i) a2 = __ffma2_rn(a2, b2, b2); x = a2.x + a2.y; }
  else if (mode == 2) { for (int i = 0; i < N; ++i) x = frcp(x) + 0.5f; }
  else if (mode == 3) { for (int i = 0; i < N; ++i) x += __shfl_xor_sync(0xffffffffu, x, 16); }
  else if (mode == 4) {   // STS -> syncwarp -> LDS (other lane) round trip
    for (int i = 0; i < N; ++i) { sm[lane] = x; __syncwarp(); x = sm[lane ^ 1] + 1.f; __syncwarp(); }
  } else if (mode == 5) {  // LDS dependent chain (pointer chase through values)
    int idx = lane;
    sm[lane] = __int_as_float((lane + 1) & 31);
    __syncwarp();
    for (int i = 0; i < N; ++i) idx = __float_as_int(sm[idx]);
    x = (float)idx;
  } else if (mode == 6) {  // syncwarp alone
    for (int i = 0; i < N; ++i) { __syncwarp(); x = fmaf(x, y, 0.5f); }
  } else if (mode == 7) {  // STS.64 x8 -> syncwarp -> 16 LDS.32 (transpose pattern) -> adds
    float* prow = sm + lane * 18;
    const float* pcol = sm + (lane >> 4) * 16 * 18 + (lane & 15);
    for (int i = 0; i < N; ++i) {
#pragma unroll
      for (int q = 0; q < 8; ++q) *reinterpret_cast<float2*>(prow + 2 * q) = make_float2(x + q, x - q);
      __syncwarp();
      float e0 = 0, e1 = 0, e2 = 0, e3 = 0;
#pragma unroll
      for (int j = 0; j < 16; j += 4) { e0 += pcol[j * 18]; e1 += pcol[(j + 1) * 18]; e2 += pcol[(j + 2) * 18]; e3 += pcol[(j + 3) * 18]; }
      x = (e0 + e1) + (e2 + e3);
      x = x * 1e-3f;
      __syncwarp();
    }
  } else if (mode == 8) {  // predicated STS (16 lanes) -> syncwarp -> LDS.128 x4 broadcast
    for (int i = 0; i < N; ++i) {
      if (lane < 16) sm[lane] = x;
      __syncwarp();
      const float4* wp = reinterpret_cast<const float4*>(sm);
      float4 a = wp[0], b = wp[1], c = wp[2], d = wp[3];
      x = (a.x + b.y) + (c.z + d.w);
      x *= 0.25f;
      __syncwarp();
    }
  } else if (mode == 9) {  // clock64 back to back
    for (int i = 0; i < N; ++i) { x += (float)(clock64() & 1); }
  }
  long long t1 = clock64();
  if (lane == 0) out[mode] = t1 - t0;
  sink[lane + 32 * mode] = x;
}

int main() {
  long long* out; float* sink;
  cudaMalloc(&out, 16 * 8); cudaMalloc(&sink, 32 * 16 * 4);
  const char* names[] = {"FFMA dep", "FFMA2 dep", "MUFU.RCP+FADD dep", "SHFL+FADD dep", "STS->syncwarp->LDS->syncwarp", "LDS dep",
                         "syncwarp+FFMA", "transpose 8xSTS.64->16xLDS->adds", "w broadcast STS->4xLDS.128", "clock64"};
  for (int rep = 0; rep < 2; ++rep)
    for (int m = 0; m < 10; ++m) { k<<<1, 32>>>(out, sink, m); }
  cudaDeviceSynchronize();
  long long h[16];
  cudaMemcpy(h, out, 16 * 8, cudaMemcpyDeviceToHost);
  for (int m = 0; m < 10; ++m) printf("%-40s %7.1f cycles/iter\n", names[m], (double)h[m] / N);
  return 0;
}
