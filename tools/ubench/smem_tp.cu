// Shared-memory pipe throughput of the instructions the OT kernels use (cycles per warp instruction, per SM):
// nw warps issue independent, conflict-free accesses back to back.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define N 1024
#define U 8
__global__ void k(long long* out, float* sink, int mode) {
  extern __shared__ __align__(128) uint8_t sm[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int RS = 1040;
  for (int i = threadIdx.x; i < 16 * RS / 4 * 8; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = i;
  __syncthreads();
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm) + (w & 7) * 16 * RS + (lane & 15) * RS + (lane >> 4) * 16;
  const uint32_t lin = (uint32_t)__cvta_generic_to_shared(sm) + (w & 7) * 16 * RS + lane * 16;
  uint32_t acc = 0;
  uint32_t r[U][4];
  long long t0 = clock64();
  for (int i = 0; i < N; i += U) {
    if (mode == 0) {
#pragma unroll
      for (int u = 0; u < U; ++u) asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[u][0]), "=r"(r[u][1]), "=r"(r[u][2]), "=r"(r[u][3]) : "r"(base + u * 32));
    } else if (mode == 1) {
#pragma unroll
      for (int u = 0; u < U; ++u) asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[u][0]), "=r"(r[u][1]), "=r"(r[u][2]), "=r"(r[u][3]) : "r"(base + u * 32));
    } else if (mode == 2) {
#pragma unroll
      for (int u = 0; u < U; ++u) asm volatile("stmatrix.sync.aligned.m8n8.x4.shared.b16 [%0], {%1,%2,%3,%4};" ::"r"(base + u * 32), "r"(acc), "r"(acc + 1), "r"(acc + 2), "r"(acc + 3) : "memory");
    } else if (mode == 3) {   // LDS.128, each lane 16 B, linear (conflict-free: 4 wavefronts)
#pragma unroll
      for (int u = 0; u < U; ++u) asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r[u][0]), "=r"(r[u][1]), "=r"(r[u][2]), "=r"(r[u][3]) : "r"(lin + u * 512));
    } else if (mode == 4) {   // STS.128 linear
#pragma unroll
      for (int u = 0; u < U; ++u) asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(lin + u * 512), "r"(acc), "r"(acc + 1), "r"(acc + 2), "r"(acc + 3) : "memory");
    } else if (mode == 5) {   // LDS.32 linear (1 wavefront)
#pragma unroll
      for (int u = 0; u < U; ++u) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r[u][0]) : "r"(lin / 4 + u * 128 + (uint32_t)__cvta_generic_to_shared(sm) * 0 + lane * 0));
    } else if (mode == 6) {   // STS.64, lane stride 72 B (the solver's transpose rows)
#pragma unroll
      for (int u = 0; u < U; ++u) asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"((uint32_t)__cvta_generic_to_shared(sm) + (w & 7) * 4096 + lane * 72 + u * 8), "r"(acc), "r"(acc + 1) : "memory");
    } else if (mode == 7) {   // ldmatrix x4 from a dense 128B-swizzled tile (rows of 128 B)
#pragma unroll
      for (int u = 0; u < U; ++u) asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[u][0]), "=r"(r[u][1]), "=r"(r[u][2]), "=r"(r[u][3]) : "r"((uint32_t)__cvta_generic_to_shared(sm) + (w & 7) * 4096 + (lane & 15) * 128 + ((((lane >> 4) + 2 * (u & 3)) ^ (lane & 7)) << 4)));
    }
    if (mode != 2 && mode != 4 && mode != 6) {
#pragma unroll
      for (int u = 0; u < U; ++u) acc ^= r[u][0];
    } else acc += i;
  }
  long long t1 = clock64();
  if (lane == 0) out[w] = t1 - t0;
  sink[threadIdx.x] = __uint_as_float(acc);
}
int main() {
  long long* out; float* sink;
  cudaMalloc(&out, 64 * 8); cudaMalloc(&sink, 4096 * 4);
  size_t smem = 8 * 16 * 1040 + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const char* names[] = {"LDSM.x4 padded rows", "LDSM.x4.trans padded rows", "STSM.x4 padded rows", "LDS.128 linear", "STS.128 linear", "LDS.32", "STS.64 stride 72B", "LDSM.x4 128B-swizzled dense"};
  for (int nw : {1, 4, 8, 16})
    for (int m = 0; m < 8; ++m) {
      k<<<1, nw * 32, smem>>>(out, sink, m);
      k<<<1, nw * 32, smem>>>(out, sink, m);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[16]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
      printf("warps %2d  %-30s %6.2f cycles per instruction per warp -> %6.2f cycles per instruction on the SM  %s\n", nw, names[m], (double)h[0] / N, (double)h[0] / N / nw, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  return 0;
}
