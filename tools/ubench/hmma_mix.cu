// Does a co-resident FP32-heavy warp (the IPOT solver's profile) slow mma.sync / ldmatrix.trans / stmatrix
// loops on the same SM sub-partition?   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/hmma_mix hmma_mix.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define N 512
__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t* r, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4(uint32_t* r, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void stsm_x4(uint32_t addr, const uint32_t* r) {
  asm volatile("stmatrix.sync.aligned.m8n8.x4.shared.b16 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
// warps 0..3: measured loop (mode); warps 4..4+nbg-1 (same sub-partitions: warp % 4): background FFMA2 / LDS+STS load
__global__ void k(long long* out, float* sink, int mode, int bg) {
  extern __shared__ __align__(128) uint8_t sm[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int RS = 1040;
  for (int i = threadIdx.x; i < 16 * RS / 4 * 8; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;
  __syncthreads();
  if (w >= 4) {
    float2 a2 = make_float2(1.f + lane * 1e-3f, 1.f), b2 = make_float2(0.999f, 0.999f), c2 = a2, d2 = a2;
    float* scr = reinterpret_cast<float*>(sm + 8 * 16 * RS) + (w - 4) * 1024;
    volatile int* flag = reinterpret_cast<volatile int*>(sm + 8 * 16 * RS + 64 * 1024);
    long long it = 0;
    while (*flag < 4 && it < 200000) {
      if (bg == 1) {
#pragma unroll
        for (int i = 0; i < 16; ++i) { a2 = __ffma2_rn(a2, b2, b2); c2 = __ffma2_rn(c2, b2, b2); d2 = __ffma2_rn(d2, b2, b2); }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) { scr[lane * 18 + 2 * i] = a2.x + i; }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 16; ++i) a2.x += scr[i * 18 + (lane & 15)];
        __syncwarp();
      }
      ++it;
    }
    sink[threadIdx.x] = a2.x + a2.y + c2.x + d2.y;
    return;
  }
  uint32_t base = (uint32_t)__cvta_generic_to_shared(sm) + w * 16 * RS + (lane & 15) * RS + (lane >> 4) * 16;
  uint32_t a[4] = {0x3f803f80u, 0x3f803f80u, 0, 0};
  float c[8][4];
  for (int j = 0; j < 8; ++j) for (int q = 0; q < 4; ++q) c[j][q] = 0.f;
  long long t0 = clock64();
  if (mode == 0) { for (int i = 0; i < N; i += 8) { for (int j = 0; j < 8; ++j) mma16816(c[j], a, a[0], a[1]); } }
  else if (mode == 1) {  // LDSM.T -> 2 HMMA, 4 independent accumulators, next load before the MMAs
    uint32_t f0[4], f1[4];
    ldsm_x4_t(f0, base);
    for (int i = 0; i < N; i += 2) {
      ldsm_x4_t(f1, base + ((i + 1) & 15) * 32);
      mma16816(c[0], a, f0[0], f0[1]); mma16816(c[1], a, f0[2], f0[3]);
      ldsm_x4_t(f0, base + ((i + 2) & 15) * 32);
      mma16816(c[2], a, f1[0], f1[1]); mma16816(c[3], a, f1[2], f1[3]);
    }
  } else if (mode == 2) {  // the same with non-transposed loads
    uint32_t f0[4], f1[4];
    ldsm_x4(f0, base);
    for (int i = 0; i < N; i += 2) {
      ldsm_x4(f1, base + ((i + 1) & 15) * 32);
      mma16816(c[0], a, f0[0], f0[1]); mma16816(c[1], a, f0[2], f0[3]);
      ldsm_x4(f0, base + ((i + 2) & 15) * 32);
      mma16816(c[2], a, f1[0], f1[1]); mma16816(c[3], a, f1[2], f1[3]);
    }
  } else if (mode == 3) {  // STSM stream
    uint32_t r[4] = {1, 2, 3, 4};
    for (int i = 0; i < N; ++i) { r[0] += i; stsm_x4(base + (i & 15) * 32, r); }
  } else if (mode == 4) {  // LDSM.T only (throughput)
    uint32_t f0[4]; uint32_t acc = 0;
    for (int i = 0; i < N; ++i) { ldsm_x4_t(f0, base + (i & 15) * 32); acc ^= f0[0] ^ f0[3]; }
    c[0][0] = __uint_as_float(acc);
  }
  long long t1 = clock64();
  if (lane == 0) out[mode * 8 + w] = t1 - t0;
  float s = 0; for (int j = 0; j < 8; ++j) for (int q = 0; q < 4; ++q) s += c[j][q];
  sink[threadIdx.x + 2048] = s;
  __syncwarp();
  if (lane == 0) atomicAdd((int*)(sm + 8 * 16 * RS + 64 * 1024), 1);
}
int main() {
  long long* out; float* sink;
  cudaMalloc(&out, 64 * 8); cudaMalloc(&sink, 8192 * 4);
  size_t smem = 8 * 16 * 1040 + 64 * 1024 + 64;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const char* names[] = {"HMMA 8 chains", "LDSM.T + 2 HMMA pipelined", "LDSM + 2 HMMA pipelined", "STSM stream", "LDSM.T stream"};
  for (int bg = 0; bg <= 2; ++bg)
    for (int nbg : {0, 4, 8}) {
      if ((bg == 0) != (nbg == 0)) continue;
      for (int m = 0; m < 5; ++m) {
        k<<<1, (4 + nbg) * 32, smem>>>(out, sink, m, bg);
        k<<<1, (4 + nbg) * 32, smem>>>(out, sink, m, bg);
      }
      cudaError_t e = cudaDeviceSynchronize();
      long long h[64]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
      for (int m = 0; m < 5; ++m)
        printf("bg %s x%d  %-28s %7.1f cycles per loop step (warp 0)  %s\n", bg == 0 ? "none " : bg == 1 ? "FFMA2" : "smem ", nbg, names[m], (double)h[m * 8] / N, cudaGetErrorString(e));
    }
  return 0;
}
