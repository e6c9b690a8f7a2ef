// mma.sync (legacy tensor path, HMMA) latency / throughput on sm_100a, and ldmatrix latency.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/build/hmma tools/ubench/hmma.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define N 512
__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t* r, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__global__ void k(long long* out, float* sink, int mode) {
  __shared__ __align__(128) uint32_t sm[4096];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = 0x3c003c00u + (i & 0) ;
  __syncthreads();
  uint32_t a[4] = {0x3f803f80u, 0x3f803f80u, 0, 0};
  uint32_t b0 = 0x3f003f00u, b1 = 0x3f003f00u;
  float c[8][4];
  for (int j = 0; j < 8; ++j) for (int q = 0; q < 4; ++q) c[j][q] = 0.f;
  uint32_t base = (uint32_t)__cvta_generic_to_shared(sm) + (lane & 15) * 64 + (lane >> 4) * 16 + w * 1024;
  long long t0 = clock64();
  if (mode == 0) { for (int i = 0; i < N; ++i) mma16816(c[0], a, b0, b1); }                       // dependent chain
  else if (mode == 1) { for (int i = 0; i < N; i += 2) { mma16816(c[0], a, b0, b1); mma16816(c[1], a, b0, b1); } }
  else if (mode == 2) { for (int i = 0; i < N; i += 4) { mma16816(c[0], a, b0, b1); mma16816(c[1], a, b0, b1); mma16816(c[2], a, b0, b1); mma16816(c[3], a, b0, b1);} }
  else if (mode == 3) { for (int i = 0; i < N; i += 8) { for (int j = 0; j < 8; ++j) mma16816(c[j], a, b0, b1); } }
  else if (mode == 4) {   // ldmatrix dependent chain (address depends on loaded value)
    uint32_t r[4]; uint32_t ad = base;
    for (int i = 0; i < N; ++i) { ldsm_x4(r, ad); ad = base + (r[0] & 0); }
    a[0] = r[1];
    mma16816(c[0], a, b0, b1);
  } else if (mode == 5) {  // ldsm -> mma dependent (as in a k loop without unrolling)
    uint32_t r[4];
    for (int i = 0; i < N; ++i) { ldsm_x4(r, base + (i & 7) * 32); mma16816(c[0], r, b0, b1); mma16816(c[1], r, b1, b0); }
  }
  long long t1 = clock64();
  if (lane == 0) out[mode * 32 + w] = t1 - t0;
  float s = 0; for (int j = 0; j < 8; ++j) for (int q = 0; q < 4; ++q) s += c[j][q];
  sink[threadIdx.x + 1024 * mode] = s;
}
int main() {
  long long* out; float* sink;
  cudaMalloc(&out, 32 * 8 * 8); cudaMalloc(&sink, 1024 * 8 * 4);
  const char* names[] = {"HMMA dependent chain", "HMMA 2 chains", "HMMA 4 chains", "HMMA 8 chains", "LDSM dependent", "LDSM -> 2 HMMA loop"};
  for (int nw : {1, 4, 8, 16}) {
    for (int rep = 0; rep < 2; ++rep) for (int m = 0; m < 6; ++m) k<<<1, nw * 32>>>(out, sink, m);
    cudaDeviceSynchronize();
    long long h[32 * 8]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    for (int m = 0; m < 6; ++m) printf("warps %2d  %-24s %7.1f cycles per HMMA/LDSM (warp 0)\n", nw, names[m], (double)h[m * 32] / N);
  }
  return 0;
}
