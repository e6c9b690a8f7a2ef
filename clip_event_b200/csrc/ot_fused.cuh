// Fused, shared-memory-resident OT kernel (csrc/ot_fused.cu): launcher interface.
#pragma once

#include <stdlib.h>

#include "ce_common.cuh"

namespace ce {

struct OtFusedArgs {
  const void* txt;            // [B, M, D] bf16, sample stride txt_bs elements
  const void* img;            // [B, N, D] bf16 (after the whole-image slot), sample stride img_bs
  int64_t txt_bs, img_bs;
  const void* txt_mask;
  const void* img_mask;
  int64_t txt_ms, img_ms;
  int mask_kind;
  int B, M, N, D;
  float beta, eps, scale;
  int iters, k;
  float* dist;                // [B]
  void* dtxt;                 // nullable (forward only): same layout as txt
  void* dimg;                 // same layout as img
  void* dslot0;               // nullable: [B] rows of D elements, stride img_bs, zero-filled
  // stream kernel only -- packed (variable-length) node layout: rows of sample b are txt_off[b] .. txt_off[b+1]-1 of a
  // dense [sum_m, D] matrix (likewise img_off); every row is a valid node, M / N are the maxima over the batch,
  // masks and sample strides are ignored and gradients come back in the same packed layout
  const int* txt_off;
  const int* img_off;
  int slots;                  // filled in by the launcher (fused: resident slots; stream: parks)
  int cy_depth, gy_depth;     // stream kernel: chunk ring depths of the cost / gradient stage
  int poll_mode;              // debug (CE_OT_POLL)
  int dbg;                    // debug (CE_OT_DBG): 1 = no MMA work, 2 = no bulk copies (results are garbage)
  long long* trace;           // debug timeline buffer (CE_OT_TRACE_PTR), normally null
};

// bf16, M <= 16, N <= 64, D a multiple of 128 up to 768, and at least two samples fit in shared memory
bool ot_fused_supported(int M, int N, int D, int dtype);
int ot_fused_slots(int M, int N, int D);
size_t ot_fused_smem_bytes(int M, int N, int D, int slots);
int launch_ot_fused(OtFusedArgs a, cudaStream_t st);

// Streaming variant (csrc/ot_stream.cu): bf16, M <= 16, N <= 64, D a multiple of 64 up to 512
bool ot_stream_supported(int M, int N, int D, int dtype);
size_t ot_stream_smem_bytes(int D, int parks, int cy_depth, int gy_depth);
int launch_ot_stream(OtFusedArgs a, cudaStream_t st);

// Gradient contraction for the plans the streaming kernel does not take (csrc/ot_wide.cu): bf16, D a multiple of 64.
// TMA-fed persistent kernel; W / ax / ay are the solver's outputs (ot_ipot_kernel), dx_acc the fp32 accumulator
// of samples whose image rows are split over several CTAs (ot_wide_nsplit > 1).
struct OtWideGradArgs {
  const void* txt; const void* img;
  int64_t txt_bs, img_bs;
  int B, M, N, D;
  const float* W; const float* ax; const float* ay;
  int Nld;
  void* dtxt; void* dimg;
  float* dx_acc;
};
bool ot_wide_supported(int M, int N, int D, int dtype);
int ot_wide_nsplit(int M, int N, int D);
int launch_ot_wide_grad(const OtWideGradArgs& g, cudaStream_t st);
// Cost contraction of the same family: S = y x^t (fp32 [B, N, MP]) and the rows' sums of squares.
struct OtWideCostArgs {
  const void* txt; const void* img;
  int64_t txt_bs, img_bs;
  int B, M, N, D;
  float* S; float* nx2; float* ny2;
  int Nld;
};
bool ot_wide_cost_supported(int M, int N, int D, int dtype);
int launch_ot_wide_cost(const OtWideCostArgs& g, cudaStream_t st);

}  // namespace ce
