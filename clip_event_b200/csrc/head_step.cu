// The glue either side of the one-call loss head (functional._LossHeadStep / distributed._GlobalLossHeadStep;
// engine.py:48-67 forms the losses, :67 sums them, :88 back-propagates the sum):
//   ce_head_losses_cast  the three fp32 losses -> the dtypes the reference's criteria return them in
//                        (model_clip.py:646-659 in the logits' dtype, :699-707 in the node embeddings'), one launch
//                        instead of a copy and three casts;
//   ce_head_step_scale   every stashed gradient buffer times the upstream gradient of ITS loss, read in the dtype
//                        autograd delivers it in -- one launch that returns on the device when they are all 1 (as
//                        under sum(loss_dict.values()).backward()), instead of three casts and three scale launches.
#include <algorithm>

#include "ce_common.cuh"

namespace ce {
namespace {

constexpr int kMaxSeg = 8;

struct ScaleSeg {
  void* p;
  int64_t n;
  int dtype;
  int which;   // 0: scaled by g_i (== g_t, or everything becomes NaN), 1: scaled by g_ot
};
struct ScaleArgs {
  ScaleSeg seg[kMaxSeg];
  int nseg;
  const void *g_i, *g_t, *g_ot;
  int gdt_c, gdt_o;
};

__device__ __forceinline__ float ld_scalar(const void* p, int dt) {
  return dt == CE_BF16 ? __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(p)) : *reinterpret_cast<const float*>(p);
}

template <int DT>
__device__ __forceinline__ void scale_span(void* vp, int64_t n, float s) {
  using T = typename In<DT>::type;
  constexpr int V = In<DT>::kVec;
  T* p = reinterpret_cast<T*>(vp);
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (int64_t)gridDim.x * blockDim.x;
  // scalar head up to the first 16-byte boundary, 16-byte body, scalar tail
  int64_t head = ((16 - (reinterpret_cast<uintptr_t>(p) & 15)) & 15) / sizeof(T);
  if (head > n) head = n;
  const int64_t nvec = (n - head) / V;
  for (int64_t i = tid; i < head; i += nthr) In<DT>::st(p + i, In<DT>::ld(p + i) * s);
  T* body = p + head;
  for (int64_t i = tid; i < nvec; i += nthr) {
    float v[V];
    In<DT>::load16(body + i * V, v);
    if constexpr (DT == CE_F32) {
      *reinterpret_cast<float4*>(body + i * V) = make_float4(v[0] * s, v[1] * s, v[2] * s, v[3] * s);
    } else {
      uint32_t w[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * j] * s, v[2 * j + 1] * s);
        w[j] = *reinterpret_cast<uint32_t*>(&h);
      }
      *reinterpret_cast<uint4*>(body + i * V) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
  for (int64_t i = head + nvec * V + tid; i < n; i += nthr) In<DT>::st(p + i, In<DT>::ld(p + i) * s);
}

__global__ void __launch_bounds__(256) head_step_scale_kernel(const __grid_constant__ ScaleArgs a) {
  float sc = 1.f, so = 1.f;
  if (a.g_i != nullptr) {
    sc = ld_scalar(a.g_i, a.gdt_c);
    // the contrastive gradients were formed for EQUAL upstream gradients of loss_i and loss_t: anything else must
    // not pass silently
    if (ld_scalar(a.g_t, a.gdt_c) != sc) sc = __int_as_float(0x7fc00000);
  }
  if (a.g_ot != nullptr) so = ld_scalar(a.g_ot, a.gdt_o);
  if (sc == 1.f && so == 1.f) return;
  for (int k = 0; k < a.nseg; ++k) {
    const ScaleSeg& sg = a.seg[k];
    const float s = sg.which == 0 ? sc : so;
    if (s == 1.f) continue;
    if (sg.dtype == CE_BF16) scale_span<CE_BF16>(sg.p, sg.n, s);
    else scale_span<CE_F32>(sg.p, sg.n, s);
  }
}

__device__ __forceinline__ void st_scalar(void* p, int idx, int dt, float v) {
  if (dt == CE_BF16) reinterpret_cast<__nv_bfloat16*>(p)[idx] = __float2bfloat16_rn(v);
  else reinterpret_cast<float*>(p)[idx] = v;
}

__global__ void head_losses_cast_kernel(const float* li, const float* lt, const float* lo, void* out_c, int dt_c,
                                        void* out_o, int dt_o) {
  const int t = threadIdx.x;
  if (t == 0 && li != nullptr) st_scalar(out_c, 0, dt_c, *li);
  if (t == 1 && lt != nullptr) st_scalar(out_c, 1, dt_c, *lt);
  if (t == 2 && lo != nullptr) st_scalar(out_o, 0, dt_o, *lo);
}

bool known(int dt) { return dt == CE_F32 || dt == CE_BF16; }

}  // namespace
}  // namespace ce

using namespace ce;

extern "C" int ce_head_losses_cast(const float* loss_i, const float* loss_t, const float* loss_ot, void* out_c,
                                   int dtype_c, void* out_o, int dtype_o, ce_stream_t stream) {
  CE_TRY(check_device());
  if ((loss_i != nullptr || loss_t != nullptr) && (out_c == nullptr || !known(dtype_c)))
    return fail(CE_ERR_DTYPE, "losses cast: contrastive output missing or of unknown dtype %d", dtype_c);
  if (loss_ot != nullptr && (out_o == nullptr || !known(dtype_o)))
    return fail(CE_ERR_DTYPE, "losses cast: OT output missing or of unknown dtype %d", dtype_o);
  head_losses_cast_kernel<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(loss_i, loss_t, loss_ot, out_c, dtype_c,
                                                                               out_o, dtype_o);
  CE_LAUNCH_CHECK();
  return CE_OK;
}

extern "C" int ce_head_step_scale(void* const* bufs, const int64_t* counts, const int* dtypes, const int* which,
                                  int nbuf, const void* g_i, const void* g_t, int g_dtype_c, const void* g_ot,
                                  int g_dtype_o, ce_stream_t stream) {
  CE_TRY(check_device());
  if (nbuf < 0 || nbuf > kMaxSeg) return fail(CE_ERR_SHAPE, "head step scale: %d buffers (at most %d)", nbuf, kMaxSeg);
  if ((g_i == nullptr) != (g_t == nullptr))
    return fail(CE_ERR_SHAPE, "head step scale: loss_i and loss_t must be back-propagated together");
  if (g_i != nullptr && !known(g_dtype_c)) return fail(CE_ERR_DTYPE, "head step scale: unknown gradient dtype %d", g_dtype_c);
  if (g_ot != nullptr && !known(g_dtype_o)) return fail(CE_ERR_DTYPE, "head step scale: unknown gradient dtype %d", g_dtype_o);
  ScaleArgs a{};
  int64_t longest = 0;
  for (int k = 0; k < nbuf; ++k) {
    if (!known(dtypes[k])) return fail(CE_ERR_DTYPE, "head step scale: buffer %d of unknown dtype %d", k, dtypes[k]);
    if (which[k] != 0 && which[k] != 1) return fail(CE_ERR_SHAPE, "head step scale: buffer %d belongs to loss %d", k, which[k]);
    if ((which[k] == 0 ? g_i : g_ot) == nullptr || bufs[k] == nullptr || counts[k] <= 0) continue;   // nothing to scale by
    a.seg[a.nseg++] = ScaleSeg{bufs[k], counts[k], dtypes[k], which[k]};
    longest = std::max(longest, counts[k]);
  }
  if (a.nseg == 0) return CE_OK;
  a.g_i = g_i; a.g_t = g_t; a.g_ot = g_ot; a.gdt_c = g_dtype_c; a.gdt_o = g_dtype_o;
  const int blocks = (int)std::min<int64_t>((longest + 2047) / 2048, (int64_t)num_sms() * 8);
  head_step_scale_kernel<<<std::max(blocks, 1), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  CE_LAUNCH_CHECK();
  return CE_OK;
}
