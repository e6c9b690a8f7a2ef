// CriterionContrastive on MATERIALISED logits (src/clip-event/model_clip.py:633-662): the reference's
// criterion accepts any logits tensors, so besides the fused GEMM+CE path (contrastive.cu) the
// drop-in needs the plain form: mean cross-entropy over the rows of a dense [rows, cols] matrix,
// optionally over the rows picked by an index list (logits_per_text.index_select(index_pos)), and
// BCE-with-logits for the 'bce' image side.  Memory-bound: every selected row is read once in the
// forward and once in the backward (16-byte loads), one CTA per row.
#include "ce_common.cuh"

namespace ce {
namespace {

constexpr int kThreads = 256;

template <int DT>
__device__ __forceinline__ float load_elem(const void* base, int64_t i) {
  return In<DT>::ld(reinterpret_cast<const typename In<DT>::type*>(base) + i);
}

__device__ __forceinline__ void merge_ml(float& m, float& l, float m2, float l2) {
  const float mn = fmaxf(m, m2);
  if (mn == -INFINITY) return;
  l = l * expf(m - mn) + l2 * expf(m2 - mn);
  m = mn;
}

// items[p] = LSE(row) - row[label];  lse[p] kept for the backward.  kind 1: BCE-with-logits row sums.
template <int DT>
__global__ void __launch_bounds__(kThreads) dense_ce_rows_kernel(const void* logits, int64_t ld, int64_t rows_total,
                                                                 int cols, const int64_t* row_index, int n,
                                                                 const void* labels, int labels_by_row, int kind,
                                                                 float* lse, float* items, int* bad) {
  using T = typename In<DT>::type;
  constexpr int V = In<DT>::kVec;
  const int p = blockIdx.x;
  int64_t r = row_index != nullptr ? row_index[p] : p;
  bool err = r < 0 || r >= rows_total;
  if (err) r = 0;
  const T* row = reinterpret_cast<const T*>(logits) + r * ld;
  const bool vec = (ld % V == 0) && (cols % V == 0) && ((reinterpret_cast<uintptr_t>(logits) & 15) == 0);
  __shared__ float sm[kThreads / 32], sl[kThreads / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (kind == 1) {   // BCEWithLogitsLoss: mean over all elements of softplus(l) - y l
    const float* y = reinterpret_cast<const float*>(labels) + (int64_t)(labels_by_row ? r : p) * cols;
    float s = 0.f;
    for (int c = threadIdx.x; c < cols; c += kThreads) {
      const float l = In<DT>::ld(row + c);
      s += fmaxf(l, 0.f) + log1pf(expf(-fabsf(l))) - y[c] * l;
    }
    s = warp_sum(s);
    if (lane == 0) sl[w] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int i = 0; i < kThreads / 32; ++i) t += sl[i];
      items[p] = t / (float)cols;
      lse[p] = 0.f;
    }
    return;
  }
  float m = -INFINITY, l = 0.f;
  if (vec) {
    for (int c = threadIdx.x * V; c < cols; c += kThreads * V) {
      float v[8];
      In<DT>::load16(row + c, v);
      float cm = -INFINITY;
#pragma unroll
      for (int i = 0; i < V; ++i) cm = (c + i < cols) ? fmaxf(cm, v[i]) : cm;
      const float mn = fmaxf(m, cm);
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < V; ++i) s += (c + i < cols) ? expf(v[i] - mn) : 0.f;
      l = l * expf(m - mn) + s;
      m = mn;
    }
  } else {
    for (int c = threadIdx.x; c < cols; c += kThreads) {
      const float v = In<DT>::ld(row + c);
      const float mn = fmaxf(m, v);
      l = l * expf(m - mn) + expf(v - mn);
      m = mn;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), l2 = __shfl_xor_sync(0xffffffffu, l, o);
    merge_ml(m, l, m2, l2);
  }
  if (lane == 0) { sm[w] = m; sl[w] = l; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float M = -INFINITY, L = 0.f;
    for (int i = 0; i < kThreads / 32; ++i) merge_ml(M, L, sm[i], sl[i]);
    const float e = M + logf(L);
    int64_t lab = reinterpret_cast<const int64_t*>(labels)[labels_by_row ? r : p];
    if (lab < 0 || lab >= cols) { err = true; lab = 0; }
    lse[p] = e;
    items[p] = e - In<DT>::ld(row + lab);
    if (err) atomicOr(bad, 1);     // the reference raises an IndexError here; the loss comes back as NaN
  }
}

// loss = mean(items), fixed order; NaN when an index was out of range.
__global__ void __launch_bounds__(1024) dense_ce_mean_kernel(const float* items, int n, const int* bad, float* loss) {
  __shared__ double sh[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 1024) s += (double)items[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 32; ++i) t += sh[i];
    *loss = *bad ? __int_as_float(0x7fc00000) : (float)(t / (double)n);
  }
}

// d logits[r, :] (+)= g/n * (softmax(row) - onehot)   (kind 1: g/(n cols) * (sigmoid(l) - y)).
// Rows named twice by the index list accumulate (atomic adds into the zero-filled fp32 matrix).
template <int DT>
__global__ void __launch_bounds__(kThreads) dense_ce_bwd_kernel(const void* logits, int64_t ld, int64_t rows_total,
                                                                int cols, const int64_t* row_index, int n,
                                                                const void* labels, int labels_by_row, int kind,
                                                                const float* lse, const float* g, float* dlogits,
                                                                int64_t ldd) {
  using T = typename In<DT>::type;
  const int p = blockIdx.x;
  int64_t r = row_index != nullptr ? row_index[p] : p;
  if (r < 0 || r >= rows_total) return;
  const T* row = reinterpret_cast<const T*>(logits) + r * ld;
  float* out = dlogits + r * ldd;
  const float coef = __ldg(g) / (float)n;
  if (kind == 1) {
    const float* y = reinterpret_cast<const float*>(labels) + (int64_t)(labels_by_row ? r : p) * cols;
    for (int c = threadIdx.x; c < cols; c += kThreads) {
      const float l = In<DT>::ld(row + c);
      out[c] = coef / (float)cols * (1.f / (1.f + expf(-l)) - y[c]);
    }
    return;
  }
  const float e = lse[p];
  int64_t lab = reinterpret_cast<const int64_t*>(labels)[labels_by_row ? r : p];
  const bool accumulate = row_index != nullptr;
  for (int c = threadIdx.x; c < cols; c += kThreads) {
    const float v = coef * (expf(In<DT>::ld(row + c) - e) - (c == lab ? 1.f : 0.f));
    if (accumulate) atomicAdd(out + c, v);
    else out[c] = v;
  }
}

}  // namespace
}  // namespace ce

using namespace ce;

extern "C" size_t ce_dense_ce_workspace_bytes(int n) { return sizeof(float) * 2 * (size_t)(n > 0 ? n : 0) + 256; }

extern "C" int ce_dense_ce_fwd(const void* logits, int64_t ld, int64_t rows_total, int cols,
                               const int64_t* row_index, int n, const void* labels, int labels_by_row,
                               int kind, int dtype, float* loss, void* workspace, size_t workspace_bytes,
                               ce_stream_t stream) {
  CE_TRY(check_device());
  if (dtype != CE_F32 && dtype != CE_BF16) return fail(CE_ERR_DTYPE, "dense CE: unknown dtype %d", dtype);
  if (rows_total < 1 || cols < 1 || n < 1 || ld < cols) return fail(CE_ERR_SHAPE, "dense CE: bad shape (rows=%lld cols=%d n=%d ld=%lld)", (long long)rows_total, cols, n, (long long)ld);
  if (kind != 0 && kind != 1) return fail(CE_ERR_ARG, "dense CE: kind must be 0 (cross-entropy) or 1 (BCE with logits)");
  if (workspace_bytes < ce_dense_ce_workspace_bytes(n)) return fail(CE_ERR_WORKSPACE, "dense CE: workspace too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float* lse = reinterpret_cast<float*>(workspace);
  float* items = lse + n;
  int* bad = reinterpret_cast<int*>(items + n);
  CE_CUDA_TRY(cudaMemsetAsync(bad, 0, sizeof(int), st));
  if (dtype == CE_F32) dense_ce_rows_kernel<CE_F32><<<n, kThreads, 0, st>>>(logits, ld, rows_total, cols, row_index, n, labels, labels_by_row, kind, lse, items, bad);
  else dense_ce_rows_kernel<CE_BF16><<<n, kThreads, 0, st>>>(logits, ld, rows_total, cols, row_index, n, labels, labels_by_row, kind, lse, items, bad);
  CE_LAUNCH_CHECK();
  dense_ce_mean_kernel<<<1, 1024, 0, st>>>(items, n, bad, loss);
  CE_LAUNCH_CHECK();
  return CE_OK;
}

extern "C" int ce_dense_ce_bwd(const void* logits, int64_t ld, int64_t rows_total, int cols,
                               const int64_t* row_index, int n, const void* labels, int labels_by_row,
                               int kind, int dtype, const float* g, float* dlogits, int64_t ldd,
                               const void* workspace, ce_stream_t stream) {
  CE_TRY(check_device());
  if (dtype != CE_F32 && dtype != CE_BF16) return fail(CE_ERR_DTYPE, "dense CE: unknown dtype %d", dtype);
  if (rows_total < 1 || cols < 1 || n < 1 || ld < cols || ldd < cols) return fail(CE_ERR_SHAPE, "dense CE: bad shape");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const float* lse = reinterpret_cast<const float*>(workspace);
  if (row_index != nullptr) CE_CUDA_TRY(cudaMemsetAsync(dlogits, 0, sizeof(float) * (size_t)rows_total * ldd, st));
  if (dtype == CE_F32) dense_ce_bwd_kernel<CE_F32><<<n, kThreads, 0, st>>>(logits, ld, rows_total, cols, row_index, n, labels, labels_by_row, kind, lse, g, dlogits, ldd);
  else dense_ce_bwd_kernel<CE_BF16><<<n, kThreads, 0, st>>>(logits, ld, rows_total, cols, row_index, n, labels, labels_by_row, kind, lse, g, dlogits, ldd);
  CE_LAUNCH_CHECK();
  return CE_OK;
}

// ------------------------------------------------------------------------------------------
// engine.py:89-90 for the loss head's own parameter: torch.nn.utils.clip_grad_norm_(params, max_norm)
// followed by optimizer.step() (torch.optim.SGD with momentum / torch.optim.Adam, engine.py:133-149),
// on the scalar logit_scale.  One launch instead of the dozen element-wise launches the foreach
// optimiser spends on a one-element tensor; the clip coefficient uses the squared gradient norm of
// all OTHER parameters (the encoders', computed by the caller) plus this parameter's own.
// ------------------------------------------------------------------------------------------
namespace ce {
namespace {
__global__ void head_param_step_kernel(float* p, float* grad, float* s0, float* s1, float* step,
                                       const float* other_sq, float max_norm, int kind, float lr, float b1,
                                       float b2, float eps, float wd, float* clip_coef_out) {
  float g = *grad;
  float coef = 1.f;
  if (max_norm > 0.f) {
    const float total = sqrtf((other_sq != nullptr ? *other_sq : 0.f) + g * g);
    coef = fminf(max_norm / (total + 1e-6f), 1.f);       // clip_grad_norm_: clamp(max_norm / (norm + 1e-6), max = 1)
    g *= coef;
    *grad = g;                                            // the reference scales .grad in place
  }
  if (clip_coef_out != nullptr) *clip_coef_out = coef;
  float w = *p;
  const float t = *step + 1.f;
  *step = t;
  if (wd != 0.f) g = fmaf(wd, w, g);                      // L2 weight decay as both optimisers apply it
  if (kind == 0) {                                        // torch.optim.SGD(momentum = b1)
    float buf = g;
    if (b1 != 0.f) {
      buf = t == 1.f ? g : fmaf(b1, *s0, g);
      *s0 = buf;
    }
    w -= lr * buf;
  } else {                                                // torch.optim.Adam
    const float m = fmaf(b1, *s0, (1.f - b1) * g);        // lerp(exp_avg, g, 1 - b1)
    const float v = fmaf(b2, *s1, (1.f - b2) * g * g);
    *s0 = m; *s1 = v;
    const float bc1 = 1.f - powf(b1, t), bc2 = 1.f - powf(b2, t);
    const float denom = sqrtf(v) / sqrtf(bc2) + eps;
    w -= (lr / bc1) * (m / denom);
  }
  *p = w;
}
}  // namespace
}  // namespace ce

extern "C" int ce_head_param_step(float* param, float* grad, float* state0, float* state1, float* step,
                                  const float* other_grad_sq, float max_norm, int kind, float lr,
                                  float beta1, float beta2, float eps, float weight_decay,
                                  float* clip_coef_out, ce_stream_t stream) {
  CE_TRY(check_device());
  if (kind != 0 && kind != 1) return fail(CE_ERR_ARG, "head_param_step: kind must be 0 (SGD) or 1 (Adam)");
  if (param == nullptr || grad == nullptr || state0 == nullptr || step == nullptr || (kind == 1 && state1 == nullptr))
    return fail(CE_ERR_ARG, "head_param_step: null parameter / gradient / optimiser state");
  head_param_step_kernel<<<1, 1, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      param, grad, state0, state1, step, other_grad_sq, max_norm, kind, lr, beta1, beta2, eps, weight_decay, clip_coef_out);
  CE_LAUNCH_CHECK();
  return CE_OK;
}
