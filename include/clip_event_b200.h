/* clip_event_b200.h -- C ABI of the B200-native CLIP-Event training loss head.
 *
 * Drop-in boundary for ONE hot path of limanling/clip-event:
 *   similarity scoring   src/clip-event/model_clip.py:495-528   (CLIP.forward tail)
 *   InfoNCE criterion    src/clip-event/model_clip.py:620-662   (CriterionContrastive)
 *   OT alignment loss    src/clip-event/model_clip.py:664-715   (CriterionAlignment)
 *   IPOT solver          src/clip-event/model_ot.py:8-84
 * forward and backward.  The reference is pure PyTorch and has no FFI of its own, so these
 * entry points are what a ctypes/cffi stub inside the reference's model_clip.py / model_ot.py
 * would bind (INTEGRATION.md shows that stub).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; the caller owns all memory,
 *     including workspaces (sizes from the *_workspace_bytes queries); the library never allocates
 *     device memory; its only state is a thread-local error string, the launch counter behind
 *     ce_debug_launch_count() and a few debug switches read once from the environment
 *     (CE_GEMM_PAIR, CE_CTR_STORED, CE_OT_STREAM, CE_OT_FUSED, CE_OT_PARKS / CE_OT_CY / CE_OT_GY, CE_OT_TRACE_PTR --
 *     tuning aids, not configuration); CE_PDL=0 (read at every launch) turns off the programmatic dependent
 *     launches between the kernels of one ce_contrastive_* call;
 *   - inputs are assumed finite: the streaming OT kernel re-uses its shared-memory row buffers from sample to
 *     sample and multiplies rows that are not loaded for a sample (padding beyond the node count) by exact
 *     zeros, so a NaN / Inf in one sample can surface in the next sample handled by the same thread block;
 *   - tensors are row-major and contiguous unless a stride argument says otherwise; embedding
 *     pointers must be 16-byte aligned and D a multiple of 8;
 *   - dtype: CE_F32 = fp32 in / fp32 out, tensor-core products as 3xTF32 (fp32-level accuracy);
 *            CE_BF16 = bf16 in / bf16 out, fp32 accumulation and fp32 softmax / solver state;
 *   - all work is enqueued on `stream` (a cudaStream_t); no host synchronisation inside;
 *   - return 0 on success, a negative CE_ERR_* for bad arguments, or a positive cudaError_t;
 *     ce_last_error() describes the last failure on the calling thread.
 *   - sm_100a only: every entry point returns CE_ERR_ARCH on another device.  No CPU fallback.
 */
#ifndef CLIP_EVENT_B200_H_
#define CLIP_EVENT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* ce_stream_t; /* cudaStream_t */

enum {
  CE_OK = 0,
  CE_ERR_SHAPE = -1,
  CE_ERR_DTYPE = -2,
  CE_ERR_ALIGN = -3,
  CE_ERR_WORKSPACE = -4,
  CE_ERR_ARCH = -5,
  CE_ERR_ARG = -6
};
enum { CE_F32 = 0, CE_BF16 = 1 };
/* image-side loss of CriterionContrastive (model_clip.py:624-627, 646-651); the text side is always
 * the over-batch cross-entropy of the positive descriptions (model_clip.py:504, 655-659) */
enum { CE_IMG_CE_OVERBATCH = 0 /* labels_i: int64 [R] GLOBAL column of each image's positive      */,
       CE_IMG_CE_INSTANCE = 1  /* labels_i: int64 [b] in [0,T), image vs its own T descriptions    */,
       CE_IMG_BCE_INSTANCE = 2 /* labels_i: fp32 [b, T] targets, BCEWithLogits over the [b, T] tile */ };
/* node-mask encodings accepted by the OT entry points */
enum { CE_MASK_NUM_I64 = 0 /* reference `*_num`: int64, valid = nonzero  (model_clip.py:688-690) */,
       CE_MASK_PAD_U8 = 1  /* reference `*_pad`: bool/uint8, pad = nonzero (model_ot.py:66-74)   */ };

int ce_version(void);
const char* ce_last_error(void);
/* 0 if the current CUDA device is sm_100 (B200), CE_ERR_ARCH otherwise. */
int ce_device_check(void);

/* ------------------------------------------------------------------------------------------
 * Similarity scoring + InfoNCE, over-batch mode  (model_clip.py:496-508 + 633-662, 'ce')
 *
 * Rows are images, columns are descriptions.  One fused tcgen05 GEMM L = s * I^ T^t is never
 * materialised: its epilogue produces the row log-sum-exp (image side) and, through the identity
 * logits_per_text[index_pos] == L[:, index_pos]^t, the column log-sum-exp of the positive columns
 * (text side).
 *
 * The entry points are written for a column-sharded multi-GPU run: `img` holds ALL
 * R (gathered) images, `txt` the C LOCAL descriptions whose global column index starts at
 * `col_offset`.  Single GPU: R = B, C = B*T, col_offset = 0, world = 1.
 *
 *   img          [R, D]   image_features (un-normalised)
 *   txt          [C, D]   text_features  (un-normalised), local columns
 *   logit_scale  [1]      fp32, the log of the temperature inverse (model_clip.py:330,502)
 *   labels_i     see image_loss; over batch: int64 [R] GLOBAL column index of each image's positive
 *                (labels_per_image).  Over-instance modes score the b = C/T LOCAL images, which
 *                are rows [row_offset, row_offset + b) of `img`, against their own T descriptions
 *   labels_t     [C]      int64 GLOBAL row index each local description belongs to (labels_per_text)
 *   index_pos    [P]      int64 LOCAL column indices used for the text-side loss  (index_pos)
 *
 * Index errors (labels_per_image outside [0, B*T) on one GPU / negative when sharded, index_pos
 * outside [0, C), labels_per_text[index_pos] outside [0, R), a description listed twice in
 * index_pos) raise an IndexError in the reference.  Here the indices are clamped -- nothing is read
 * or written out of bounds -- and BOTH LOSSES COME BACK AS NaN, which engine.py:79-82 turns into
 * "Loss is nan, stopping training".
 * ------------------------------------------------------------------------------------------ */
size_t ce_contrastive_workspace_bytes(int R, int C, int P, int D, int dtype);

/* Phase 1: GEMM + local statistics.
 *   row_part  [R, 4] fp32 out: per image (max2, sum2, positive logit, 0): running max and sum of
 *             the base-2 scaled logits over the LOCAL columns, and L[r, labels_i[r]] if that column
 *             is local (else 0).  All-gather these blocks across ranks.
 *   sums      [4]    fp32 out: {sum_p (colLSE_p - L[labels_t[pos_p], pos_p]), P, 0, 0}
 * Stashes norms / column LSE / split operands in `workspace` for the backward. */
int ce_contrastive_fwd_partial(const void* img, const void* txt, const float* logit_scale,
                               const void* labels_i, const int64_t* labels_t,
                               const int64_t* index_pos, int R, int C, int P, int D,
                               int64_t col_offset, int image_loss, int T, int64_t row_offset,
                               int dtype, float* row_part, float* sums,
                               void* workspace, size_t workspace_bytes, ce_stream_t stream);

/* Phase 2: merge `world` row_part blocks (block w at row_part_all + w*rank_stride floats, [R, 4]
 * each, this rank's own included), reduce the `world` sums blocks (block w at sums_all +
 * w*rank_stride, [4] each) and emit the losses -- so one all-gathered buffer of per-rank
 * [row_part | sums] records can be passed without repacking.  loss_i = mean_r(rowLSE - L[r,label]),
 * loss_t = mean_p(...) over the GLOBAL P.  Writes the global row LSE into the workspace (same
 * R, C, P, D, dtype as phase 1) for the backward. */
int ce_contrastive_fwd_finish(const float* row_part_all, const float* sums_all, int64_t rank_stride,
                              int world, int R, int C, int P, int D, int dtype, float* loss_i, float* loss_t,
                              void* workspace, size_t workspace_bytes, ce_stream_t stream);

/* Backward phase 1.  g_i / g_t are device scalars dL/dloss_i, dL/dloss_t; R_total / P_total the
 * global means' denominators.  Produces
 *   dtxt          [C, D] in `dtype` -- complete (normalisation backward applied)
 *   dimg_hat_part [R, D] fp32       -- s * G T^ for the LOCAL columns, BEFORE the normalisation
 *                                      backward; sum over ranks (reduce-scatter), then call
 *                                      ce_contrastive_bwd_finish on the local rows
 *   dlogit_scale_part [1] fp32      -- sum over ranks. */
int ce_contrastive_bwd_partial(const void* img, const void* txt, const float* logit_scale,
                               const void* labels_i, const int64_t* labels_t,
                               const int64_t* index_pos, int R, int C, int P, int D,
                               int64_t col_offset, int image_loss, int T, int64_t row_offset,
                               int dtype, const float* g_i, const float* g_t,
                               int R_total, int P_total, void* dtxt, float* dimg_hat_part,
                               float* dlogit_scale_part, void* workspace, size_t workspace_bytes,
                               ce_stream_t stream);

/* Backward phase 2: dimg = (d - i^ (i^ . d)) / |i| on `rows` rows (model_clip.py:496 backward). */
int ce_contrastive_bwd_finish(const void* img_rows, const float* dimg_hat_rows, int rows, int D,
                              int dtype, void* dimg_rows, ce_stream_t stream);

/* Single-GPU convenience wrappers (world = 1): forward then backward with the same workspace. */
int ce_contrastive_fwd(const void* img, const void* txt, const float* logit_scale,
                       const void* labels_i, const int64_t* labels_t, const int64_t* index_pos,
                       int B, int BT, int P, int D, int image_loss, int dtype, float* loss_i, float* loss_t,
                       void* workspace, size_t workspace_bytes, ce_stream_t stream);
int ce_contrastive_bwd(const void* img, const void* txt, const float* logit_scale,
                       const void* labels_i, const int64_t* labels_t, const int64_t* index_pos,
                       int B, int BT, int P, int D, int image_loss, int dtype, const float* g_i, const float* g_t,
                       void* dimg, void* dtxt, float* dlogit_scale, void* workspace,
                       size_t workspace_bytes, ce_stream_t stream);

/* Materialised logits for consumers that need the matrix itself
 * (src/preprocess/preprocess_description_contrastive.py:127-134):
 *   out[r, c] = exp(logit_scale) * <a_r, b_c> / (|a_r| |b_c|),  out is [Ra, Rb] fp32.
 * CLIP.forward returns (logits_per_image = f(img, txt), logits_per_text = f(txt, img)). */
int ce_similarity_logits(const void* a, const void* b, const float* logit_scale, int Ra, int Rb,
                         int D, int dtype, float* out, void* workspace, size_t workspace_bytes,
                         ce_stream_t stream);
size_t ce_similarity_workspace_bytes(int Ra, int Rb, int D, int dtype);

/* ------------------------------------------------------------------------------------------
 * OT graph alignment  (model_clip.py:679-715 + model_ot.py:8-84)
 *
 *   txt   [B, M, D]  text node embeddings, sample stride txt_bstride elements
 *   img   [B, N, D]  image node embeddings AFTER the whole-image slot: pass object_vec + D with
 *                    img_bstride = (N+1)*D to drop slot 0 without a copy (model_clip.py:686)
 *   masks per `mask_kind`; *_mstride is the per-sample stride of the mask arrays in elements
 *   beta / iters / k   IPOT parameters (model_ot.py:68: 0.5 / 50 / 1)
 *
 * ce_ot_fwd_bwd computes in ONE pass over the embeddings
 *   dist   [B] fp32      model_ot.py:83
 *   loss   [1] fp32      loss_scale * sum_b dist[b]         (model_clip.py:707: loss_scale = 0.01)
 *   dtxt / dimg          loss_scale * d(sum_b dist[b]) / d(txt, img), same layout/strides/dtype as
 *                        the inputs (pass NULL for both to skip the backward); IPOT is not
 *                        differentiated through (model_ot.py:32,81,83); padded entries get 0.
 *   dimg_slot0           if non-NULL, [B] rows of D elements with stride img_bstride that are
 *                        zero-filled (the dropped slot's gradient).
 * Scale the returned gradients by dL/dloss with ce_scale_inplace (a no-op launch when it is 1).
 * ------------------------------------------------------------------------------------------ */
size_t ce_ot_workspace_bytes(int B, int M, int N, int D);
int ce_ot_fwd_bwd(const void* txt, int64_t txt_bstride, const void* img, int64_t img_bstride,
                  const void* txt_mask, int64_t txt_mstride, const void* img_mask,
                  int64_t img_mstride, int mask_kind, int B, int M, int N, int D, int dtype,
                  float beta, int iters, int k, float loss_scale, float* dist, float* loss,
                  void* dtxt, void* dimg, void* dimg_slot0, void* workspace,
                  size_t workspace_bytes, ce_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * SURVEY.md 8f-2 -- the projections that feed the head: image side model_clip.py:253-260
 * (x[:, 0, :] -> ln_post -> @ proj), text side model_clip.py:412-415 (ln_final -> x[arange, eot] ->
 * @ text_projection).  hidden: [rows, L, W] last hidden states (sample_stride = L * W elements);
 * token_index: [rows] int64 token position per sample (NULL = token 0, the class token);
 * ln_w / ln_b [W] in the I/O dtype (both NULL = no LayerNorm); proj [W, D] row-major in the I/O dtype.
 * ce_proj_fwd writes feat [rows, D] (I/O dtype) and, if non-NULL, norm2 [rows] = the squared L2 norm of
 * each stored feature row (model_clip.py:496-497 needs it next).  ce_proj_bwd, called with the SAME
 * workspace after ce_proj_fwd, returns the gradient of the gathered token rows dx_rows [rows, W] (I/O
 * dtype; every other token's gradient is zero), dln_w / dln_b [W] fp32 and dproj [W, D] fp32.
 * W, D multiples of 8; fp32 mode runs the GEMMs as 3xTF32 like the rest of the library.
 * ------------------------------------------------------------------------------------------ */
size_t ce_proj_workspace_bytes(int rows, int W, int D, int dtype);
int ce_proj_fwd(const void* hidden, int64_t sample_stride, const int64_t* token_index, const void* ln_w,
                const void* ln_b, float eps, const void* proj, int rows, int W, int D, int dtype,
                void* feat, float* norm2, void* workspace, size_t workspace_bytes, ce_stream_t stream);
int ce_proj_bwd(const void* hidden, int64_t sample_stride, const int64_t* token_index, const void* ln_w,
                const void* proj, const void* dfeat, int rows, int W, int D, int dtype, void* dx_rows,
                float* dln_w, float* dln_b, float* dproj, void* workspace, size_t workspace_bytes,
                ce_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Device-side exchange for the sharded loss head (what utils.py:192-206 gathers, and the gradient return),
 * as plain loads from the peers' buffers over NVLink / NVSwitch -- no collective library in the data path.
 * peer_ptrs_host: HOST array of `world` device addresses, one per rank, of buffers with the same layout on
 * every rank (torch symmetric memory: hdl.buffer_ptrs).  The caller orders the steps with its own cross-rank
 * barrier (hdl.barrier()) before the call and before the buffers are overwritten.
 *   ce_p2p_gather     dst[r * bytes_each ...] = peer r's first bytes_each bytes            (all-gather)
 *   ce_p2p_reduce_f32 dst[i] = sum_r peer_r[offset_elems + i], i < n, summed in rank order (reduce-scatter:
 *                     offset = this rank's slice); tail_dst[j] = sum_r peer_r[tail_offset_elems + j], j < tail_n <= 32
 * Sizes / offsets in multiples of 16 bytes; world <= 16.
 * ------------------------------------------------------------------------------------------ */
int ce_p2p_gather(const int64_t* peer_ptrs_host, int world, int64_t bytes_each, void* dst, ce_stream_t stream);
int ce_p2p_reduce_f32(const int64_t* peer_ptrs_host, int world, int64_t offset_elems, int64_t n, float* dst,
                      int64_t tail_offset_elems, int tail_n, float* tail_dst, ce_stream_t stream);

/* Packed (variable-length) node sets -- SURVEY.md 8f-3; the reference pads every sample to the batch maximum
 * (model_clip.py:531-552, dataset_voa.py:532-544,566-577) and masks the padding afterwards.  Here the rows of
 * sample b are txt_off[b] .. txt_off[b+1]-1 of a dense [sum_m, D] matrix (img_off likewise, the whole-image slot
 * already removed); every row is a valid node.  bf16, max_m <= 16, max_n <= 64, D a multiple of 64 up to 512
 * (CE_ERR_SHAPE otherwise: pad and call ce_ot_fwd_bwd).  dtxt_rows / dimg_rows (both or neither) receive
 * loss_scale * d(sum dist)/d(rows) in the same packed layout.  Same value as ce_ot_fwd_bwd on the padded,
 * masked batch. */
int ce_ot_fwd_bwd_packed(const void* txt_rows, const int32_t* txt_off, const void* img_rows,
                         const int32_t* img_off, int B, int max_m, int max_n, int D, int dtype,
                         float beta, int iters, int k, float loss_scale, float* dist, float* loss,
                         void* dtxt_rows, void* dimg_rows, ce_stream_t stream);

/* The solver's building blocks with the reference's own granularity (model_ot.py), fp32:
 *   ce_ot_cost_matrix : cost_matrix_cosine  [B,M,D],[B,N,D] -> 1 - cos [B,M,N]  (model_ot.py:8-18)
 *   ce_ot_ipot        : ipot                C [B,M,N] -> T [B,N,M]              (model_ot.py:32-63)
 *   ce_ot_trace       : trace               [B,n,n] -> [B]                      (model_ot.py:21-29)
 */
int ce_ot_cost_matrix(const void* x, const void* y, int B, int M, int N, int D, int dtype,
                      float eps, float* cost, ce_stream_t stream);
int ce_ot_ipot(const float* cost, const uint8_t* x_pad, const uint8_t* y_pad, int B, int M, int N,
               float beta, int iters, int k, float* plan, ce_stream_t stream);
int ce_ot_trace(const float* x, int B, int n, float* out, ce_stream_t stream);

/* x[i] *= *g for i < n (n elements of `dtype`, rows of `row_len` elements `row_stride` apart);
 * returns immediately on the device when *g == 1. */
int ce_scale_inplace(void* x, int64_t rows, int64_t row_len, int64_t row_stride, int dtype,
                     const float* g, ce_stream_t stream);

/* Same, for gradients that were formed assuming EQUAL upstream gradients of two losses (the fused
 * step of the loss head under `sum(loss_dict.values()).backward()`, engine.py:67,88): x[i] *= *g, and
 * every element becomes NaN if *g_same != *g. */
int ce_scale_inplace_same(void* x, int64_t n, int dtype, const float* g, const float* g_same,
                          ce_stream_t stream);

/* The glue either side of the one-call loss head (engine.py:48-67 forms the losses, :67 sums them, :88
 * back-propagates the sum), one launch each:
 *   ce_head_losses_cast   out_c[0], out_c[1] = *loss_i, *loss_t in `dtype_c`; out_o[0] = *loss_ot in `dtype_o`
 *                         -- the dtypes the reference's criteria return (model_clip.py:646-659 follow the logits,
 *                         :699-707 the node embeddings); a NULL loss pointer is skipped.
 *   ce_head_step_scale    bufs[k] (counts[k] elements of dtypes[k]) *= the upstream gradient of its loss:
 *                         which[k] == 0 -> *g_i, and every element becomes NaN if *g_t != *g_i (the
 *                         contrastive gradients were formed for equal upstream gradients); which[k] == 1 ->
 *                         *g_ot.  The upstream gradients are device scalars of g_dtype_c / g_dtype_o (what
 *                         autograd delivers for losses of those dtypes); a NULL gradient leaves its buffers
 *                         alone.  Returns immediately on the device when every gradient is 1.  The arrays are
 *                         host arrays read during the call; at most 8 buffers. */
int ce_head_losses_cast(const float* loss_i, const float* loss_t, const float* loss_ot, void* out_c,
                        int dtype_c, void* out_o, int dtype_o, ce_stream_t stream);
int ce_head_step_scale(void* const* bufs, const int64_t* counts, const int* dtypes, const int* which,
                       int nbuf, const void* g_i, const void* g_t, int g_dtype_c, const void* g_ot,
                       int g_dtype_o, ce_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * CriterionContrastive on MATERIALISED logits (model_clip.py:633-662): the reference's criterion
 * accepts any logits tensors.  Memory-bound row kernels (one CTA per row, 16-byte loads):
 *   loss = mean_p ( LSE(logits[r_p, :]) - logits[r_p, label_p] ),  r_p = row_index ? row_index[p] : p
 *          -- nn.CrossEntropyLoss()(logits.index_select(0, row_index), labels...)   (kind 0)
 *   loss = mean over all elements of softplus(l) - y l  -- nn.BCEWithLogitsLoss()       (kind 1,
 *          labels fp32 [n, cols])
 *   labels: int64, indexed by p (labels_by_row = 0) or by r_p (labels_by_row = 1, i.e.
 *           labels_per_text.index_select(0, index_pos)); out-of-range indices give a NaN loss.
 * Backward: dlogits [rows_total, ldd] fp32 = g/n (softmax - onehot) on the selected rows (rows named
 * twice accumulate; with a row_index the matrix is zero-filled first).  `workspace` carries the row
 * LSEs from the forward.
 * ------------------------------------------------------------------------------------------ */
size_t ce_dense_ce_workspace_bytes(int n);
int ce_dense_ce_fwd(const void* logits, int64_t ld, int64_t rows_total, int cols,
                    const int64_t* row_index, int n, const void* labels, int labels_by_row, int kind,
                    int dtype, float* loss, void* workspace, size_t workspace_bytes, ce_stream_t stream);
int ce_dense_ce_bwd(const void* logits, int64_t ld, int64_t rows_total, int cols,
                    const int64_t* row_index, int n, const void* labels, int labels_by_row, int kind,
                    int dtype, const float* g, float* dlogits, int64_t ldd, const void* workspace,
                    ce_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * The step either side of the path, for the loss head's OWN parameter logit_scale:
 * torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm) + optimizer.step()  (engine.py:89-90,
 * optimisers built at engine.py:133-149) in one launch.  All pointers are fp32 device scalars.
 *   other_grad_sq   nullable: sum of squared gradients of every OTHER parameter (the encoders');
 *                   total_norm = sqrt(other_grad_sq + grad^2), clip_coef = min(1, max_norm/(total_norm+1e-6));
 *                   max_norm <= 0 disables clipping.  *grad is scaled in place, as the reference does.
 *   kind 0          torch.optim.SGD:  g += wd p; buf = (step 1 ? g : beta1 buf + g); p -= lr buf
 *                   (state0 = momentum buffer, beta1 = momentum; beta1 = 0: no buffer)
 *   kind 1          torch.optim.Adam: g += wd p; state0/state1 = exp_avg / exp_avg_sq; bias-corrected
 *   step            fp32 counter, incremented here;  clip_coef_out nullable (the encoders' gradients
 *                   need the same coefficient).
 * ------------------------------------------------------------------------------------------ */
int ce_head_param_step(float* param, float* grad, float* state0, float* state1, float* step,
                       const float* other_grad_sq, float max_norm, int kind, float lr, float beta1,
                       float beta2, float eps, float weight_decay, float* clip_coef_out,
                       ce_stream_t stream);

/* Number of kernels this library has launched in this process (for benchmark bookkeeping). */
unsigned long long ce_debug_launch_count(void);

/* Debug / self-test: C[M,N] (fp32) = A * B^t through the same tcgen05 + TMA main loop the
 * similarity GEMM uses.  a_mn_major / b_mn_major select the operand layouts:
 *   K-major : A is [M, K] row-major, B is [N, K] row-major
 *   MN-major: A is [K, M] row-major, B is [K, N] row-major
 * CE_F32 runs the three-product tf32 path with (hi, lo) = (A, A), i.e. returns 3 * A * B^t for
 * tf32-representable inputs. */
int ce_debug_gemm(const void* A, const void* B, float* C, int M, int N, int K, int dtype,
                  int a_mn_major, int b_mn_major, int split_k, ce_stream_t stream);

/* Same contract, CE_BF16 only, through the CTA-pair main loop (`tcgen05.mma.cta_group::2`: two SMs
 * on one 256 x 256 tile, each staging its 128 rows of A and half of the B tile). */
int ce_debug_gemm_pair(const void* A, const void* B, float* C, int M, int N, int K, int dtype,
                       int a_mn_major, int b_mn_major, int split_k, ce_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CLIP_EVENT_B200_H_ */
