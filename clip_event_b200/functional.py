"""Autograd bindings of the CUDA loss-head kernels (thin: argument checks, buffers, ctypes calls)."""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch

from . import _lib as L

OT_LOSS_WEIGHT = 0.01      # model_clip.py:707
IPOT_BETA, IPOT_ITERS, IPOT_K = 0.5, 50, 1   # model_ot.py:68


def _i64(t: torch.Tensor, device) -> torch.Tensor:
    if t.dtype != torch.int64 or t.device != device or not t.is_contiguous():
        t = t.to(device=device, dtype=torch.int64).contiguous()
    return t


def _scalar_f32(t: torch.Tensor, device) -> torch.Tensor:
    return t.detach().to(device=device, dtype=torch.float32).reshape(1).contiguous()


_ONES = {}


def unit_gradient(device) -> torch.Tensor:
    """A device scalar 1.0 (the upstream gradient the eager gradient launches are formed for), made once per device
    -- not while a CUDA graph is being captured, where a first use would tie the constant to that graph's pool."""
    if device.type == "cuda" and torch.cuda.is_current_stream_capturing():
        return torch.ones(1, dtype=torch.float32, device=device)
    t = _ONES.get(device)
    if t is None:
        t = _ONES[device] = torch.ones(1, dtype=torch.float32, device=device)
        if device.type == "cuda":
            torch.cuda.current_stream(device).synchronize()     # once: later calls may come from any stream
    return t


def _upstream(g: Optional[torch.Tensor], device):
    """(tensor kept alive, pointer, dtype code) of an upstream gradient for ce_head_step_scale: read on the device in
    the dtype autograd delivers it in; anything but an fp32 / bf16 device scalar goes through one cast."""
    if g is None:
        return None, 0, L.CE_F32
    g = g.detach()
    if g.device != device or g.dtype not in (torch.float32, torch.bfloat16) or g.numel() != 1:
        g = _scalar_f32(g, device)
    return g, g.data_ptr(), L.dtype_code(g.dtype)


def head_step_scale(contrastive_bufs, ot_bufs, g_i, g_t, g_ot, what: str) -> None:
    """Every stashed gradient buffer of a one-call loss head step times the upstream gradient of its loss, ONE
    launch (``ce_head_step_scale``): a no-op on the device under ``sum(loss_dict.values()).backward()``."""
    import ctypes
    bufs = [(t, 0) for t in contrastive_bufs if t is not None] + [(t, 1) for t in ot_bufs if t is not None]
    if not bufs:
        return
    if any(not t.is_contiguous() for t, _ in bufs):
        raise RuntimeError("clip_event_b200 " + what + ": gradient buffers must be contiguous")
    dev = bufs[0][0].device
    n = len(bufs)
    keep_i, p_i, dt_c = _upstream(g_i, dev)
    keep_t, p_t, dt_t = _upstream(g_t, dev)
    keep_o, p_o, dt_o = _upstream(g_ot, dev)
    if keep_i is not None and keep_t is not None and dt_c != dt_t:
        keep_i, keep_t = _scalar_f32(keep_i, dev), _scalar_f32(keep_t, dev)
        p_i, p_t, dt_c = keep_i.data_ptr(), keep_t.data_ptr(), L.CE_F32
    ptrs = (ctypes.c_void_p * n)(*[t.data_ptr() for t, _ in bufs])
    counts = (ctypes.c_int64 * n)(*[t.numel() for t, _ in bufs])
    dts = (ctypes.c_int * n)(*[L.dtype_code(t.dtype) for t, _ in bufs])
    which = (ctypes.c_int * n)(*[w for _, w in bufs])
    L.check(L.load().ce_head_step_scale(ptrs, counts, dts, which, n, p_i, p_t, dt_c, p_o, dt_o, L.stream_ptr()), what)


def _debug_check_finite(losses: torch.Tensor, what: str) -> None:
    """Index errors come back from the kernels as NaN losses (no host synchronisation on the hot
    path).  With CE_CHECK_INPUTS=1 the losses are read back here and a RuntimeError names the cause,
    which is what the reference's IndexError would have said."""
    if os.environ.get("CE_CHECK_INPUTS") == "1" and bool(torch.isnan(losses).any()):
        raise RuntimeError("clip_event_b200 " + what)


# --------------------------------------------------------------------------------------------
# similarity + InfoNCE (over batch)
# --------------------------------------------------------------------------------------------
IMAGE_LOSS = {"ce_overbatch": L.CE_IMG_CE_OVERBATCH, "ce_instance": L.CE_IMG_CE_INSTANCE,
              "bce_instance": L.CE_IMG_BCE_INSTANCE}


class _ContrastiveOverBatch(torch.autograd.Function):
    """(image_features, text_features, logit_scale) -> (loss_i, loss_t); model_clip.py:496-520,633-662.

    ``image_loss`` selects the image side: 'ce_overbatch' (labels int64 [B] = positive column),
    'ce_instance' (labels int64 [B] in [0,T)) or 'bce_instance' (labels float [B,T]).  The text
    side is always the over-batch cross-entropy of the ``index_pos`` rows.
    """

    @staticmethod
    def forward(ctx, img, txt, logit_scale, labels_i, labels_t, index_pos, image_loss="ce_overbatch"):
        L.require_cuda(img, txt, logit_scale)
        if img.dtype != txt.dtype:
            raise RuntimeError("image_features and text_features must share a dtype")
        dt = L.dtype_code(img.dtype)
        if img.dim() != 2 or txt.dim() != 2 or img.shape[1] != txt.shape[1]:
            raise RuntimeError("expected image_features [B,D] and text_features [B*T,D]")
        dev = img.device
        img_c, txt_c = img.detach().contiguous(), txt.detach().contiguous()
        ls = _scalar_f32(logit_scale, dev)
        mode = IMAGE_LOSS[image_loss]
        labels_t, index_pos = _i64(labels_t, dev), _i64(index_pos, dev)
        B, D = img_c.shape
        BT, P = txt_c.shape[0], index_pos.numel()
        # the reference index_selects labels_per_text with index_pos (model_clip.py:656): a shorter
        # vector is an IndexError there, and an out-of-bounds device read here -- refuse it
        if labels_t.numel() != BT:
            raise RuntimeError("labels_per_text must have one entry per description (%d), got %d"
                               % (BT, labels_t.numel()))
        if P < 1:
            raise RuntimeError("index_pos is empty: the text-side loss has no rows (model_clip.py:655-659)")
        if mode == L.CE_IMG_BCE_INSTANCE:
            labels_i = labels_i.to(device=dev, dtype=torch.float32).contiguous()
            if B == 0 or BT % B or labels_i.numel() != BT:
                raise RuntimeError("'bce' labels_per_image must be [B, T] with B*T descriptions")
        else:
            labels_i = _i64(labels_i, dev)
            if labels_i.numel() != B:
                raise RuntimeError("labels_per_image must have one entry per image")
            if mode == L.CE_IMG_CE_INSTANCE and (B == 0 or BT % B):
                raise RuntimeError("over-instance logits need B*T descriptions")
        lib = L.load()
        nbytes = lib.ce_contrastive_workspace_bytes(B, BT, P, D, dt)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        out = torch.empty(2, dtype=torch.float32, device=dev)
        L.check(lib.ce_contrastive_fwd(L.ptr(img_c), L.ptr(txt_c), L.ptr(ls), L.ptr(labels_i), L.ptr(labels_t),
                                       L.ptr(index_pos), B, BT, P, D, mode, dt, out.data_ptr(), out.data_ptr() + 4,
                                       ws.data_ptr(), nbytes, L.stream_ptr()), "contrastive forward")
        _debug_check_finite(out, "contrastive forward: an index in labels_per_image / labels_per_text / index_pos "
                                 "is out of range or index_pos lists a description twice")
        ctx.save_for_backward(img_c, txt_c, ls, labels_i, labels_t, index_pos, ws)
        ctx.dims = (B, BT, P, D, dt, mode)
        ctx.ls_dtype = logit_scale.dtype
        ctx.ls_shape = logit_scale.shape
        return out[0], out[1]

    @staticmethod
    def backward(ctx, g_i, g_t):
        img, txt, ls, labels_i, labels_t, index_pos, ws = ctx.saved_tensors
        B, BT, P, D, dt, mode = ctx.dims
        dev = img.device
        # a missing upstream gradient is a zero; no fill launch when both are present
        gi = torch.zeros(1, dtype=torch.float32, device=dev) if g_i is None else _scalar_f32(g_i, dev)
        gt = torch.zeros(1, dtype=torch.float32, device=dev) if g_t is None else _scalar_f32(g_t, dev)
        dimg, dtxt = torch.empty_like(img), torch.empty_like(txt)
        dls = torch.empty(1, dtype=torch.float32, device=dev)
        lib = L.load()
        L.check(lib.ce_contrastive_bwd(L.ptr(img), L.ptr(txt), L.ptr(ls), L.ptr(labels_i), L.ptr(labels_t),
                                       L.ptr(index_pos), B, BT, P, D, mode, dt, L.ptr(gi), L.ptr(gt), L.ptr(dimg),
                                       L.ptr(dtxt), L.ptr(dls), ws.data_ptr(), ws.numel(), L.stream_ptr()),
                "contrastive backward")
        return dimg, dtxt, dls.reshape(ctx.ls_shape).to(ctx.ls_dtype), None, None, None, None


def contrastive_over_batch(image_features, text_features, logit_scale, labels_per_image,
                           labels_per_text, index_pos) -> Tuple[torch.Tensor, torch.Tensor]:
    """loss_i, loss_t of ``CriterionContrastive('ce')`` applied to ``CLIP.forward``'s over-batch logits."""
    return _ContrastiveOverBatch.apply(image_features, text_features, logit_scale, labels_per_image,
                                       labels_per_text, index_pos, "ce_overbatch")


def contrastive_over_instance(image_features, text_features, logit_scale, labels_per_image,
                              labels_per_text, index_pos, loss="ce") -> Tuple[torch.Tensor, torch.Tensor]:
    """loss_i over each image's own T descriptions ('ce' or 'bce', model_clip.py:509-520,624-651) and
    the over-batch text-side loss_t."""
    if loss not in ("ce", "bce"):
        raise RuntimeError("Invalid constrastive_loss '{}'. ".format(loss))
    return _ContrastiveOverBatch.apply(image_features, text_features, logit_scale, labels_per_image,
                                       labels_per_text, index_pos, loss + "_instance")


class _DenseCrossEntropy(torch.autograd.Function):
    """mean cross-entropy (or BCE with logits) over the rows of a MATERIALISED logits matrix,
    optionally over ``logits.index_select(0, row_index)`` -- model_clip.py:648-659 on plain tensors."""

    @staticmethod
    def forward(ctx, logits, labels, row_index, labels_by_row, kind):
        L.require_cuda(logits, labels)
        if logits.dim() != 2:
            raise RuntimeError("expected a [rows, classes] logits matrix")
        dt = L.dtype_code(logits.dtype)
        dev = logits.device
        lg = logits.detach().contiguous()
        rows, cols = lg.shape
        if kind == 1:
            lab = labels.to(device=dev, dtype=torch.float32).contiguous()
            if lab.shape != lg.shape:
                raise RuntimeError("'bce' labels_per_image must have the shape of logits_per_image")
        else:
            lab = _i64(labels, dev)
        ri = None if row_index is None else _i64(row_index, dev)
        n = rows if ri is None else ri.numel()
        if kind == 0 and lab.numel() != (rows if (labels_by_row or ri is None) else n):
            raise RuntimeError("labels must have one entry per logits row")
        if n < 1:
            raise RuntimeError("no rows to average over")
        lib = L.load()
        nbytes = lib.ce_dense_ce_workspace_bytes(n)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        L.check(lib.ce_dense_ce_fwd(lg.data_ptr(), cols, rows, cols, L.ptr(ri), n, lab.data_ptr(), int(labels_by_row),
                                    int(kind), dt, loss.data_ptr(), ws.data_ptr(), nbytes, L.stream_ptr()), "dense CE forward")
        _debug_check_finite(loss, "dense cross-entropy: a label or index_pos entry is out of range")
        ctx.save_for_backward(lg, lab, ri if ri is not None else torch.empty(0, device=dev), ws)
        ctx.meta = (rows, cols, n, ri is not None, int(labels_by_row), int(kind), dt)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        lg, lab, ri, ws = ctx.saved_tensors
        rows, cols, n, has_ri, by_row, kind, dt = ctx.meta
        gg = _scalar_f32(g, lg.device)
        d = torch.empty(rows, cols, dtype=torch.float32, device=lg.device)
        L.check(L.load().ce_dense_ce_bwd(lg.data_ptr(), cols, rows, cols, ri.data_ptr() if has_ri else 0, n, lab.data_ptr(),
                                         by_row, kind, dt, gg.data_ptr(), d.data_ptr(), cols, ws.data_ptr(), L.stream_ptr()),
                "dense CE backward")
        return d.to(lg.dtype), None, None, None, None


def dense_contrastive(logits_per_image, logits_per_text, labels_per_image, labels_per_text, index_pos, loss="ce"):
    """``CriterionContrastive.forward`` on materialised logits (any producer): loss_i over the rows of
    ``logits_per_image`` ('ce' or 'bce'), loss_t over ``logits_per_text[index_pos]`` with
    ``labels_per_text[index_pos]`` (model_clip.py:648-659)."""
    if loss not in ("ce", "bce"):
        raise RuntimeError("Invalid constrastive_loss '{}'. ".format(loss))
    loss_i = _DenseCrossEntropy.apply(logits_per_image, labels_per_image, None, False, 1 if loss == "bce" else 0)
    loss_t = _DenseCrossEntropy.apply(logits_per_text, labels_per_text, index_pos, True, 0)
    return loss_i, loss_t


def similarity_logits(a: torch.Tensor, b: torch.Tensor, logit_scale: torch.Tensor) -> torch.Tensor:
    """Dense ``exp(logit_scale) * normalize(a) @ normalize(b).T`` (fp32), no autograd."""
    L.require_cuda(a, b, logit_scale)
    dt = L.dtype_code(a.dtype)
    if a.dtype != b.dtype or a.dim() != 2 or b.dim() != 2 or a.shape[1] != b.shape[1]:
        raise RuntimeError("similarity_logits expects two [rows, D] matrices of one dtype")
    a_c, b_c = a.detach().contiguous(), b.detach().contiguous()
    ls = _scalar_f32(logit_scale, a.device)
    Ra, D = a_c.shape
    Rb = b_c.shape[0]
    lib = L.load()
    nbytes = lib.ce_similarity_workspace_bytes(Ra, Rb, D, dt)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=a.device)
    out = torch.empty(Ra, Rb, dtype=torch.float32, device=a.device)
    L.check(lib.ce_similarity_logits(L.ptr(a_c), L.ptr(b_c), L.ptr(ls), Ra, Rb, D, dt, out.data_ptr(),
                                     ws.data_ptr(), nbytes, L.stream_ptr()), "similarity_logits")
    return out


# --------------------------------------------------------------------------------------------
# OT alignment
# --------------------------------------------------------------------------------------------
def _mask_args(mask: torch.Tensor):
    """(tensor kept alive, mask_kind): int64 *_num arrays or bool/uint8 *_pad arrays."""
    if mask.dtype == torch.int64:
        return mask, L.CE_MASK_NUM_I64
    if mask.dtype == torch.bool:
        return mask.view(torch.uint8), L.CE_MASK_PAD_U8
    if mask.dtype == torch.uint8:
        return mask, L.CE_MASK_PAD_U8
    return mask.to(torch.int64), L.CE_MASK_NUM_I64


def num_mask(mask: torch.Tensor) -> torch.Tensor:
    """``*_num`` arrays of CriterionAlignment (model_clip.py:679-690): nonzero = valid node, whatever
    the dtype (the reference applies ``mask2pad(x) = (x == 0)``); the kernels read them as int64."""
    return mask if mask.dtype == torch.int64 else mask.to(torch.int64)


class _OtAlignment(torch.autograd.Function):
    """(txt_nodes [B,M,D], img_nodes view [B,N,D]) -> (loss, dist[B]).

    The forward launch also produces loss_scale * d(sum dist)/d(inputs) (IPOT is not
    differentiated through, model_ot.py:32,81), so the backward is a scale by the incoming
    gradient -- a no-op launch when that is 1, as it is under ``sum(loss_dict.values()).backward()``.
    A second backward over a retained graph launches the kernel again on the saved inputs.
    """

    @staticmethod
    def forward(ctx, txt, obj, txt_mask, obj_mask, drop_slot0, beta, iters, k, loss_scale):
        L.require_cuda(txt, obj, txt_mask, obj_mask)
        if txt.dtype != obj.dtype:
            raise RuntimeError("text and image node embeddings must share a dtype")
        dt = L.dtype_code(txt.dtype)
        if txt.dim() != 3 or obj.dim() != 3 or txt.shape[0] != obj.shape[0] or txt.shape[2] != obj.shape[2]:
            raise RuntimeError("expected [B,M,D] text nodes and [B,N(+1),D] image nodes")
        txt_c, obj_c = txt.detach().contiguous(), obj.detach().contiguous()
        tm, kind_t = _mask_args(txt_mask.contiguous())
        om, kind_o = _mask_args(obj_mask.contiguous())
        if kind_t != kind_o:
            raise RuntimeError("text and image node masks must use the same encoding")
        B, M, D = txt_c.shape
        slot = 1 if drop_slot0 else 0
        N = obj_c.shape[1] - slot
        if tm.shape != (B, M) or om.shape != (B, N + slot):
            raise RuntimeError("node masks must be [B,M] and [B,N(+1)]")
        need_grad = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        # both gradients live in one allocation, so the backward is ONE scale launch over it
        loss, dist, dtxt, dobj, gbuf, _ = _ot_launch(txt_c, obj_c, tm, om, kind_t, drop_slot0, beta, iters, k, loss_scale,
                                                     need_grad, L.stream_ptr())
        # what a second backward (retain_graph=True, as the reference's autograd allows) needs to form the gradients again
        ctx.replay = (txt_c, obj_c, tm, om, kind_t, drop_slot0, beta, iters, k, loss_scale) if need_grad else None
        ctx.stash = (dtxt, dobj, gbuf)
        ctx.loss_scale = float(loss_scale)
        ctx.consumed = False
        ctx.set_materialize_grads(False)   # an unused output arrives as None, not as zeros
        return loss[0], dist

    @staticmethod
    def backward(ctx, g_loss, g_dist):
        if ctx.replay is None:
            return (None,) * 9
        if ctx.consumed:
            # the stashed gradients went to autograd with the first backward (it may own or have changed them):
            # launch the kernel again on the saved inputs
            _, _, dtxt, dobj, gbuf, _ = _ot_launch(*ctx.replay, True, L.stream_ptr())
        else:
            dtxt, dobj, gbuf = ctx.stash
            ctx.consumed = True
            ctx.stash = (None, None, None)
        dev = dtxt.device
        if g_dist is not None:
            # per-sample upstream gradients (optimal_transport_dist users): general path
            per = g_dist.to(torch.float32) / ctx.loss_scale
            if g_loss is not None:
                per = per + g_loss.to(torch.float32)
            per = per.view(-1, 1, 1)
            return (dtxt.float() * per).to(dtxt.dtype), (dobj.float() * per).to(dobj.dtype), None, None, None, None, None, None, None
        g = _scalar_f32(g_loss, dev)
        lib = L.load()
        dt = L.dtype_code(dtxt.dtype)
        L.check(lib.ce_scale_inplace(gbuf.data_ptr(), 1, gbuf.numel(), gbuf.numel(), dt, g.data_ptr(), L.stream_ptr()),
                "OT backward scale")
        return dtxt, dobj, None, None, None, None, None, None, None


def ot_alignment(txt_nodes, object_vec, txt_mask, object_mask, drop_slot0=True, beta=IPOT_BETA,
                 iters=IPOT_ITERS, k=IPOT_K, loss_scale=OT_LOSS_WEIGHT):
    """(loss_scale * sum_b dist[b], dist[B]) for text nodes vs image nodes (slot 0 dropped if asked)."""
    return _OtAlignment.apply(txt_nodes, object_vec, txt_mask, object_mask, drop_slot0, beta, iters, k, loss_scale)


# --------------------------------------------------------------------------------------------
# the projections that feed the head (model_clip.py:253-260 and :412-415)
# --------------------------------------------------------------------------------------------
class _ProjectionTail(torch.autograd.Function):
    """hidden [rows, L, W] -> LayerNorm(hidden[arange, token]) @ proj  ([rows, D]); gradients for the hidden
    states (non-zero at the selected token only), the LayerNorm vectors and the projection."""

    @staticmethod
    def forward(ctx, hidden, token_index, ln_w, ln_b, proj, eps):
        L.require_cuda(hidden, proj)
        if hidden.dim() != 3 or proj.dim() != 2 or hidden.shape[2] != proj.shape[0]:
            raise RuntimeError("expected hidden [rows, L, W] and proj [W, D]")
        if proj.dtype != hidden.dtype or (ln_w is not None and (ln_w.dtype != hidden.dtype or ln_b.dtype != hidden.dtype)):
            raise RuntimeError("hidden states, LayerNorm vectors and the projection must share a dtype")
        dt = L.dtype_code(hidden.dtype)
        h = hidden.detach().contiguous()
        pj = proj.detach().contiguous()
        rows, Lt, W = h.shape
        D = pj.shape[1]
        dev = h.device
        tok = None if token_index is None else token_index.to(device=dev, dtype=torch.int64).contiguous()
        lw = None if ln_w is None else ln_w.detach().contiguous()
        lb = None if ln_b is None else ln_b.detach().contiguous()
        lib = L.load()
        nbytes = lib.ce_proj_workspace_bytes(rows, W, D, dt)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        feat = torch.empty(rows, D, dtype=h.dtype, device=dev)
        norm2 = torch.empty(rows, dtype=torch.float32, device=dev)
        L.check(lib.ce_proj_fwd(h.data_ptr(), Lt * W, L.ptr(tok), L.ptr(lw), L.ptr(lb), float(eps), pj.data_ptr(), rows, W, D, dt,
                                feat.data_ptr(), norm2.data_ptr(), ws.data_ptr(), nbytes, L.stream_ptr()), "projection forward")
        ctx.save_for_backward(h, tok, lw, pj, ws)
        ctx.has_ln = lw is not None
        ctx.mark_non_differentiable(norm2)
        return feat, norm2

    @staticmethod
    def backward(ctx, dfeat, _):
        h, tok, lw, pj, ws = ctx.saved_tensors
        rows, Lt, W = h.shape
        D = pj.shape[1]
        dev = h.device
        dt = L.dtype_code(h.dtype)
        df = dfeat.detach().to(h.dtype).contiguous()
        dx_rows = torch.empty(rows, W, dtype=h.dtype, device=dev)
        dlw = torch.empty(W, dtype=torch.float32, device=dev)
        dlb = torch.empty(W, dtype=torch.float32, device=dev)
        dproj = torch.empty(W, D, dtype=torch.float32, device=dev)
        L.check(L.load().ce_proj_bwd(h.data_ptr(), Lt * W, L.ptr(tok), L.ptr(lw), pj.data_ptr(), df.data_ptr(), rows, W, D, dt,
                                     dx_rows.data_ptr(), dlw.data_ptr(), dlb.data_ptr(), dproj.data_ptr(), ws.data_ptr(),
                                     ws.numel(), L.stream_ptr()), "projection backward")
        dh = None
        if ctx.needs_input_grad[0]:
            dh = torch.zeros_like(h)
            ar = torch.arange(rows, device=dev)
            dh[ar, tok if tok is not None else torch.zeros_like(ar)] = dx_rows
        return (dh, None, dlw.to(h.dtype) if ctx.has_ln else None, dlb.to(h.dtype) if ctx.has_ln else None,
                dproj.to(h.dtype), None)


def projection_tail(hidden, proj, ln_weight=None, ln_bias=None, token_index=None, eps=1e-5):
    """(features [rows, D], squared L2 norms [rows]) = LayerNorm(hidden[arange, token_index]) @ proj.
    ``token_index`` None = token 0 (the class token, model_clip.py:256); pass ``text.argmax(dim=-1)`` for
    the text side (model_clip.py:415)."""
    return _ProjectionTail.apply(hidden, token_index, ln_weight, ln_bias, proj, eps)


# --------------------------------------------------------------------------------------------
# packed (variable-length) node sets: SURVEY.md 8f-3
# --------------------------------------------------------------------------------------------
class PackedNodes:
    """Node embeddings of a batch without the padding: ``rows`` [sum_b count[b], D] (sample after sample, each
    sample's valid nodes in their original order) and ``offsets`` [B+1] int32 (``rows[offsets[b]:offsets[b+1]]``
    belongs to sample b).  ``max_count`` is a Python int so that no launch has to wait for the device."""

    def __init__(self, rows, offsets, max_count):
        self.rows, self.offsets, self.max_count = rows, offsets, int(max_count)

    @property
    def batch(self):
        return self.offsets.numel() - 1

    def to_padded(self, width=None):
        """[B, width, D] zero-padded copy and its [B, width] int64 validity mask (the reference's layout)."""
        B, D = self.batch, self.rows.shape[1]
        width = self.max_count if width is None else width
        counts = (self.offsets[1:] - self.offsets[:-1]).to(torch.int64)
        pos = torch.arange(width, device=self.rows.device).unsqueeze(0)
        mask = pos < counts.unsqueeze(1)
        out = self.rows.new_zeros(B, width, D)
        out[mask] = self.rows
        return out, mask.to(torch.int64)


def pack_nodes(vec, num, drop_first=False):
    """Padded [B, S, D] embeddings + [B, S] validity mask (``*_num`` semantics: nonzero = valid, any pattern)
    -> :class:`PackedNodes`.  ``drop_first`` removes slot 0 first (the whole-image slot, model_clip.py:686).
    This is the converter for callers that still produce the reference's padded layout
    (model_clip.py:531-552, dataset_voa.py:532-544,566-577); a packing-aware encoder would emit the rows
    directly.  One host read (the maximum count) -- do it in the data loader, not in the step."""
    if drop_first:
        vec, num = vec[:, 1:], num[:, 1:]
    valid = num != 0
    counts = valid.sum(dim=1)
    offsets = torch.zeros(vec.shape[0] + 1, dtype=torch.int32, device=vec.device)
    offsets[1:] = counts.cumsum(0)
    rows = vec[valid]            # boolean-mask gather keeps sample order and node order; autograd-connected
    return PackedNodes(rows.contiguous(), offsets, int(counts.max().item()) if vec.shape[0] else 0)


class _OtAlignmentPacked(torch.autograd.Function):
    """(text rows, image rows) in packed layout -> (loss, dist[B]); gradients come back packed."""

    @staticmethod
    def forward(ctx, txt_rows, img_rows, txt_off, img_off, max_m, max_n, beta, iters, k, loss_scale):
        L.require_cuda(txt_rows, img_rows, txt_off, img_off)
        if txt_rows.dtype != img_rows.dtype or txt_rows.dim() != 2 or img_rows.dim() != 2 or txt_rows.shape[1] != img_rows.shape[1]:
            raise RuntimeError("expected [sum_m, D] text rows and [sum_n, D] image rows of one dtype")
        if txt_off.dtype != torch.int32 or img_off.dtype != torch.int32 or txt_off.numel() != img_off.numel():
            raise RuntimeError("offsets must be int32 [B+1] vectors")
        B, D = txt_off.numel() - 1, txt_rows.shape[1]
        dev = txt_rows.device
        tr, ir = txt_rows.detach().contiguous(), img_rows.detach().contiguous()
        need_grad = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        dist = torch.empty(B, dtype=torch.float32, device=dev)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        if need_grad:
            n_t = tr.numel()
            n_t_pad = (n_t + 7) // 8 * 8
            gbuf = torch.empty(n_t_pad + ir.numel(), dtype=tr.dtype, device=dev)
            if n_t_pad != n_t:
                gbuf[n_t:n_t_pad].zero_()
            dtxt, dimg = gbuf[:n_t].view_as(tr), gbuf[n_t_pad:].view_as(ir)
        else:
            gbuf = dtxt = dimg = None
        L.check(L.load().ce_ot_fwd_bwd_packed(
            tr.data_ptr(), txt_off.data_ptr(), ir.data_ptr(), img_off.data_ptr(), B, int(max_m), int(max_n), D,
            L.dtype_code(tr.dtype), float(beta), int(iters), int(k), float(loss_scale), dist.data_ptr(), loss.data_ptr(),
            L.ptr(dtxt), L.ptr(dimg), L.stream_ptr()), "packed OT forward")
        ctx.stash = (dtxt, dimg, gbuf)
        ctx.loss_scale = float(loss_scale)
        ctx.set_materialize_grads(False)
        return loss[0], dist

    @staticmethod
    def backward(ctx, g_loss, g_dist):
        dtxt, dimg, gbuf = ctx.stash
        if dtxt is None:
            return (None,) * 10
        if g_dist is not None:
            raise RuntimeError("packed OT: per-sample upstream gradients are not supported; use the padded optimal_transport_dist")
        ctx.stash = (None, None, None)
        g = _scalar_f32(g_loss, dtxt.device)
        L.check(L.load().ce_scale_inplace(gbuf.data_ptr(), 1, gbuf.numel(), gbuf.numel(), L.dtype_code(dtxt.dtype),
                                          g.data_ptr(), L.stream_ptr()), "packed OT backward scale")
        return dtxt, dimg, None, None, None, None, None, None, None, None


def ot_alignment_packed(txt: PackedNodes, img: PackedNodes, beta=IPOT_BETA, iters=IPOT_ITERS, k=IPOT_K,
                        loss_scale=OT_LOSS_WEIGHT):
    """(loss_scale * sum_b dist[b], dist[B]) for packed node sets (bf16, at most 16 text and 64 image nodes per
    sample, D a multiple of 64 up to 512 -- the streaming kernel; anything else is padded on the device and
    goes through :func:`ot_alignment`).  Equal to the padded, masked call."""
    D = txt.rows.shape[1]
    ok = (txt.rows.dtype == torch.bfloat16 and 1 <= txt.max_count <= 16 and 1 <= img.max_count <= 64
          and D % 64 == 0 and D <= 512)
    if ok:
        return _OtAlignmentPacked.apply(txt.rows, img.rows, txt.offsets, img.offsets, txt.max_count, img.max_count,
                                        beta, iters, k, loss_scale)
    tp, tm = txt.to_padded(max(txt.max_count, 1))
    ip, im = img.to_padded(max(img.max_count, 1))
    return ot_alignment(tp, ip, tm, im, drop_slot0=False, beta=beta, iters=iters, k=k, loss_scale=loss_scale)


# --------------------------------------------------------------------------------------------
# one call for engine.py:48-67,88: both criteria, two streams, gradients formed with the losses
# --------------------------------------------------------------------------------------------
_SIDE_STREAMS = {}


def side_stream(device) -> "torch.cuda.Stream":
    """The library's second stream on ``device`` (the OT chain runs there, next to the GEMM chain)."""
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    st = _SIDE_STREAMS.get(key)
    if st is None:
        st = _SIDE_STREAMS[key] = torch.cuda.Stream(device=key)
    return st


def _ot_launch(txt_c, obj_c, tm, om, kind, drop_slot0, beta, iters, k, loss_scale, need_grad, stream_ptr):
    """Allocate the OT buffers (on the CURRENT stream's pool) and enqueue ce_ot_fwd_bwd on ``stream_ptr``."""
    B, M, D = txt_c.shape
    slot = 1 if drop_slot0 else 0
    N = obj_c.shape[1] - slot
    esz, msz, dev = txt_c.element_size(), om.element_size(), txt_c.device
    lib = L.load()
    nbytes = lib.ce_ot_workspace_bytes(B, M, N, D)
    if nbytes == 0:
        raise RuntimeError("clip_event_b200 OT: unsupported node counts M=%d N=%d" % (M, N))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    dist = torch.empty(B, dtype=torch.float32, device=dev)
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    if need_grad:
        n_t = txt_c.numel()
        n_t_pad = (n_t + 7) // 8 * 8
        gbuf = torch.empty(n_t_pad + obj_c.numel(), dtype=txt_c.dtype, device=dev)
        if n_t_pad != n_t:
            gbuf[n_t:n_t_pad].zero_()
        dtxt, dobj = gbuf[:n_t].view_as(txt_c), gbuf[n_t_pad:].view_as(obj_c)
    else:
        gbuf = dtxt = dobj = None
    L.check(lib.ce_ot_fwd_bwd(
        txt_c.data_ptr(), M * D, obj_c.data_ptr() + slot * D * esz, (N + slot) * D,
        tm.data_ptr(), M, om.data_ptr() + slot * msz, N + slot, kind, B, M, N, D, L.dtype_code(txt_c.dtype),
        float(beta), int(iters), int(k), float(loss_scale), dist.data_ptr(), loss.data_ptr(),
        L.ptr(dtxt), 0 if dobj is None else dobj.data_ptr() + slot * D * esz,
        0 if (dobj is None or not slot) else dobj.data_ptr(), ws.data_ptr(), nbytes, stream_ptr), "OT forward")
    return loss, dist, dtxt, dobj, gbuf, ws


class _LossHeadStep(torch.autograd.Function):
    """(image_features, text_features, logit_scale, entitytxt_vec, object_vec) -> (loss_i, loss_t, loss_ot).

    engine.py:48-67 computes the two criteria back to back and line 88 back-propagates their plain
    sum.  Neither gradient depends on the upstream gradient except as a scale, so this call forms
    losses AND gradients at once: the similarity/InfoNCE chain (tensor-core bound) on the current
    stream, the OT chain (HBM/ALU bound) on the library's side stream, joined before it returns.
    ``backward`` is two scale launches that return immediately on the device when the upstream
    gradients are 1.  loss_i and loss_t must receive the SAME upstream gradient (they do under
    ``sum(loss_dict.values())``); anything else turns the contrastive gradients into NaN rather than
    passing silently -- use the separate criteria for weighted sums.
    """

    @staticmethod
    def forward(ctx, img, txt, logit_scale, etxt, obj, labels_i, labels_t, index_pos, tnum, onum, image_loss,
                cast_losses=False):
        L.require_cuda(img, txt, logit_scale, etxt, obj, tnum, onum)
        if img.dtype != txt.dtype or etxt.dtype != obj.dtype:
            raise RuntimeError("features must share a dtype")
        dev = img.device
        dt = L.dtype_code(img.dtype)
        img_c, txt_c = img.detach().contiguous(), txt.detach().contiguous()
        etxt_c, obj_c = etxt.detach().contiguous(), obj.detach().contiguous()
        ls = _scalar_f32(logit_scale, dev)
        mode = IMAGE_LOSS[image_loss]
        labels_t, index_pos = _i64(labels_t, dev), _i64(index_pos, dev)
        labels_i = labels_i.to(device=dev, dtype=torch.float32).contiguous() if mode == L.CE_IMG_BCE_INSTANCE else _i64(labels_i, dev)
        B, D = img_c.shape
        BT, P = txt_c.shape[0], index_pos.numel()
        if labels_t.numel() != BT or P < 1:
            raise RuntimeError("labels_per_text must have one entry per description and index_pos at least one")
        tm, kind_t = _mask_args(num_mask(tnum).contiguous())
        om, kind_o = _mask_args(num_mask(onum).contiguous())
        need_c = any(ctx.needs_input_grad[:3])
        need_o = any(ctx.needs_input_grad[3:5])
        lib = L.load()
        cur = torch.cuda.current_stream()
        side = side_stream(dev)
        # OT chain on the side stream (buffers come from the current stream's pool; the join below orders every later use)
        side.wait_stream(cur)
        loss_ot, dist, detxt, dobj, gbuf, ws_ot = _ot_launch(etxt_c, obj_c, tm, om, kind_t, True, IPOT_BETA, IPOT_ITERS,
                                                             IPOT_K, OT_LOSS_WEIGHT, need_o, side.cuda_stream)
        # similarity + InfoNCE chain on the current stream
        nbytes = lib.ce_contrastive_workspace_bytes(B, BT, P, D, dt)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        out = torch.empty(3, dtype=torch.float32, device=dev)
        sp = cur.cuda_stream
        L.check(lib.ce_contrastive_fwd(L.ptr(img_c), L.ptr(txt_c), L.ptr(ls), L.ptr(labels_i), L.ptr(labels_t),
                                       L.ptr(index_pos), B, BT, P, D, mode, dt, out.data_ptr(), out.data_ptr() + 4,
                                       ws.data_ptr(), nbytes, sp), "contrastive forward")
        if need_c:
            one = unit_gradient(dev)
            # dimg | dtxt in one allocation (one scale launch in the backward), dls apart (fp32)
            n_i = img_c.numel()
            n_i_pad = (n_i + 7) // 8 * 8
            cbuf = torch.empty(n_i_pad + txt_c.numel(), dtype=img_c.dtype, device=dev)
            if n_i_pad != n_i:
                cbuf[n_i:n_i_pad].zero_()
            dimg, dtxt = cbuf[:n_i].view_as(img_c), cbuf[n_i_pad:].view_as(txt_c)
            dls = torch.empty(1, dtype=torch.float32, device=dev)
            L.check(lib.ce_contrastive_bwd(L.ptr(img_c), L.ptr(txt_c), L.ptr(ls), L.ptr(labels_i), L.ptr(labels_t),
                                           L.ptr(index_pos), B, BT, P, D, mode, dt, L.ptr(one), L.ptr(one), L.ptr(dimg),
                                           L.ptr(dtxt), L.ptr(dls), ws.data_ptr(), nbytes, sp), "contrastive backward")
        else:
            cbuf = dimg = dtxt = dls = None
        cur.wait_stream(side)           # join: everything below and after sees both chains
        ctx.stash = (cbuf, dimg, dtxt, dls, gbuf, detxt, dobj)
        ctx.ls_meta = (logit_scale.dtype, logit_scale.shape)
        ctx.set_materialize_grads(False)
        if cast_losses:
            # the dtypes the reference's criteria return: the logits' for loss_i / loss_t, the nodes' for loss_ot
            lc = torch.empty(2, dtype=img_c.dtype, device=dev)
            lo = torch.empty(1, dtype=etxt_c.dtype, device=dev)
            L.check(lib.ce_head_losses_cast(out.data_ptr(), out.data_ptr() + 4, loss_ot.data_ptr(), lc.data_ptr(), dt,
                                            lo.data_ptr(), L.dtype_code(etxt_c.dtype), sp), "loss head step losses")
            return lc[0], lc[1], lo[0]
        out[2:3].copy_(loss_ot)
        return out[0], out[1], out[2]

    @staticmethod
    def backward(ctx, g_i, g_t, g_ot):
        if getattr(ctx, "consumed", False):
            raise RuntimeError("clip_event_b200 loss head step: backward a second time; run the step again "
                               "(retain_graph is not supported on this path)")
        cbuf, dimg, dtxt, dls, gbuf, detxt, dobj = ctx.stash
        if cbuf is None and gbuf is None:
            return (None,) * 12
        ctx.consumed = True
        ctx.stash = (None,) * 7
        out = [None] * 12
        use_c = cbuf is not None and (g_i is not None or g_t is not None)
        use_o = gbuf is not None and g_ot is not None
        if use_c and (g_i is None or g_t is None):
            raise RuntimeError("clip_event_b200 loss head step: loss_i and loss_t must be back-propagated together "
                               "(use CriterionContrastive for a single-sided or weighted loss)")
        head_step_scale((cbuf, dls) if use_c else (), (gbuf,) if use_o else (), g_i if use_c else None,
                        g_t if use_c else None, g_ot if use_o else None, "loss head step backward")
        if use_c:
            ls_dtype, ls_shape = ctx.ls_meta
            out[0], out[1], out[2] = dimg, dtxt, dls.reshape(ls_shape).to(ls_dtype)
        if use_o:
            out[3], out[4] = detxt, dobj
        return tuple(out)


def loss_head_step(image_features, text_features, logit_scale, labels_per_image, labels_per_text, index_pos,
                   entitytxt_vec, object_vec, entitytxt_num, object_num, image_loss="ce_overbatch", cast_losses=False):
    """(loss_i, loss_t, loss_ot) of the whole loss head in one call; see :class:`_LossHeadStep`.  fp32 losses, or
    (``cast_losses``) the dtypes the reference's criteria return them in, cast inside the call."""
    return _LossHeadStep.apply(image_features, text_features, logit_scale, entitytxt_vec, object_vec,
                               labels_per_image, labels_per_text, index_pos, entitytxt_num, object_num, image_loss,
                               cast_losses)


# --------------------------------------------------------------------------------------------
# engine.py:89-90 for the head's own parameter
# --------------------------------------------------------------------------------------------
class HeadParamStep:
    """``clip_grad_norm_(params, max_norm)`` + ``optimizer.step()`` for ``logit_scale`` in ONE launch.

    ``kind``: 'sgd' (torch.optim.SGD, momentum) or 'adam' (torch.optim.Adam), the two optimisers
    engine.build_optimizer offers (engine.py:133-149).  ``other_grad_sq`` is the squared gradient norm
    of all other parameters (a device scalar; None = the head's parameter is clipped alone).  Returns
    the clip coefficient (device scalar) for the caller's other parameters.
    """

    def __init__(self, param: torch.Tensor, kind="adam", lr=1e-6, momentum=0.9, betas=(0.9, 0.999), eps=1e-8,
                 weight_decay=0.0, max_norm=1.0):
        L.require_cuda(param)
        if param.numel() != 1 or param.dtype != torch.float32:
            raise RuntimeError("HeadParamStep drives the scalar fp32 logit_scale parameter")
        if kind not in ("sgd", "adam"):
            raise RuntimeError("Invalid optimizer '{}'. ".format(kind))      # engine.py:149
        self.param, self.kind = param, kind
        self.lr, self.eps, self.wd, self.max_norm = lr, eps, weight_decay, max_norm
        self.b1, self.b2 = (momentum, 0.0) if kind == "sgd" else betas
        self.state = torch.zeros(4, dtype=torch.float32, device=param.device)    # state0, state1, step, clip_coef

    @torch.no_grad()
    def step(self, other_grad_sq: Optional[torch.Tensor] = None, lr: Optional[float] = None):
        g = self.param.grad
        if g is None:
            raise RuntimeError("logit_scale has no gradient")
        st = self.state
        o = None if other_grad_sq is None else _scalar_f32(other_grad_sq, st.device)
        L.check(L.load().ce_head_param_step(
            self.param.data_ptr(), g.data_ptr(), st.data_ptr(), st.data_ptr() + 4, st.data_ptr() + 8, L.ptr(o),
            float(self.max_norm), 0 if self.kind == "sgd" else 1, float(self.lr if lr is None else lr), float(self.b1),
            float(self.b2), float(self.eps), float(self.wd), st.data_ptr() + 12, L.stream_ptr()), "head parameter step")
        return st[3]
