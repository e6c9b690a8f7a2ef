"""Autograd bindings of the CUDA loss-head kernels (thin: argument checks, buffers, ctypes calls)."""
from __future__ import annotations

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


class _OtAlignment(torch.autograd.Function):
    """(txt_nodes [B,M,D], img_nodes view [B,N,D]) -> (loss, dist[B]).

    The forward launch also produces loss_scale * d(sum dist)/d(inputs) (IPOT is not
    differentiated through, model_ot.py:32,81), so the backward is a scale by the incoming
    gradient -- a no-op launch when that is 1, as it is under ``sum(loss_dict.values()).backward()``.
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
        esz = txt_c.element_size()
        msz = om.element_size()
        need_grad = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        dev = txt_c.device
        lib = L.load()
        nbytes = lib.ce_ot_workspace_bytes(B, M, N, D)
        if nbytes == 0:
            raise RuntimeError("clip_event_b200 OT: unsupported node counts M=%d N=%d" % (M, N))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        dist = torch.empty(B, dtype=torch.float32, device=dev)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        # both gradients live in one allocation, so the backward is ONE scale launch over it
        if need_grad:
            n_t = txt_c.numel()
            n_t_pad = (n_t + 7) // 8 * 8                     # keeps dobj 16-byte aligned
            gbuf = torch.empty(n_t_pad + obj_c.numel(), dtype=txt_c.dtype, device=dev)
            if n_t_pad != n_t:
                gbuf[n_t:n_t_pad].zero_()
            dtxt = gbuf[:n_t].view_as(txt_c)
            dobj = gbuf[n_t_pad:].view_as(obj_c)
        else:
            gbuf = dtxt = dobj = None
        L.check(lib.ce_ot_fwd_bwd(
            txt_c.data_ptr(), M * D, obj_c.data_ptr() + slot * D * esz, (N + slot) * D,
            tm.data_ptr(), M, om.data_ptr() + slot * msz, N + slot, kind_t, B, M, N, D, dt,
            float(beta), int(iters), int(k), float(loss_scale), dist.data_ptr(), loss.data_ptr(),
            L.ptr(dtxt), 0 if dobj is None else dobj.data_ptr() + slot * D * esz,
            0 if (dobj is None or not slot) else dobj.data_ptr(), ws.data_ptr(), nbytes, L.stream_ptr()),
            "OT forward")
        ctx.stash = (dtxt, dobj, gbuf)
        ctx.loss_scale = float(loss_scale)
        ctx.consumed = False
        ctx.set_materialize_grads(False)   # an unused output arrives as None, not as zeros
        return loss[0], dist

    @staticmethod
    def backward(ctx, g_loss, g_dist):
        if ctx.consumed:
            raise RuntimeError("clip_event_b200 OT: backward through the stashed gradients a second time; "
                               "run the forward again (retain_graph is not supported on this path)")
        dtxt, dobj, gbuf = ctx.stash
        if dtxt is None:
            return (None,) * 9
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
