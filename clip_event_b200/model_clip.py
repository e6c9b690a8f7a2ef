"""CUDA-backed mirror of the loss head in the reference's ``model_clip.py``.

Reference (src/clip-event/model_clip.py):
  :495-528  tail of ``CLIP.forward``  -> :class:`ClipEventHead`
  :620-662  ``CriterionContrastive`` -> :class:`CriterionContrastive`
  :664-715  ``CriterionAlignment``   -> :class:`CriterionAlignment`
Call signatures, argument meaning, returned dict keys and error style (RuntimeError) are the
reference's, so ``engine.train_one_epoch`` (engine.py:48-67) runs unchanged on top of them.

The fused similarity+cross-entropy kernel needs the *features*, while the reference's criterion is
handed *logits*.  The head therefore returns :class:`LazyLogits` handles -- tensor-like objects that
remember (features, logit_scale) -- and the criterion recognises them.  Anything else that touches
a handle (e.g. ``.softmax`` in preprocess_description_contrastive.py:129-131) materialises the dense
matrix with the same tcgen05 GEMM.
"""
from __future__ import annotations

import numpy as np
import torch
from torch import nn

from . import functional as F_
from . import _lib as L


class LazyLogits:
    """``exp(logit_scale) * normalize(rows) @ normalize(cols).T`` kept unevaluated."""

    def __init__(self, rows, cols, logit_scale, role, per_instance=False):
        self.rows, self.cols, self.logit_scale = rows, cols, logit_scale
        self.role = role                    # 'per_image' | 'per_text'
        self.per_instance = per_instance    # model_clip.py:509-520 ([B,T] against own descriptions)
        self._dense = None

    @property
    def shape(self):
        if self.per_instance:
            return torch.Size((self.rows.shape[0], self.cols.shape[0] // self.rows.shape[0]))
        return torch.Size((self.rows.shape[0], self.cols.shape[0]))

    def size(self, dim=None):
        return self.shape if dim is None else self.shape[dim]

    @property
    def device(self):
        return self.rows.device

    @property
    def dtype(self):
        return self.rows.dtype

    def materialize(self) -> torch.Tensor:
        """Dense logits (no autograd), computed by the same tensor-core GEMM."""
        if self._dense is None:
            full = F_.similarity_logits(self.rows, self.cols, self.logit_scale)
            if self.per_instance:
                B = self.rows.shape[0]
                T = self.cols.shape[0] // B
                idx = torch.arange(B, device=full.device)
                full = full.view(B, B, T)[idx, idx]
            self._dense = full.to(self.rows.dtype)
        return self._dense

    def __getattr__(self, name):  # tensor methods (softmax, argmax, cpu, ...) act on the dense matrix
        if name.startswith("_") or name in ("rows", "cols", "logit_scale", "role", "per_instance"):
            raise AttributeError(name)
        return getattr(self.materialize(), name)

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        conv = lambda a: a.materialize() if isinstance(a, LazyLogits) else a
        return func(*[conv(a) for a in args], **{k: conv(v) for k, v in kwargs.items()})

    def __repr__(self):
        return "LazyLogits(%s, shape=%s)" % (self.role, tuple(self.shape))


class ClipEventHead(nn.Module):
    """The part of ``CLIP`` between the encoders and the criteria (model_clip.py:330,343-346,495-528).

    ``forward(image_features, text_features)`` takes the encoder outputs ([B,D], [B*T,D]) and returns
    ``(logits_per_image, logits_per_text)`` as :class:`LazyLogits`.
    """

    def __init__(self, constrastive_overbatch=True, alignment=True):
        super().__init__()
        self.logit_scale = nn.Parameter(torch.ones([]) * np.log(1 / 0.07))
        self.set_hyps(constrastive_overbatch, alignment)

    def set_hyps(self, constrastive_overbatch, alignment, multiattention=False):
        self.constrastive_overbatch = constrastive_overbatch
        self.alignment = alignment
        self.multiattention = multiattention

    def forward(self, image_features, text_features):
        L.require_cuda(image_features, text_features)
        lpt = LazyLogits(text_features, image_features, self.logit_scale, "per_text")
        lpi = LazyLogits(image_features, text_features, self.logit_scale, "per_image",
                         per_instance=not self.constrastive_overbatch)
        return lpi, lpt

    def sim_entity(self, image_features, text_features):
        """model_clip.py:531-552 after the encoders: the node embeddings pass through un-normalised."""
        return image_features, text_features

    def sim_entity_packed(self, image_features, text_features, object_num, entitytxt_num):
        """The same node sets without their padding: ``(image nodes, text nodes)`` as
        :class:`functional.PackedNodes`, the whole-image slot dropped (model_clip.py:686), ready for
        ``CriterionAlignment.forward_packed``.  The padded slots of dataset_voa.py:532-544,566-577 then never
        reach the OT kernel: with prefix lengths uniform in 1..max it moves about half of the bytes."""
        return (F_.pack_nodes(image_features, F_.num_mask(object_num), drop_first=True),
                F_.pack_nodes(text_features, F_.num_mask(entitytxt_num)))


class ProjectionTail(nn.Module):
    """The last step of either encoder, SURVEY.md 8f-2: ``LayerNorm(hidden[arange, token]) @ proj``.

    * image side (model_clip.py:253-260): ``ProjectionTail(width, embed_dim)`` holds ``ln_post`` and ``proj``
      (initialised as model_clip.py:214-216); call it on the transformer output ``[B, L, width]``.
    * text side (model_clip.py:412-415): the same module holds ``ln_final`` and ``text_projection``; pass
      ``token_index=text.argmax(dim=-1)`` (the eot token).  LayerNorm acts per token, so normalising only the
      selected token equals the reference's ``ln_final(x)[arange, eot]``.

    One row kernel (gather + LayerNorm, fp32 statistics as the reference's LayerNorm subclass,
    model_clip.py:157-163) feeds the tcgen05 GEMM; the features come back together with their squared L2
    norms (``self.last_norm2``), which is what ``ClipEventHead`` computes next (model_clip.py:496-497).
    """

    def __init__(self, width: int, embed_dim: int, layer_norm: bool = True, eps: float = 1e-5):
        super().__init__()
        self.eps = eps
        if layer_norm:
            self.weight = nn.Parameter(torch.ones(width))
            self.bias = nn.Parameter(torch.zeros(width))
        else:
            self.register_parameter("weight", None)
            self.register_parameter("bias", None)
        self.proj = nn.Parameter((width ** -0.5) * torch.randn(width, embed_dim))
        self.last_norm2 = None

    def forward(self, hidden, token_index=None):
        feat, norm2 = F_.projection_tail(hidden, self.proj, self.weight, self.bias, token_index, self.eps)
        self.last_norm2 = norm2
        return feat


def _localise_labels(labels_per_image, labels_per_text, b, cols_local, group):
    """The reference's collate_fn numbers rows and columns within ONE rank's batch
    (dataset_voa.py:617-663).  The sharded kernels want labels_per_image as GLOBAL column indices and
    labels_per_text as GLOBAL row indices: shift by this rank's offsets."""
    import torch.distributed as dist
    rank = dist.get_rank(group)
    # a training loop hands over the same label tensors every step: shift them once
    key = (labels_per_image.data_ptr(), labels_per_text.data_ptr(), labels_per_image._version, labels_per_text._version,
           rank, b, cols_local)
    hit = _LOCALISED.get(key)
    if hit is None:
        if len(_LOCALISED) > 16:
            _LOCALISED.clear()
        hit = _LOCALISED[key] = (labels_per_image + rank * cols_local, labels_per_text + rank * b,
                                 labels_per_image, labels_per_text)   # keep the sources alive: the key is their address
    return hit[0], hit[1]


_LOCALISED = {}


class CriterionContrastive(nn.Module):
    """model_clip.py:620-662.

    ``forward`` takes what ``ClipEventHead.forward`` returns (LazyLogits -> fused tcgen05 GEMM +
    cross-entropy, the logits are never written to HBM) or, like the reference, any materialised
    logits tensors (memory-bound row kernels, ``functional.dense_contrastive``).

    ``group`` (a torch.distributed process group, or ``True`` for the default group) turns on global
    negatives: image embeddings are all-gathered so that every rank scores against the global batch
    and the image gradient is reduce-scattered back (the reference's own
    ``utils.gather_tensors``, utils.py:192-206, is dead code).  Labels stay the per-rank ones the
    reference's collate_fn builds.  ``ddp_average`` (default True with a group): feature gradients are
    scaled for an encoder wrapped in DistributedDataParallel -- see ``distributed.global_contrastive``.
    """

    def __init__(self, constrastive_loss, group=None, ddp_average=True, compute=None):
        super().__init__()
        if constrastive_loss not in ("ce", "bce", "kl"):
            raise RuntimeError("Invalid constrastive_loss '{}'. ".format(constrastive_loss))
        self.constrastive_loss = constrastive_loss
        self.group = group
        self.ddp_average = ddp_average
        self._compute = compute     # tests inject a CPU stand-in for the kernels

    def _pg(self):
        return None if self.group is True else self.group

    def forward(self, logits_per_image, logits_per_text, labels_per_image=None, labels_per_text=None,
                index_pos=None, constrastive_overbatch=True):
        lazy = isinstance(logits_per_image, LazyLogits) and isinstance(logits_per_text, LazyLogits)
        if not lazy and (isinstance(logits_per_image, LazyLogits) or isinstance(logits_per_text, LazyLogits)):
            logits_per_image = logits_per_image.materialize() if isinstance(logits_per_image, LazyLogits) else logits_per_image
            logits_per_text = logits_per_text.materialize() if isinstance(logits_per_text, LazyLogits) else logits_per_text
        if lazy:
            img, txt, ls = logits_per_image.rows, logits_per_image.cols, logits_per_image.logit_scale
            B, dev, out_dtype = img.shape[0], img.device, img.dtype
        else:
            B, dev, out_dtype = logits_per_image.shape[0], logits_per_image.device, logits_per_image.dtype
        if labels_per_image is None:
            labels_per_image = torch.arange(B, device=dev)     # model_clip.py:635-637
        if labels_per_text is None:
            labels_per_text = torch.arange(B, device=dev)      # model_clip.py:638-640
        if index_pos is None:
            raise RuntimeError("index_pos is required (the reference index_selects with it, model_clip.py:655)")
        if self.constrastive_loss == "kl":
            # the reference's 'kl' path is unusable: its collate_fn calls torch.zeros() with no shape
            # (dataset_voa.py:642) and feeds raw logits to KLDivLoss (model_clip.py:628-629)
            raise RuntimeError("constrastive_loss 'kl' is broken in the reference and not provided")
        if not lazy:
            if self.group is not None:
                raise RuntimeError("global negatives need the features: pass the LazyLogits of ClipEventHead.forward")
            loss_i, loss_t = F_.dense_contrastive(logits_per_image, logits_per_text, labels_per_image, labels_per_text,
                                                  index_pos, self.constrastive_loss)
            return {"loss_i": loss_i.to(out_dtype), "loss_t": loss_t.to(out_dtype)}
        if logits_per_image.per_instance != (not constrastive_overbatch):
            raise RuntimeError("constrastive_overbatch=%s does not match the logits the head produced "
                               "(ClipEventHead.set_hyps)" % constrastive_overbatch)
        if constrastive_overbatch:
            if self.constrastive_loss != "ce":
                # dataset_voa.py:628-631: "Set constrastive_overbatch=false for constrative_loss=='bce'."
                raise RuntimeError("Set constrastive_overbatch=false for constrative_loss=='bce'.")
            if self.group is not None:
                from . import distributed as cd
                lpi, lpt = _localise_labels(labels_per_image.to(dev), labels_per_text.to(dev), B, txt.shape[0], self._pg())
                loss_i, loss_t = cd.global_contrastive(img, txt, ls, lpi, lpt, index_pos, group=self._pg(),
                                                       compute=self._compute, ddp_average=self.ddp_average)
            else:
                loss_i, loss_t = F_.contrastive_over_batch(img, txt, ls, labels_per_image, labels_per_text, index_pos)
        else:
            if self.group is not None:
                raise RuntimeError("global negatives are defined for constrastive_overbatch=True")
            loss_i, loss_t = F_.contrastive_over_instance(img, txt, ls, labels_per_image, labels_per_text,
                                                          index_pos, self.constrastive_loss)
        return {"loss_i": loss_i.to(out_dtype), "loss_t": loss_t.to(out_dtype)}


class CriterionAlignment(nn.Module):
    """model_clip.py:664-715: OT distance between text nodes and image nodes, summed, times 0.01.

    With ``group`` the returned loss is the sum over the GLOBAL batch (one scalar all-reduce); the
    gradient w.r.t. this rank's nodes is the local term, as in the reference under DDP."""

    def __init__(self, group=None):
        super().__init__()
        self.group = group

    def mask2pad(self, x_mask):
        return x_mask == 0

    def forward(self, entitytxt_vec, object_vec, entitytxt_num, object_num):
        # *_num semantics whatever the dtype: nonzero = valid node (the reference applies mask2pad, :688-690)
        tnum, onum = F_.num_mask(entitytxt_num), F_.num_mask(object_num)
        if self.group is not None:
            from . import distributed as cd
            loss = cd.sharded_alignment(entitytxt_vec, object_vec, tnum, onum, group=None if self.group is True else self.group)
        else:
            loss, _ = F_.ot_alignment(entitytxt_vec, object_vec, tnum, onum, drop_slot0=True)
        return {"loss_ot": loss.to(entitytxt_vec.dtype)}

    def forward_packed(self, entitytxt_nodes, object_nodes):
        """The criterion on packed node sets (``functional.PackedNodes`` from ``sim_entity_packed`` /
        ``pack_nodes``): same ``{'loss_ot': ...}`` as :meth:`forward` on the padded, masked batch."""
        if self.group is not None:
            raise RuntimeError("forward_packed: shard the packed batch by sample and all-reduce the scalar yourself")
        loss, _ = F_.ot_alignment_packed(entitytxt_nodes, object_nodes)
        return {"loss_ot": loss.to(entitytxt_nodes.rows.dtype)}


class LossHeadStep(nn.Module):
    """engine.py:48-67 in ONE call: both criteria of the loss head, ready for ``sum(...).backward()``.

        step = LossHeadStep(head)                       # head: ClipEventHead (owns logit_scale)
        loss_dict = step(image_features, text_features, labels_per_image, labels_per_text, index_pos,
                         entitytxt_vec, object_vec, entitytxt_num, object_num)
        losses = sum(loss_dict.values()); losses.backward()          # engine.py:67,88 unchanged

    Same numbers as ``CriterionContrastive('ce')`` + ``CriterionAlignment()`` called one after the
    other; what the single call buys is overlap: the tensor-core chain and the OT chain run on two
    streams, and their gradients are formed together with the losses (``backward`` only scales).
    With ``group`` the batch is sharded over the ranks with global negatives (three NCCL launches
    per step, see ``distributed.global_loss_head_step``); labels are the per-rank ones.
    """

    def __init__(self, head: ClipEventHead, group=None, ddp_average=True):
        super().__init__()
        self.head = head
        self.group = group
        self.ddp_average = ddp_average

    def forward(self, image_features, text_features, labels_per_image, labels_per_text, index_pos,
                entitytxt_vec, object_vec, entitytxt_num, object_num):
        if not self.head.constrastive_overbatch:
            raise RuntimeError("LossHeadStep covers constrastive_overbatch=True; use the criteria for the over-instance modes")
        ls = self.head.logit_scale
        if self.group is not None:
            from . import distributed as cd
            pg = None if self.group is True else self.group
            B = image_features.shape[0]
            lpi, lpt = _localise_labels(labels_per_image, labels_per_text, B, text_features.shape[0], pg)
            li, lt, lo = cd.global_loss_head_step(image_features, text_features, ls, lpi, lpt, index_pos, entitytxt_vec,
                                                  object_vec, entitytxt_num, object_num, group=pg,
                                                  ddp_average=self.ddp_average)
        else:
            # the losses come back in the dtypes the reference's criteria return them in (cast inside the call)
            li, lt, lo = F_.loss_head_step(image_features, text_features, ls, labels_per_image, labels_per_text,
                                           index_pos, entitytxt_vec, object_vec, entitytxt_num, object_num,
                                           cast_losses=True)
            return {"loss_i": li, "loss_t": lt, "loss_ot": lo}
        dt = image_features.dtype
        return {"loss_i": li.to(dt), "loss_t": lt.to(dt), "loss_ot": lo.to(entitytxt_vec.dtype)}
