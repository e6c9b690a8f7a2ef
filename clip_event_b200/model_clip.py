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


class CriterionContrastive(nn.Module):
    """model_clip.py:620-662."""

    def __init__(self, constrastive_loss):
        super().__init__()
        if constrastive_loss not in ("ce", "bce", "kl"):
            raise RuntimeError("Invalid constrastive_loss '{}'. ".format(constrastive_loss))
        self.constrastive_loss = constrastive_loss

    def forward(self, logits_per_image, logits_per_text, labels_per_image=None, labels_per_text=None,
                index_pos=None, constrastive_overbatch=True):
        if not isinstance(logits_per_image, LazyLogits) or not isinstance(logits_per_text, LazyLogits):
            raise RuntimeError("clip_event_b200.CriterionContrastive fuses the similarity GEMM with the "
                               "cross-entropy: pass the LazyLogits returned by ClipEventHead.forward")
        img, txt, ls = logits_per_image.rows, logits_per_image.cols, logits_per_image.logit_scale
        B = img.shape[0]
        dev = img.device
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
        if logits_per_image.per_instance != (not constrastive_overbatch):
            raise RuntimeError("constrastive_overbatch=%s does not match the logits the head produced "
                               "(ClipEventHead.set_hyps)" % constrastive_overbatch)
        if constrastive_overbatch:
            if self.constrastive_loss != "ce":
                # dataset_voa.py:628-631: "Set constrastive_overbatch=false for constrative_loss=='bce'."
                raise RuntimeError("Set constrastive_overbatch=false for constrative_loss=='bce'.")
            loss_i, loss_t = F_.contrastive_over_batch(img, txt, ls, labels_per_image, labels_per_text, index_pos)
        else:
            loss_i, loss_t = F_.contrastive_over_instance(img, txt, ls, labels_per_image, labels_per_text,
                                                          index_pos, self.constrastive_loss)
        return {"loss_i": loss_i.to(img.dtype), "loss_t": loss_t.to(img.dtype)}


class CriterionAlignment(nn.Module):
    """model_clip.py:664-715: OT distance between text nodes and image nodes, summed, times 0.01."""

    def __init__(self):
        super().__init__()

    def mask2pad(self, x_mask):
        return x_mask == 0

    def forward(self, entitytxt_vec, object_vec, entitytxt_num, object_num):
        loss, _ = F_.ot_alignment(entitytxt_vec, object_vec, entitytxt_num, object_num, drop_slot0=True)
        return {"loss_ot": loss.to(entitytxt_vec.dtype)}
