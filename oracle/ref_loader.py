"""Loader for the UNMODIFIED reference loss-head modules (test / benchmark infrastructure only).

The product (clip_event_b200/) never imports this file.  Only tests/, tools/, __graft_entry__.smoke()
and bench.py's CPU legs may, as the tier rules say.

The reference is a flat script directory whose files import each other by bare module name
(model_clip.py:10-11 -> model_ot, utils_image), so its directory goes on sys.path:
  1. oracle/_ref/            the byte-for-byte copy written by oracle/make_ref.sh (travels to the GPU box)
  2. /root/reference/src/clip-event   (this container only)
"""
from __future__ import annotations

import importlib
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
CANDIDATES = [os.path.join(HERE, "_ref"), "/root/reference/src/clip-event"]

_mods = None


def available() -> bool:
    return any(os.path.exists(os.path.join(d, "model_clip.py")) for d in CANDIDATES)


def load():
    """(model_clip, model_ot) of the reference, imported unmodified."""
    global _mods
    if _mods is None:
        for d in CANDIDATES:
            if os.path.exists(os.path.join(d, "model_clip.py")) and os.path.exists(os.path.join(d, "model_ot.py")):
                if d not in sys.path:
                    sys.path.insert(0, d)
                _mods = (importlib.import_module("model_clip"), importlib.import_module("model_ot"), d)
                break
        else:
            raise RuntimeError("reference modules not found (run oracle/make_ref.sh where /root/reference exists)")
    return _mods[0], _mods[1]


def source_dir() -> str:
    load()
    return _mods[2]


def head_logits(image_features, text_features, logit_scale, overbatch=True):
    """The ten lines of CLIP.forward between the encoders and the return (model_clip.py:495-520);
    they sit inside CLIP.forward behind the encoders, so the harness has to restate them."""
    image_features = image_features / image_features.norm(dim=-1, keepdim=True)
    text_features = text_features / text_features.norm(dim=-1, keepdim=True)
    s = logit_scale.exp()
    logits_per_text = s * text_features @ image_features.t()
    if overbatch:
        logits_per_image = s * image_features @ text_features.t()
    else:
        b, d = image_features.size(0), text_features.size(-1)
        logits_per_image = (s * torch.bmm(image_features.unsqueeze(1),
                                          text_features.view(b, -1, d).transpose(-2, -1))).squeeze(1)
    return logits_per_image, logits_per_text


class ReferenceLossHead:
    """engine.py:48-67,87-88 on the reference's own criteria: one fwd+bwd step on CPU tensors."""

    def __init__(self, constrastive_loss="ce"):
        mc, _ = load()
        self.crit = mc.CriterionContrastive(constrastive_loss)
        self.crit_ot = mc.CriterionAlignment()

    def step(self, img, txt, ls, lpi, lpt, idx, etxt=None, obj=None, tnum=None, onum=None, overbatch=True):
        leaves = [t.detach().clone().requires_grad_(True) for t in (img, txt, ls)]
        a, b = head_logits(leaves[0], leaves[1], leaves[2], overbatch)
        loss_dict = self.crit(a, b, lpi, lpt, index_pos=idx, constrastive_overbatch=overbatch)
        ot_leaves = []
        if etxt is not None:
            ot_leaves = [t.detach().clone().requires_grad_(True) for t in (etxt, obj)]
            loss_dict.update(self.crit_ot(ot_leaves[0], ot_leaves[1], tnum, onum))
        losses = sum(loss_dict.values())          # engine.py:67
        losses.backward()                         # engine.py:88
        grads = [t.grad for t in leaves + ot_leaves]
        return {k: v.detach() for k, v in loss_dict.items()}, grads
