"""CPU oracle for the CLIP-Event training loss head.  TEST INFRASTRUCTURE ONLY.

This file is a plain-PyTorch (CPU) restatement of the reference algorithm for the
hot path named in BASELINE.json / SURVEY.md section 8.  It is NOT part of the
product: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it.  The product
(``clip_event_b200``) never imports it and fails loudly without its CUDA library.

Parity pin: every function below is checked against the UNMODIFIED reference
(imported from /root/reference/src/clip-event in the build container) by
``tests/golden/make_golden.py``; the resulting vectors are committed under
``tests/golden/*.npz`` and re-checked on every run by ``tests/test_oracle_golden.py``.

Each function cites the reference lines it restates (paths relative to the
reference root, ``src/clip-event/``).  Everything is dtype-generic: run it in
float32 to mirror the reference bit-for-bit-ish, or float64 for a noise-free
ground truth.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch

# --------------------------------------------------------------------------- #
# Label / index contract  (dataset_voa.py:397-399, 615-663)
# --------------------------------------------------------------------------- #


def canonical_labels(batch: int, descs_per_image: int, overbatch: bool = True,
                     num_pos: int = 1, device="cpu"):
    """Labels exactly as the reference collate_fn builds them for ``'ce'``.

    dataset_voa.py:617-625  labels_per_image = arange(B)*T (over batch) or zeros (over instance)
    dataset_voa.py:652-655  labels_per_text  = [0..0, 1..1, ...] (length B*T)
    dataset_voa.py:657-663  index_pos        = nonzero([1]*num_pos + [0]*num_neg per image)
    """
    B, T = batch, descs_per_image
    ar = torch.arange(B, device=device, dtype=torch.int64)
    labels_per_image = ar * T if overbatch else torch.zeros(B, dtype=torch.int64, device=device)
    labels_per_text = ar.repeat_interleave(T)
    pos_mask = torch.zeros(B, T, dtype=torch.int64, device=device)
    pos_mask[:, :num_pos] = 1
    index_pos = pos_mask.flatten().nonzero().flatten()
    return labels_per_image, labels_per_text, index_pos


# --------------------------------------------------------------------------- #
# Similarity scoring head  (model_clip.py:495-528)
# --------------------------------------------------------------------------- #


def similarity_logits(image_features: torch.Tensor, text_features: torch.Tensor,
                      logit_scale: torch.Tensor, overbatch: bool = True
                      ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Tail of ``CLIP.forward`` after the encoders.

    model_clip.py:496-497  L2-normalise both sides, NO eps
    model_clip.py:502      s = exp(logit_scale), not clamped
    model_clip.py:504      logits_per_text = s * txt @ img.T      (always over batch)
    model_clip.py:506-520  logits_per_image = s * img @ txt.T     (over batch)
                           or the per-image bmm against its own T descriptions
    """
    img = image_features / image_features.norm(dim=-1, keepdim=True)
    txt = text_features / text_features.norm(dim=-1, keepdim=True)
    s = logit_scale.exp()
    logits_per_text = s * txt @ img.t()
    if overbatch:
        logits_per_image = s * img @ txt.t()
    else:
        B, D = img.shape
        per_inst = txt.view(B, -1, D)                       # [B, T, D]
        logits_per_image = (s * torch.bmm(img.unsqueeze(1), per_inst.transpose(1, 2))).squeeze(1)
    return logits_per_image, logits_per_text


# --------------------------------------------------------------------------- #
# InfoNCE criterion  (model_clip.py:620-662)
# --------------------------------------------------------------------------- #


def _ce_mean(logits: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """nn.CrossEntropyLoss() default = mean over rows of (logsumexp - picked)."""
    lse = torch.logsumexp(logits, dim=-1)
    picked = logits.gather(1, target.view(-1, 1)).squeeze(1)
    return (lse - picked).mean()


def contrastive_criterion(logits_per_image, logits_per_text, labels_per_image=None,
                          labels_per_text=None, index_pos=None, overbatch=True,
                          kind: str = "ce") -> Dict[str, torch.Tensor]:
    """``CriterionContrastive.forward``.

    model_clip.py:635-640  default labels = arange(B)
    model_clip.py:646-651  loss_i = loss_func_image(logits_per_image, labels_per_image)
                           ('ce': mean CE; 'bce': BCEWithLogits, mean over all entries)
    model_clip.py:655-659  rows index_pos of logits_per_text / labels_per_text -> mean CE = loss_t
    """
    B = logits_per_image.shape[0]
    if labels_per_image is None:
        labels_per_image = torch.arange(B, device=logits_per_image.device)
    if labels_per_text is None:
        labels_per_text = torch.arange(B, device=logits_per_image.device)
    if kind == "ce":
        loss_i = _ce_mean(logits_per_image, labels_per_image)
    elif kind == "bce":
        # nn.BCEWithLogitsLoss(): mean over every element of softplus(x) - y*x
        x, y = logits_per_image, labels_per_image.to(logits_per_image.dtype)
        loss_i = (torch.nn.functional.softplus(x) - y * x).mean()
    else:
        raise RuntimeError("Invalid constrastive_loss '{}'. ".format(kind))
    sel_logits = logits_per_text.index_select(0, index_pos)
    sel_labels = labels_per_text.index_select(0, index_pos)
    loss_t = _ce_mean(sel_logits, sel_labels)
    return {"loss_i": loss_i, "loss_t": loss_t}


# --------------------------------------------------------------------------- #
# OT: cosine cost, IPOT, trace, distance  (model_ot.py:8-84)
# --------------------------------------------------------------------------- #


def cost_matrix_cosine(x: torch.Tensor, y: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """model_ot.py:8-18.  [B,M,D],[B,N,D] -> 1 - cos, [B,M,N]; F.normalize clamps the norm at eps."""
    xn = x / x.norm(dim=-1, keepdim=True).clamp_min(eps)
    yn = y / y.norm(dim=-1, keepdim=True).clamp_min(eps)
    return 1 - xn @ yn.transpose(1, 2)


def trace_batched(x: torch.Tensor) -> torch.Tensor:
    """model_ot.py:21-29.  Batched trace of [B,n,n]."""
    return torch.diagonal(x, dim1=-2, dim2=-1).sum(-1)


@torch.no_grad()
def ipot(C, x_len, x_pad, y_len, y_pad, joint_pad, beta, iteration, k):
    """model_ot.py:32-63.  Inexact proximal point OT; C [B,M,N] -> plan T [B,N,M].

    :36-45  sigma = 1/x_len (0 at text pads), T = 1, A = exp(-C^T/beta), T and A zeroed at joint pads
    :52-53  pad guards 1e4 keep 1/(...) finite on padded rows / columns
    :55-61  iteration x [ Q = A*T ; k x ( delta = 1/(y_len*Q sigma + y_mask) ;
                                         sigma = 1/(x_len*delta Q + x_mask) ) ; T = delta*Q*sigma ]
    :62     final re-mask
    """
    b, m, n = C.shape
    dt = C.dtype
    sigma = (torch.ones(b, m, dtype=dt) / x_len.view(b, 1)).masked_fill(x_pad, 0)
    jp = joint_pad.transpose(1, 2)
    T = torch.ones(b, n, m, dtype=dt).masked_fill(jp, 0)
    A = torch.exp(-C.transpose(1, 2) / beta).masked_fill(jp, 0)
    xl = x_len.view(b, 1, 1)
    yl = y_len.view(b, 1, 1)
    x_guard = (x_pad.to(dt) * 1e4).view(b, 1, m)
    y_guard = (y_pad.to(dt) * 1e4).view(b, 1, n)
    delta = None
    for _ in range(iteration):
        Q = A * T
        for _ in range(k):
            # the reference reshapes sigma once per OUTER iteration, so it raises for k > 1;
            # reshaping here gives the intended recurrence and is identical for k == 1
            delta = 1 / (yl * Q.matmul(sigma.view(b, m, 1)).view(b, 1, n) + y_guard)
            sigma = 1 / (xl * delta.matmul(Q) + x_guard)
        T = delta.view(b, n, 1) * Q * sigma
    return T.masked_fill(jp, 0)


def optimal_transport_dist(txt_emb, img_emb, txt_pad, img_pad, cost=None,
                           beta: float = 0.5, iteration: int = 50, k: int = 1):
    """model_ot.py:66-84.  [B,M,D],[B,N,D],[B,M]bool,[B,N]bool -> [B].

    :71     cost = cosine cost unless given
    :73-74  cost zeroed at joint pads (the reference does it in place: grad is 0 there too)
    :76-79  valid counts as floats
    :81     T = ipot(cost.detach()) -- no gradient through the solver
    :83     distance = trace(cost @ T.detach()) = sum_{m,n} C[m,n] * T[n,m]
    """
    if cost is None:
        cost = cost_matrix_cosine(txt_emb, img_emb)
    joint_pad = txt_pad.unsqueeze(-1) | img_pad.unsqueeze(-2)
    cost = cost.masked_fill(joint_pad, 0)
    txt_len = (txt_pad.size(1) - txt_pad.sum(dim=1)).to(cost.dtype)
    img_len = (img_pad.size(1) - img_pad.sum(dim=1)).to(cost.dtype)
    T = ipot(cost.detach(), txt_len, txt_pad, img_len, img_pad, joint_pad, beta, iteration, k)
    return trace_batched(cost.matmul(T.detach()))


def alignment_criterion(entitytxt_vec, object_vec, entitytxt_num, object_num,
                        compute_dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """``CriterionAlignment.forward``  (model_clip.py:679-715).

    :686      image nodes = object_vec[:, 1:]  (slot 0, the whole image, is dropped)
    :688-690  pads = (mask == 0)
    :699-702  OT distance in fp32 (``compute_dtype`` lets tests ask for fp64), cast back to input dtype
    :707      loss_ot = sum over the batch (left to right) * 0.01
    """
    txt_nodes = entitytxt_vec
    img_nodes = object_vec[:, 1:]
    txt_pad = entitytxt_num == 0
    img_pad = object_num[:, 1:] == 0
    dist = optimal_transport_dist(txt_nodes.to(compute_dtype), img_nodes.to(compute_dtype),
                                  txt_pad, img_pad).to(txt_nodes.dtype)
    total = dist[0] * 0 if dist.numel() == 0 else dist[0]
    for i in range(1, dist.shape[0]):
        total = total + dist[i]
    return {"loss_ot": total * 0.01}


# --------------------------------------------------------------------------- #
# Closed-form gradients -- used to cross-check autograd and the kernels
# --------------------------------------------------------------------------- #


def ot_closed_form_grads(txt_emb, img_emb, txt_pad, img_pad, ddist, eps: float = 1e-5,
                         beta: float = 0.5, iteration: int = 50, k: int = 1):
    """d(sum_b ddist[b]*dist[b]) / d(txt_emb, img_emb) without autograd.

    dC = T^T (0 at pads);  dx^ = -dC y^ ;  dy^ = -dC^T x^ ;  then through x/max(|x|,eps):
    |x| >= eps: dx = (dx^ - x^ (x^ . dx^)) / |x| ;  |x| < eps: dx = dx^ / eps.
    """
    nx = txt_emb.norm(dim=-1, keepdim=True)
    ny = img_emb.norm(dim=-1, keepdim=True)
    xh = txt_emb / nx.clamp_min(eps)
    yh = img_emb / ny.clamp_min(eps)
    joint_pad = txt_pad.unsqueeze(-1) | img_pad.unsqueeze(-2)
    cost = (1 - xh @ yh.transpose(1, 2)).masked_fill(joint_pad, 0)
    txt_len = (txt_pad.size(1) - txt_pad.sum(dim=1)).to(cost.dtype)
    img_len = (img_pad.size(1) - img_pad.sum(dim=1)).to(cost.dtype)
    T = ipot(cost, txt_len, txt_pad, img_len, img_pad, joint_pad, beta, iteration, k)
    dist = (cost * T.transpose(1, 2)).sum((1, 2))
    dC = (T.transpose(1, 2) * ddist.view(-1, 1, 1)).masked_fill(joint_pad, 0)
    dxh = -dC @ yh
    dyh = -dC.transpose(1, 2) @ xh

    def through_normalize(v, vh, n, g):
        big = n >= eps
        proj = g - vh * (vh * g).sum(-1, keepdim=True)
        return torch.where(big, proj / n.clamp_min(eps), g / eps)

    return dist, through_normalize(txt_emb, xh, nx, dxh), through_normalize(img_emb, yh, ny, dyh)


def contrastive_closed_form(image_features, text_features, logit_scale, labels_per_image,
                            labels_per_text, index_pos, g_i: float = 1.0, g_t: float = 1.0):
    """Over-batch 'ce' losses and grads from ONE logits matrix (SURVEY.md 8a-2 identity).

    logits_per_text[index_pos] == logits_per_image[:, index_pos]^T, so
    loss_i = mean_b(rowLSE - L[b, lab_i[b]]),  loss_t = mean_p(colLSE[pos_p] - L[lab_t[pos_p], pos_p]).
    G = g_i*(softmax_row - onehot)/B + scatter_cols(g_t*(softmax_col - onehot)/P).
    """
    ni = image_features.norm(dim=-1, keepdim=True)
    nt = text_features.norm(dim=-1, keepdim=True)
    ih, th = image_features / ni, text_features / nt
    s = logit_scale.exp()
    L = s * ih @ th.t()
    B, P = L.shape[0], index_pos.numel()
    row_lse = torch.logsumexp(L, 1)
    loss_i = (row_lse - L.gather(1, labels_per_image.view(-1, 1)).squeeze(1)).mean()
    Lp = L[:, index_pos]                                     # [B, P]
    col_lse = torch.logsumexp(Lp, 0)
    lab_rows = labels_per_text[index_pos]
    loss_t = (col_lse - Lp[lab_rows, torch.arange(P)]).mean()
    G = torch.exp(L - row_lse.view(-1, 1))
    G[torch.arange(B), labels_per_image] -= 1
    G = G * (g_i / B)
    Gc = torch.exp(Lp - col_lse.view(1, -1))
    Gc[lab_rows, torch.arange(P)] -= 1
    G.index_add_(1, index_pos, Gc * (g_t / P))
    dih = s * G @ th
    dth = s * G.t() @ ih
    dls = (G * L).sum()
    dimg = (dih - ih * (ih * dih).sum(-1, keepdim=True)) / ni
    dtxt = (dth - th * (th * dth).sum(-1, keepdim=True)) / nt
    return loss_i, loss_t, dimg, dtxt, dls


# --------------------------------------------------------------------------- #
# Whole loss head, forward + backward (engine.py:48-67, 87-88)
# --------------------------------------------------------------------------- #


def loss_head_step(image_features, text_features, logit_scale, labels_per_image, labels_per_text,
                   index_pos, entitytxt_vec=None, object_vec=None, entitytxt_num=None,
                   object_num=None, overbatch: bool = True, kind: str = "ce",
                   compute_dtype=torch.float32):
    """One fwd+bwd of the loss head exactly as ``engine.train_one_epoch`` drives it.

    engine.py:48     logits = model(...)   -> similarity_logits
    engine.py:52-53  loss_dict = criterion(...)
    engine.py:57-64  if alignment: loss_dict.update(criterion_ot(...))
    engine.py:67     losses = sum(loss_dict.values())
    engine.py:88     losses.backward()
    Returns (loss_dict, grads) with grads keyed by input name.
    """
    leaves = {"image_features": image_features, "text_features": text_features,
              "logit_scale": logit_scale}
    if entitytxt_vec is not None:
        leaves["entitytxt_vec"] = entitytxt_vec
        leaves["object_vec"] = object_vec
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in leaves.items()}
    lpi, lpt = similarity_logits(leaves["image_features"], leaves["text_features"],
                                 leaves["logit_scale"], overbatch)
    loss_dict = contrastive_criterion(lpi, lpt, labels_per_image, labels_per_text, index_pos,
                                      overbatch, kind)
    if entitytxt_vec is not None:
        loss_dict.update(alignment_criterion(leaves["entitytxt_vec"], leaves["object_vec"],
                                             entitytxt_num, object_num, compute_dtype))
    total = sum(loss_dict.values())
    total.backward()
    grads = {k: v.grad for k, v in leaves.items()}
    return {k: v.detach() for k, v in loss_dict.items()}, grads


def ln_inv_temperature() -> float:
    """model_clip.py:330  logit_scale init = ln(1/0.07)."""
    return math.log(1 / 0.07)


# --------------------------------------------------------------------------- #
# Bounded CPU sample of a large workload (bench.py cpu_baseline / --impl reference)
# --------------------------------------------------------------------------- #


def loss_head_rowblock_step(image_features, text_features, logit_scale, descs_per_image, rows,
                            entitytxt_vec=None, object_vec=None, entitytxt_num=None, object_num=None):
    """fwd+bwd of the reference loss head for a BLOCK of ``rows`` images of a larger batch.

    Per-image work is what the full batch costs per image: the block's images are scored against
    ALL B*T descriptions (image-side CE, model_clip.py:506-508,648) and the block's positive
    descriptions against ALL B images (text-side CE, model_clip.py:504,655-659), with the
    reference's own op sequence (normalise, exp, matmul, CrossEntropyLoss, autograd backward);
    the OT criterion runs on the block's samples.  Returns the summed loss (a float).
    """
    T = descs_per_image
    lo, hi = rows
    img = image_features.detach().clone().requires_grad_(True)
    txt = text_features.detach().clone().requires_grad_(True)
    ls = logit_scale.detach().clone().requires_grad_(True)
    img_n = img / img.norm(dim=-1, keepdim=True)
    txt_n = txt / txt.norm(dim=-1, keepdim=True)
    s = ls.exp()
    ce = torch.nn.CrossEntropyLoss()
    logits_per_image = s * img_n[lo:hi] @ txt_n.t()                                  # [rows, B*T]
    labels_i = torch.arange(lo, hi) * T
    pos = torch.arange(lo, hi) * T
    logits_per_text = s * txt_n.index_select(0, pos) @ img_n.t()                     # [rows, B]
    labels_t = torch.arange(lo, hi)
    total = ce(logits_per_image, labels_i) + ce(logits_per_text, labels_t)
    if entitytxt_vec is not None:
        e = entitytxt_vec[lo:hi].detach().clone().requires_grad_(True)
        o = object_vec[lo:hi].detach().clone().requires_grad_(True)
        total = total + alignment_criterion(e, o, entitytxt_num[lo:hi], object_num[lo:hi])["loss_ot"]
    total.backward()
    return float(total.detach())
