"""Seeded synthetic inputs for the loss head.

The reference's training data is private, so every test and the benchmark drive the
loss head with tensors of the shapes the reference's collate_fn / encoders produce
(dataset_voa.py:397-399,532-577,607-663; model_clip.py:531-552).  All generators run on
the CPU with an explicit ``torch.Generator`` so the same values reach the oracle and
the CUDA path.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import torch

LOGIT_SCALE_INIT = math.log(1 / 0.07)          # model_clip.py:330


@dataclass
class Workload:
    """Shapes of one BASELINE.json config."""
    name: str
    B: int          # images per (global) batch
    K: int          # hard negatives per image; T = 1 + K descriptions
    D: int          # embedding width: 512 (ViT-B/32) or 768 (ViT-L/14)
    M: int          # text (event / argument) nodes
    N: int          # image nodes AFTER the whole-image slot is dropped
    iters: int = 50

    @property
    def T(self) -> int:
        return 1 + self.K


WORKLOADS = {
    "c1": Workload("c1", B=32, K=4, D=512, M=8, N=50),
    "c2": Workload("c2", B=256, K=8, D=512, M=16, N=50),
    "c3": Workload("c3", B=4096, K=8, D=512, M=16, N=50),
    "c4": Workload("c4", B=1024, K=8, D=768, M=32, N=257),
    "c5": Workload("c5", B=512, K=8, D=768, M=64, N=577),
}


def contrastive_labels(B: int, T: int, overbatch: bool = True):
    """labels_per_image, labels_per_text, index_pos as dataset_voa.py:617-663 builds them."""
    ar = torch.arange(B, dtype=torch.int64)
    lpi = ar * T if overbatch else torch.zeros(B, dtype=torch.int64)
    lpt = ar.repeat_interleave(T)
    idx = ar * T
    return lpi, lpt, idx


def contrastive_inputs(B: int, T: int, D: int, seed: int = 0, kind: str = "iid",
                       dtype=torch.float32):
    """image_features [B,D], text_features [B*T,D], logit_scale [].

    kind='iid'     : N(0,1) everywhere (softmax ~ uniform, loss ~ ln(B*T))
    kind='trained' : txt_pos = img + 0.5 eps, hard negatives = txt_pos + 0.3 eps_k (peaked softmax)
    Rows of image b's descriptions are b*T .. b*T+T-1, positive first (dataset_voa.py:607-611).
    """
    g = torch.Generator().manual_seed(seed)
    img = torch.randn(B, D, generator=g)
    if kind == "iid":
        txt = torch.randn(B * T, D, generator=g)
    elif kind == "trained":
        pos = img + 0.5 * torch.randn(B, D, generator=g)
        neg = pos.unsqueeze(1) + 0.3 * torch.randn(B, T - 1, D, generator=g)
        txt = torch.cat([pos.unsqueeze(1), neg], 1).reshape(B * T, D)
    else:
        raise ValueError(kind)
    ls = torch.tensor(LOGIT_SCALE_INIT)
    if dtype != torch.float32:
        img, txt = img.to(dtype), txt.to(dtype)
    return img, txt, ls


def ot_inputs(B: int, M: int, N: int, D: int, seed: int = 0, masks: str = "full",
              kind: str = "iid", dtype=torch.float32):
    """entitytxt_vec [B,M,D], object_vec [B,N+1,D], entitytxt_num [B,M], object_num [B,N+1].

    object_vec keeps the whole-image slot 0 (always valid) that CriterionAlignment drops
    (model_clip.py:686).  masks: 'full' | 'ragged' (prefix-ones, lengths U{1..}) | 'edge'
    (sample 0 has no text nodes, sample 1 no image nodes, sample 2 a zero text vector,
    rest ragged) | 'scattered' (arbitrary 0/1 pattern, at least one valid each).
    kind='correlated': x = 0.3 eps + base_b  (cost range ~0..0.27, stresses precision).
    """
    g = torch.Generator().manual_seed(seed)
    txt = torch.randn(B, M, D, generator=g)
    img = torch.randn(B, N + 1, D, generator=g)
    if kind == "correlated":
        base = torch.randn(B, 1, D, generator=g)
        txt = 0.3 * txt + base
        img = 0.3 * img + base
    elif kind != "iid":
        raise ValueError(kind)
    tnum = torch.ones(B, M, dtype=torch.int64)
    onum = torch.ones(B, N + 1, dtype=torch.int64)
    if masks in ("ragged", "edge"):
        tl = torch.randint(1, M + 1, (B,), generator=g)
        il = torch.randint(1, N + 1, (B,), generator=g)
        tnum = (torch.arange(M).view(1, M) < tl.view(B, 1)).to(torch.int64)
        onum = (torch.arange(N + 1).view(1, N + 1) < (il + 1).view(B, 1)).to(torch.int64)
        if masks == "edge":
            if B > 0:
                tnum[0] = 0
            if B > 1:
                onum[1, 1:] = 0
            if B > 2:
                txt[2, 0] = 0
                tnum[2, 0] = 1
    elif masks == "scattered":
        tnum = (torch.rand(B, M, generator=g) < 0.6).to(torch.int64)
        onum = (torch.rand(B, N + 1, generator=g) < 0.6).to(torch.int64)
        tnum[:, 0] = 1
        onum[:, :2] = 1
    elif masks != "full":
        raise ValueError(masks)
    if dtype != torch.float32:
        txt, img = txt.to(dtype), img.to(dtype)
    return txt, img, tnum, onum
