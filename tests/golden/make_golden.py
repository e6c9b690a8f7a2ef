"""Generate golden vectors by running the UNMODIFIED reference (limanling/clip-event).

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

It imports ``model_clip`` / ``model_ot`` from /root/reference/src/clip-event, drives them
with the seeded synthetic inputs of ``clip_event_b200.synthetic`` and writes
``tests/golden/golden_*.npz``.  Small cases store inputs and every output in full; the
BASELINE-sized cases store the losses, per-sample distances, gradient norms and a few
gradient rows plus an input checksum (the inputs are regenerated from the seed).

The reference has no tests or fixtures of its own, so these files
are what pins the oracle and the CUDA path to the reference's behaviour.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("CLIP_EVENT_REFERENCE", "/root/reference/src/clip-event")
sys.path.insert(0, REF)

import model_clip as ref_clip  # noqa: E402  (the reference, unmodified)
import model_ot as ref_ot      # noqa: E402

from clip_event_b200 import synthetic as syn  # noqa: E402

torch.set_num_threads(8)


def ref_head(img, txt, ls, overbatch):
    """The 10 lines of CLIP.forward after the encoders (model_clip.py:496-520), run through a
    stub CLIP object so the reference's own code executes."""
    class Stub(ref_clip.CLIP):
        def __init__(self):
            torch.nn.Module.__init__(self)
            self.logit_scale = ls
            self.constrastive_overbatch = overbatch

        def encode_image(self, x, use_grid=False):
            return x

        def encode_text(self, x):
            return x
    return Stub()(img, txt)


def contrastive_case(B, T, D, seed, kind, overbatch=True, loss="ce", labels="canonical", full=True):
    img, txt, ls = syn.contrastive_inputs(B, T, D, seed, kind)
    lpi, lpt, idx = syn.contrastive_labels(B, T, overbatch)
    g = torch.Generator().manual_seed(seed + 1000)
    if labels == "random":
        lpi = torch.randint(0, B * T if overbatch else T, (B,), generator=g)
        lpt = torch.randint(0, B, (B * T,), generator=g)
        idx = torch.randperm(B * T, generator=g)[:B].sort().values
    if loss == "bce":
        lpi = torch.zeros(B, T)
        lpi[:, 0] = 1.0
    img.requires_grad_(True)
    txt.requires_grad_(True)
    ls = ls.clone().requires_grad_(True)
    lpi_logits, lpt_logits = ref_head(img, txt, ls, overbatch)
    crit = ref_clip.CriterionContrastive(loss)
    out = crit(lpi_logits, lpt_logits, lpi, lpt, index_pos=idx, constrastive_overbatch=overbatch)
    (out["loss_i"] + out["loss_t"]).backward()
    rec = dict(B=B, T=T, D=D, seed=seed, overbatch=int(overbatch),
               labels_per_image=lpi.numpy(), labels_per_text=lpt.numpy(), index_pos=idx.numpy(),
               loss_i=out["loss_i"].item(), loss_t=out["loss_t"].item(),
               dlogit_scale=ls.grad.item(),
               dimg_norm=img.grad.norm().item(), dtxt_norm=txt.grad.norm().item(),
               dimg_head=img.grad[:4].numpy(), dtxt_head=txt.grad[:8].numpy(),
               in_checksum=float(img.detach().double().sum() + txt.detach().double().sum()))
    if full:
        rec.update(image_features=img.detach().numpy(), text_features=txt.detach().numpy(),
                   logits_per_image=lpi_logits.detach().numpy(),
                   logits_per_text=lpt_logits.detach().numpy(),
                   dimg=img.grad.numpy(), dtxt=txt.grad.numpy())
    return rec


def ot_case(B, M, N, D, seed, masks, kind="iid", full=True):
    txt, obj, tnum, onum = syn.ot_inputs(B, M, N, D, seed, masks, kind)
    txt.requires_grad_(True)
    obj.requires_grad_(True)
    crit = ref_clip.CriterionAlignment()
    out = crit(txt, obj, tnum, onum)
    out["loss_ot"].backward()
    with torch.no_grad():
        tp, ip = tnum == 0, onum[:, 1:] == 0
        dist = ref_ot.optimal_transport_dist(txt.detach(), obj.detach()[:, 1:], tp, ip)
    rec = dict(B=B, M=M, N=N, D=D, seed=seed, loss_ot=out["loss_ot"].item(), dist=dist.numpy(),
               dtxt_norm=txt.grad.norm().item(), dobj_norm=obj.grad.norm().item(),
               dtxt_head=txt.grad[:2, :4].numpy(), dobj_head=obj.grad[:2, :6].numpy(),
               in_checksum=float(txt.detach().double().sum() + obj.detach().double().sum()))
    if full:
        with torch.no_grad():
            cost = ref_ot.cost_matrix_cosine(txt.detach(), obj.detach()[:, 1:])
            jp = tp.unsqueeze(-1) | ip.unsqueeze(-2)
            cm = cost.clone().masked_fill_(jp, 0)
            tl = (M - tp.sum(1)).float()
            il = (N - ip.sum(1)).float()
            T = ref_ot.ipot(cm, tl, tp, il, ip, jp, 0.5, 50, 1)
            T10 = ref_ot.ipot(cm, tl, tp, il, ip, jp, 0.3, 10, 1)  # k>1 raises in the reference (sigma keeps shape [b,1,m])
            tr = ref_ot.trace(cm.matmul(T))
        rec.update(entitytxt_vec=txt.detach().numpy(), object_vec=obj.detach().numpy(),
                   entitytxt_num=tnum.numpy(), object_num=onum.numpy(), cost=cost.numpy(),
                   plan=T.numpy(), plan_b03_it10=T10.numpy(), trace=tr.numpy(),
                   dtxt=txt.grad.numpy(), dobj=obj.grad.numpy())
    return rec


def main():
    cases = {
        # small, stored in full
        "contrastive_small_iid": contrastive_case(6, 3, 32, 0, "iid"),
        "contrastive_small_trained": contrastive_case(8, 4, 48, 1, "trained"),
        "contrastive_small_randlabels": contrastive_case(7, 3, 40, 2, "iid", labels="random"),
        "contrastive_small_instance": contrastive_case(6, 5, 32, 3, "trained", overbatch=False),
        "contrastive_small_bce": contrastive_case(6, 5, 32, 4, "trained", overbatch=False, loss="bce"),
        "ot_small_full": ot_case(4, 4, 7, 16, 0, "full"),
        "ot_small_edge": ot_case(6, 5, 9, 24, 1, "edge"),
        "ot_small_scattered": ot_case(5, 6, 11, 16, 2, "scattered"),
        "ot_small_correlated": ot_case(4, 8, 12, 32, 3, "ragged", kind="correlated"),
        # BASELINE shapes, summaries only
        "contrastive_c1_iid": contrastive_case(32, 5, 512, 0, "iid", full=False),
        "contrastive_c1_trained": contrastive_case(32, 5, 512, 1, "trained", full=False),
        "contrastive_c2_trained": contrastive_case(256, 9, 512, 2, "trained", full=False),
        "contrastive_c2_instance": contrastive_case(256, 9, 512, 3, "trained", overbatch=False, full=False),
        "ot_c1_full": ot_case(32, 8, 50, 512, 0, "full", full=False),
        "ot_c1_ragged": ot_case(32, 8, 50, 512, 1, "ragged", full=False),
        "ot_c2_edge": ot_case(64, 16, 50, 512, 2, "edge", full=False),
        "ot_c2_correlated": ot_case(16, 16, 50, 512, 3, "full", kind="correlated", full=False),
        "ot_c4_ragged": ot_case(4, 32, 257, 768, 4, "ragged", full=False),
        "ot_c5_corner": ot_case(2, 64, 577, 768, 5, "full", full=False),
    }
    for name, rec in cases.items():
        path = os.path.join(HERE, "golden_%s.npz" % name)
        np.savez_compressed(path, **{k: np.asarray(v) for k, v in rec.items()})
        print("%-34s %8.1f KB" % (name, os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
