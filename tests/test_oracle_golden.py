"""The oracle (oracle/clip_event_oracle.py) against the golden vectors produced by the
unmodified reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err
from clip_event_b200 import synthetic as syn
from oracle import clip_event_oracle as orc

CONTRASTIVE_FULL = ["contrastive_small_iid", "contrastive_small_trained",
                    "contrastive_small_randlabels", "contrastive_small_instance",
                    "contrastive_small_bce"]
CONTRASTIVE_SUMMARY = [("contrastive_c1_iid", "iid"), ("contrastive_c1_trained", "trained"),
                       ("contrastive_c2_trained", "trained"), ("contrastive_c2_instance", "trained")]
OT_FULL = ["ot_small_full", "ot_small_edge", "ot_small_scattered", "ot_small_correlated"]
OT_SUMMARY = [("ot_c1_full", "full", "iid"), ("ot_c1_ragged", "ragged", "iid"),
              ("ot_c2_edge", "edge", "iid"), ("ot_c2_correlated", "full", "correlated"),
              ("ot_c4_ragged", "ragged", "iid"), ("ot_c5_corner", "full", "iid")]

# fp32 oracle vs fp32 reference: same maths, possibly different op order
TOL = 2e-6
# a CE loss is (logsumexp - picked logit) with |logit| up to s = 14.3, so its fp32 resolution is
# ulp(14.3) ~ 1e-6 however small the loss itself is (peaked 'trained' inputs give loss ~ 1e-4)
ATOL = 2e-6


def close(a, b, rtol=TOL, atol=ATOL):
    return abs(float(a) - float(b)) <= rtol * abs(float(b)) + atol


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


@pytest.mark.parametrize("name", CONTRASTIVE_FULL)
def test_contrastive_full(name):
    g = load_golden(name)
    img, txt = _t(g["image_features"]), _t(g["text_features"])
    ls = torch.tensor(syn.LOGIT_SCALE_INIT)
    over = bool(g["overbatch"])
    kind = "bce" if name.endswith("bce") else "ce"
    lpi, lpt = orc.similarity_logits(img, txt, ls, over)
    assert rel_err(lpi, g["logits_per_image"]) < TOL
    assert rel_err(lpt, g["logits_per_text"]) < TOL
    lab_i = _t(g["labels_per_image"])
    losses, grads = orc.loss_head_step(img, txt, ls, lab_i, _t(g["labels_per_text"]),
                                       _t(g["index_pos"]), overbatch=over, kind=kind)
    assert close(losses["loss_i"].item(), g["loss_i"])
    assert close(losses["loss_t"].item(), g["loss_t"])
    assert rel_err(grads["image_features"], g["dimg"]) < 5 * TOL
    assert rel_err(grads["text_features"], g["dtxt"]) < 5 * TOL
    assert abs(grads["logit_scale"].item() - g["dlogit_scale"]) <= 1e-5 * max(1.0, abs(g["dlogit_scale"]))


@pytest.mark.parametrize("name", [n for n in CONTRASTIVE_FULL if "instance" not in n and "bce" not in n])
def test_contrastive_closed_form_matches_reference(name):
    """The single-GEMM identity the CUDA kernels rely on."""
    g = load_golden(name)
    img, txt = _t(g["image_features"]).double(), _t(g["text_features"]).double()
    ls = torch.tensor(syn.LOGIT_SCALE_INIT, dtype=torch.float64)
    li, lt, dimg, dtxt, dls = orc.contrastive_closed_form(
        img, txt, ls, _t(g["labels_per_image"]), _t(g["labels_per_text"]), _t(g["index_pos"]))
    assert close(li.item(), g["loss_i"])
    assert close(lt.item(), g["loss_t"])
    assert rel_err(dimg, g["dimg"]) < 5 * TOL
    assert rel_err(dtxt, g["dtxt"]) < 5 * TOL
    assert abs(dls.item() - g["dlogit_scale"]) <= 1e-5 * max(1.0, abs(g["dlogit_scale"]))


@pytest.mark.parametrize("name,kind", CONTRASTIVE_SUMMARY)
def test_contrastive_summary(name, kind):
    g = load_golden(name)
    B, T, D, seed = int(g["B"]), int(g["T"]), int(g["D"]), int(g["seed"])
    over = bool(g["overbatch"])
    img, txt, ls = syn.contrastive_inputs(B, T, D, seed, kind)
    assert abs(float(img.double().sum() + txt.double().sum()) - g["in_checksum"]) < 1e-6, "RNG drift"
    losses, grads = orc.loss_head_step(img, txt, ls, _t(g["labels_per_image"]),
                                       _t(g["labels_per_text"]), _t(g["index_pos"]), overbatch=over)
    assert close(losses["loss_i"].item(), g["loss_i"])
    assert close(losses["loss_t"].item(), g["loss_t"])
    assert rel_err(grads["image_features"][:4], g["dimg_head"]) < 5 * TOL
    assert rel_err(grads["text_features"][:8], g["dtxt_head"]) < 5 * TOL
    assert abs(grads["image_features"].norm().item() - g["dimg_norm"]) <= 5 * TOL * g["dimg_norm"]


@pytest.mark.parametrize("name", OT_FULL)
def test_ot_full(name):
    g = load_golden(name)
    txt, obj = _t(g["entitytxt_vec"]), _t(g["object_vec"])
    tnum, onum = _t(g["entitytxt_num"]), _t(g["object_num"])
    tp, ip = tnum == 0, onum[:, 1:] == 0
    img = obj[:, 1:]
    cost = orc.cost_matrix_cosine(txt, img)
    assert rel_err(cost, g["cost"]) < TOL
    jp = tp.unsqueeze(-1) | ip.unsqueeze(-2)
    cm = cost.masked_fill(jp, 0)
    tl = (tp.size(1) - tp.sum(1)).float()
    il = (ip.size(1) - ip.sum(1)).float()
    plan = orc.ipot(cm, tl, tp, il, ip, jp, 0.5, 50, 1)
    assert rel_err(plan, g["plan"]) < 1e-5
    plan2 = orc.ipot(cm, tl, tp, il, ip, jp, 0.3, 10, 1)
    assert rel_err(plan2, g["plan_b03_it10"]) < 1e-5
    assert rel_err(orc.trace_batched(cm.matmul(plan)), g["trace"]) < 1e-5
    dist = orc.optimal_transport_dist(txt, img, tp, ip)
    assert rel_err(dist, g["dist"]) < 1e-5
    losses, grads = orc.loss_head_step(
        torch.randn(2, 8), torch.randn(2, 8), torch.tensor(1.0), torch.arange(2), torch.arange(2),
        torch.arange(2), txt, obj, tnum, onum)
    assert abs(losses["loss_ot"].item() - g["loss_ot"]) <= 1e-5 * abs(g["loss_ot"]) + 1e-12
    assert rel_err(grads["entitytxt_vec"], g["dtxt"]) < 1e-5
    assert rel_err(grads["object_vec"], g["dobj"]) < 1e-5
    # closed-form gradient (what the CUDA kernel implements) == reference autograd
    d, dx, dy = orc.ot_closed_form_grads(txt.double(), img.double(), tp, ip,
                                         torch.full((txt.shape[0],), 0.01, dtype=torch.float64))
    assert rel_err(d, g["dist"]) < 1e-5
    assert rel_err(dx, g["dtxt"]) < 1e-5
    assert rel_err(dy, g["dobj"][:, 1:]) < 1e-5
    assert np.all(g["dobj"][:, 0] == 0), "whole-image slot gets zero grad in the reference"
    assert torch.isfinite(dx).all() and torch.isfinite(dy).all()


@pytest.mark.parametrize("name,masks,kind", OT_SUMMARY)
def test_ot_summary(name, masks, kind):
    g = load_golden(name)
    B, M, N, D, seed = (int(g[k]) for k in ("B", "M", "N", "D", "seed"))
    txt, obj, tnum, onum = syn.ot_inputs(B, M, N, D, seed, masks, kind)
    assert abs(float(txt.double().sum() + obj.double().sum()) - g["in_checksum"]) < 1e-6, "RNG drift"
    tp, ip = tnum == 0, onum[:, 1:] == 0
    dist = orc.optimal_transport_dist(txt, obj[:, 1:], tp, ip)
    assert rel_err(dist, g["dist"]) < 1e-5
    d, dx, dy = orc.ot_closed_form_grads(txt.double(), obj[:, 1:].double(), tp, ip,
                                         torch.full((B,), 0.01, dtype=torch.float64))
    assert rel_err(dx[:2, :4], g["dtxt_head"]) < 2e-5
    assert rel_err(dy[:2, :5], g["dobj_head"][:, 1:]) < 2e-5
    assert abs(dx.norm().item() - g["dtxt_norm"]) <= 2e-5 * g["dtxt_norm"]
    assert abs(0.01 * d.sum().item() - g["loss_ot"]) <= 1e-5 * abs(g["loss_ot"])


def test_reference_rejects_k_gt_1_note():
    """model_ot.py:55-61: with k>1 the reference raises (sigma keeps shape [b,1,m]); the
    oracle implements the intended recurrence, so k>1 has no reference parity claim."""
    C = torch.rand(2, 3, 4)
    pad_x = torch.zeros(2, 3, dtype=torch.bool)
    pad_y = torch.zeros(2, 4, dtype=torch.bool)
    jp = pad_x.unsqueeze(-1) | pad_y.unsqueeze(-2)
    T = orc.ipot(C, torch.full((2,), 3.0), pad_x, torch.full((2,), 4.0), pad_y, jp, 0.5, 5, 2)
    assert T.shape == (2, 4, 3) and torch.isfinite(T).all()


def test_canonical_labels_contract():
    lpi, lpt, idx = orc.canonical_labels(4, 3)
    assert lpi.tolist() == [0, 3, 6, 9]
    assert lpt.tolist() == [0, 0, 0, 1, 1, 1, 2, 2, 2, 3, 3, 3]
    assert idx.tolist() == [0, 3, 6, 9]
    a, b, c = syn.contrastive_labels(4, 3)
    assert a.tolist() == lpi.tolist() and b.tolist() == lpt.tolist() and c.tolist() == idx.tolist()
    lpi0, _, _ = orc.canonical_labels(4, 3, overbatch=False)
    assert lpi0.tolist() == [0, 0, 0, 0]
