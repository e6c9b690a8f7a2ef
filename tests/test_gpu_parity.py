"""Parity of the CUDA loss head (through the Python mirrors -> C ABI) against the oracle and the
golden vectors generated from the unmodified reference.  Needs a B200: run with -m gpu.

Tolerances (BASELINE.json north_star): fp32 mode loss 1e-5 relative (+2e-6 absolute for a CE loss,
whose fp32 resolution is ulp(14.3) ~ 1e-6 however small the loss is); bf16 mode loss 2e-3,
gradients 1e-2 (relative L2).  fp32-mode gradients are held to 2e-5 relative L2.
"""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err
import clip_event_b200 as ce
from clip_event_b200 import functional as F_
from clip_event_b200 import synthetic as syn
from oracle import clip_event_oracle as orc

pytestmark = pytest.mark.gpu

F32_LOSS_RTOL, F32_LOSS_ATOL, F32_GRAD = 1e-5, 2e-6, 2e-5
BF16_LOSS_RTOL, BF16_GRAD = 2e-3, 1e-2


def close(a, b, rtol, atol=0.0):
    return abs(float(a) - float(b)) <= rtol * abs(float(b)) + atol


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def run_contrastive(img, txt, ls, lpi, lpt, idx, g_i=1.0, g_t=1.0):
    """fp32 losses (functional layer, no output cast) + grads from the CUDA path."""
    ig, tg = img.cuda().requires_grad_(True), txt.cuda().requires_grad_(True)
    lsg = ls.cuda().requires_grad_(True)
    li, lt = F_.contrastive_over_batch(ig, tg, lsg, lpi.cuda(), lpt.cuda(), idx.cuda())
    (g_i * li + g_t * lt).backward()
    torch.cuda.synchronize()
    return li.item(), lt.item(), ig.grad.cpu(), tg.grad.cpu(), lsg.grad.item()


def run_ot(txt, obj, tnum, onum, scale=1.0):
    tg, og = txt.cuda().requires_grad_(True), obj.cuda().requires_grad_(True)
    loss, dist = F_.ot_alignment(tg, og, tnum.cuda(), onum.cuda())
    (scale * loss).backward()
    torch.cuda.synchronize()
    return loss.item(), dist.cpu(), tg.grad.cpu(), og.grad.cpu()


# ------------------------------------------------------------------------------------------
# golden vectors (reference outputs)
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["contrastive_small_iid", "contrastive_small_trained",
                                  "contrastive_small_randlabels"])
def test_contrastive_golden_full(name):
    g = load_golden(name)
    li, lt, dimg, dtxt, dls = run_contrastive(_t(g["image_features"]), _t(g["text_features"]),
                                              torch.tensor(syn.LOGIT_SCALE_INIT), _t(g["labels_per_image"]),
                                              _t(g["labels_per_text"]), _t(g["index_pos"]))
    assert close(li, g["loss_i"], F32_LOSS_RTOL, F32_LOSS_ATOL)
    assert close(lt, g["loss_t"], F32_LOSS_RTOL, F32_LOSS_ATOL)
    assert rel_err(dimg, g["dimg"]) < F32_GRAD
    assert rel_err(dtxt, g["dtxt"]) < F32_GRAD
    assert close(dls, g["dlogit_scale"], 1e-4, 1e-5)


@pytest.mark.parametrize("name,kind", [("contrastive_c1_iid", "iid"), ("contrastive_c1_trained", "trained"),
                                       ("contrastive_c2_trained", "trained")])
def test_contrastive_golden_baseline_shapes(name, kind):
    g = load_golden(name)
    B, T, D, seed = int(g["B"]), int(g["T"]), int(g["D"]), int(g["seed"])
    img, txt, ls = syn.contrastive_inputs(B, T, D, seed, kind)
    li, lt, dimg, dtxt, dls = run_contrastive(img, txt, ls, _t(g["labels_per_image"]),
                                              _t(g["labels_per_text"]), _t(g["index_pos"]))
    assert close(li, g["loss_i"], F32_LOSS_RTOL, F32_LOSS_ATOL)
    assert close(lt, g["loss_t"], F32_LOSS_RTOL, F32_LOSS_ATOL)
    assert rel_err(dimg[:4], g["dimg_head"]) < F32_GRAD
    assert rel_err(dtxt[:8], g["dtxt_head"]) < F32_GRAD
    assert close(dimg.norm().item(), g["dimg_norm"], 1e-5)
    assert close(dtxt.norm().item(), g["dtxt_norm"], 1e-5)


@pytest.mark.parametrize("name", ["ot_small_full", "ot_small_edge", "ot_small_scattered", "ot_small_correlated"])
def test_ot_golden_full(name):
    g = load_golden(name)
    loss, dist, dtxt, dobj = run_ot(_t(g["entitytxt_vec"]), _t(g["object_vec"]), _t(g["entitytxt_num"]),
                                    _t(g["object_num"]))
    assert close(loss, g["loss_ot"], F32_LOSS_RTOL, 1e-12)
    assert rel_err(dist, g["dist"]) < F32_LOSS_RTOL
    assert rel_err(dtxt, g["dtxt"]) < F32_GRAD
    assert rel_err(dobj, g["dobj"]) < F32_GRAD
    assert torch.isfinite(dtxt).all() and torch.isfinite(dobj).all()
    assert (dobj[:, 0] == 0).all()


@pytest.mark.parametrize("name,masks,kind", [("ot_c1_full", "full", "iid"), ("ot_c1_ragged", "ragged", "iid"),
                                             ("ot_c2_edge", "edge", "iid"), ("ot_c2_correlated", "full", "correlated"),
                                             ("ot_c4_ragged", "ragged", "iid"), ("ot_c5_corner", "full", "iid")])
def test_ot_golden_baseline_shapes(name, masks, kind):
    g = load_golden(name)
    B, M, N, D, seed = (int(g[k]) for k in ("B", "M", "N", "D", "seed"))
    txt, obj, tnum, onum = syn.ot_inputs(B, M, N, D, seed, masks, kind)
    loss, dist, dtxt, dobj = run_ot(txt, obj, tnum, onum)
    assert close(loss, g["loss_ot"], F32_LOSS_RTOL)
    assert rel_err(dist, g["dist"]) < F32_LOSS_RTOL
    assert rel_err(dtxt[:2, :4], g["dtxt_head"]) < F32_GRAD
    assert rel_err(dobj[:2, :6], g["dobj_head"]) < F32_GRAD
    assert close(dtxt.norm().item(), g["dtxt_norm"], 2e-5)
    assert close(dobj.norm().item(), g["dobj_norm"], 2e-5)


# ------------------------------------------------------------------------------------------
# oracle on seeded inputs, both modes
# ------------------------------------------------------------------------------------------
CONTRASTIVE_CASES = [(6, 3, 32, "iid"), (9, 4, 40, "trained"), (32, 5, 512, "trained"),
                     (256, 9, 512, "trained"), (256, 9, 512, "iid"), (130, 7, 768, "trained"),
                     (1024, 9, 768, "trained")]


@pytest.mark.parametrize("B,T,D,kind", CONTRASTIVE_CASES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_contrastive_vs_oracle(B, T, D, kind, dtype):
    img, txt, ls = syn.contrastive_inputs(B, T, D, 11, kind, dtype=dtype)
    lpi, lpt, idx = syn.contrastive_labels(B, T)
    ri, rt, rdi, rdt, rdls = orc.contrastive_closed_form(img.double(), txt.double(), ls.double(), lpi, lpt, idx,
                                                         g_i=1.0, g_t=0.7)
    li, lt, dimg, dtxt, dls = run_contrastive(img, txt, ls, lpi, lpt, idx, 1.0, 0.7)
    if dtype == torch.float32:
        assert close(li, ri, F32_LOSS_RTOL, F32_LOSS_ATOL) and close(lt, rt, F32_LOSS_RTOL, F32_LOSS_ATOL)
        assert rel_err(dimg, rdi) < F32_GRAD and rel_err(dtxt, rdt) < F32_GRAD
        assert close(dls, rdls, 1e-4, 1e-5)
    else:
        assert close(li, ri, BF16_LOSS_RTOL, 1e-4) and close(lt, rt, BF16_LOSS_RTOL, 1e-4)
        assert rel_err(dimg, rdi) < BF16_GRAD and rel_err(dtxt, rdt) < BF16_GRAD
        assert close(dls, rdls, 1e-2, 1e-3)


def test_contrastive_arbitrary_labels_and_index_pos():
    B, T, D = 40, 4, 64
    img, txt, ls = syn.contrastive_inputs(B, T, D, 3, "iid")
    g = torch.Generator().manual_seed(5)
    lpi = torch.randint(0, B * T, (B,), generator=g)
    lpt = torch.randint(0, B, (B * T,), generator=g)
    idx = torch.randperm(B * T, generator=g)[:B].sort().values
    ri, rt, rdi, rdt, rdls = orc.contrastive_closed_form(img.double(), txt.double(), ls.double(), lpi, lpt, idx)
    li, lt, dimg, dtxt, dls = run_contrastive(img, txt, ls, lpi, lpt, idx)
    assert close(li, ri, F32_LOSS_RTOL, F32_LOSS_ATOL) and close(lt, rt, F32_LOSS_RTOL, F32_LOSS_ATOL)
    assert rel_err(dimg, rdi) < F32_GRAD and rel_err(dtxt, rdt) < F32_GRAD


OT_CASES = [(4, 4, 7, 16, "full", "iid"), (6, 5, 9, 24, "edge", "iid"), (5, 6, 11, 16, "scattered", "iid"),
            (32, 8, 50, 512, "ragged", "iid"), (16, 16, 50, 512, "full", "correlated"),
            (8, 32, 257, 768, "ragged", "iid"), (2, 64, 577, 768, "full", "iid"), (3, 20, 300, 64, "ragged", "iid"),
            (3, 64, 130, 64, "ragged", "correlated"), (3, 32, 577, 64, "edge", "iid"), (3, 4, 197, 64, "ragged", "iid")]


@pytest.mark.parametrize("B,M,N,D,masks,kind", OT_CASES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_ot_vs_oracle(B, M, N, D, masks, kind, dtype):
    txt, obj, tnum, onum = syn.ot_inputs(B, M, N, D, 13, masks, kind, dtype=dtype)
    tp, ip = tnum == 0, onum[:, 1:] == 0
    d_ref, dx_ref, dy_ref = orc.ot_closed_form_grads(txt.double(), obj.double()[:, 1:], tp, ip,
                                                     torch.full((B,), 0.01, dtype=torch.float64))
    loss, dist, dtxt, dobj = run_ot(txt, obj, tnum, onum)
    if dtype == torch.float32:
        assert close(loss, 0.01 * d_ref.sum().item(), F32_LOSS_RTOL)
        assert rel_err(dist, d_ref) < F32_LOSS_RTOL
        assert rel_err(dtxt, dx_ref) < F32_GRAD and rel_err(dobj[:, 1:], dy_ref) < F32_GRAD
    else:
        assert close(loss, 0.01 * d_ref.sum().item(), BF16_LOSS_RTOL)
        assert rel_err(dtxt, dx_ref) < BF16_GRAD and rel_err(dobj[:, 1:], dy_ref) < BF16_GRAD
    assert (dobj[:, 0] == 0).all() and torch.isfinite(dtxt).all() and torch.isfinite(dobj).all()
    empty = (tp.all(1) | ip.all(1))
    assert (dist[empty] == 0).all()                       # model_ot.py:62: empty node set -> distance 0


def test_ot_upstream_gradient_scaling_and_per_sample_path():
    txt, obj, tnum, onum = syn.ot_inputs(6, 8, 20, 64, 2, "ragged")
    _, _, d1, o1 = run_ot(txt, obj, tnum, onum, scale=1.0)
    _, _, d3, o3 = run_ot(txt, obj, tnum, onum, scale=-2.5)
    assert rel_err(d3, -2.5 * d1) < 1e-6 and rel_err(o3, -2.5 * o1) < 1e-6
    # optimal_transport_dist: per-sample distances with per-sample upstream gradients
    tp, ip = (tnum == 0).cuda(), (onum[:, 1:] == 0).cuda()
    tg = txt.cuda().requires_grad_(True)
    og = obj[:, 1:].contiguous().cuda().requires_grad_(True)
    dist = ce.optimal_transport_dist(tg, og, tp, ip)
    wts = torch.linspace(0.5, 2.0, 6).cuda()
    (dist * wts).sum().backward()
    d_ref, dx_ref, dy_ref = orc.ot_closed_form_grads(txt.double(), obj.double()[:, 1:], tp.cpu(), ip.cpu(), wts.cpu().double())
    assert rel_err(dist.cpu(), d_ref) < F32_LOSS_RTOL
    assert rel_err(tg.grad.cpu(), dx_ref) < F32_GRAD and rel_err(og.grad.cpu(), dy_ref) < F32_GRAD


def test_ot_second_backward_over_a_retained_graph():
    """The reference's autograd allows ``backward(retain_graph=True)`` followed by another backward; the stashed
    gradients belong to autograd after the first one, so the second launches the kernel again: .grad doubles."""
    txt, obj, tnum, onum = syn.ot_inputs(5, 4, 6, 16, 0, "ragged")
    tg, og = txt.cuda().requires_grad_(True), obj.cuda().requires_grad_(True)
    loss, dist = F_.ot_alignment(tg, og, tnum.cuda(), onum.cuda())
    loss.backward(retain_graph=True)
    g1t, g1o = tg.grad.clone(), og.grad.clone()
    (3.0 * loss).backward(retain_graph=True)
    assert rel_err(tg.grad, 4 * g1t) < 1e-6 and rel_err(og.grad, 4 * g1o) < 1e-6
    dist.sum().backward()                               # per-sample path on the third pass: d(sum dist) = grads / 0.01
    assert rel_err(tg.grad, 104 * g1t) < 1e-5 and rel_err(og.grad, 104 * g1o) < 1e-5


# ------------------------------------------------------------------------------------------
# the reference-granularity blocks
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["ot_small_full", "ot_small_edge", "ot_small_scattered"])
def test_ot_blocks_golden(name):
    g = load_golden(name)
    txt, obj = _t(g["entitytxt_vec"]), _t(g["object_vec"])
    tnum, onum = _t(g["entitytxt_num"]), _t(g["object_num"])
    tp, ip = tnum == 0, onum[:, 1:] == 0
    img = obj[:, 1:].contiguous()
    cost = ce.cost_matrix_cosine(txt.cuda(), img.cuda())
    assert rel_err(cost.cpu(), g["cost"]) < 2e-6
    jp = tp.unsqueeze(-1) | ip.unsqueeze(-2)
    cm = _t(g["cost"]).masked_fill(jp, 0).cuda()
    tl = (tp.size(1) - tp.sum(1)).float().cuda()
    il = (ip.size(1) - ip.sum(1)).float().cuda()
    plan = ce.ipot(cm, tl, tp.cuda(), il, ip.cuda(), jp.cuda(), 0.5, 50, 1)
    assert rel_err(plan.cpu(), g["plan"]) < 1e-5
    plan2 = ce.ipot(cm, tl, tp.cuda(), il, ip.cuda(), jp.cuda(), 0.3, 10, 1)
    assert rel_err(plan2.cpu(), g["plan_b03_it10"]) < 1e-5
    tr = ce.trace(cm.matmul(plan))
    assert rel_err(tr.cpu(), g["trace"]) < 1e-5
    dist = ce.optimal_transport_dist(txt.cuda(), img.cuda(), tp.cuda(), ip.cuda())
    assert rel_err(dist.cpu(), g["dist"]) < F32_LOSS_RTOL
    dist_c = ce.optimal_transport_dist(txt.cuda(), img.cuda(), tp.cuda(), ip.cuda(), cost=_t(g["cost"]).cuda())
    assert rel_err(dist_c.cpu(), g["dist"]) < 1e-5


@pytest.mark.parametrize("name", ["contrastive_small_iid", "contrastive_small_trained"])
def test_materialised_logits_golden(name):
    g = load_golden(name)
    head = ce.ClipEventHead().cuda()
    lpi, lpt = head(_t(g["image_features"]).cuda(), _t(g["text_features"]).cuda())
    assert tuple(lpi.shape) == g["logits_per_image"].shape and tuple(lpt.shape) == g["logits_per_text"].shape
    assert rel_err(lpi.materialize().cpu(), g["logits_per_image"]) < 1e-5
    assert rel_err(lpt.materialize().cpu(), g["logits_per_text"]) < 1e-5
    # preprocess_description_contrastive.py:129-131 style consumer
    probs = lpi.softmax(dim=-1)
    assert rel_err(probs.cpu(), _t(g["logits_per_image"]).softmax(-1)) < 1e-4


# ------------------------------------------------------------------------------------------
# the drop-in modules as engine.py drives them, BASELINE shapes, plus size-independent properties
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("wl,dtype", [("c1", torch.float32), ("c2", torch.float32), ("c2", torch.bfloat16)])
def test_engine_style_step(wl, dtype):
    w = syn.WORKLOADS[wl]
    img, txt, ls = syn.contrastive_inputs(w.B, w.T, w.D, 0, "trained", dtype=dtype)
    lpi, lpt, idx = syn.contrastive_labels(w.B, w.T)
    etxt, obj, tnum, onum = syn.ot_inputs(w.B, w.M, w.N, w.D, 0, "ragged", dtype=dtype)
    ref_losses, ref_grads = orc.loss_head_step(img.float(), txt.float(), ls, lpi, lpt, idx, etxt.float(),
                                               obj.float(), tnum, onum)
    head = ce.ClipEventHead().cuda()
    criterion, criterion_ot = ce.CriterionContrastive("ce"), ce.CriterionAlignment()
    leaves = {k: v.cuda().requires_grad_(True) for k, v in dict(img=img, txt=txt, etxt=etxt, obj=obj).items()}
    a, b = head(leaves["img"], leaves["txt"])                                               # engine.py:48
    loss_dict = criterion(a, b, lpi.cuda(), lpt.cuda(), index_pos=idx.cuda(),
                          constrastive_overbatch=head.constrastive_overbatch)               # engine.py:52-53
    loss_dict.update(criterion_ot(leaves["etxt"], leaves["obj"], tnum.cuda(), onum.cuda()))  # engine.py:63
    losses = sum(loss for loss in loss_dict.values())                                       # engine.py:67
    losses.backward()                                                                       # engine.py:88
    torch.cuda.synchronize()
    lt, gt = (F32_LOSS_RTOL, F32_GRAD) if dtype == torch.float32 else (4e-3, BF16_GRAD)   # bf16 OUTPUT cast: 2^-8
    for k in ("loss_i", "loss_t", "loss_ot"):
        assert loss_dict[k].dtype == dtype
        assert close(loss_dict[k].item(), ref_losses[k].item(), lt, 2e-6 if dtype == torch.float32 else 1e-4), k
    for mine, theirs in [("img", "image_features"), ("txt", "text_features"), ("etxt", "entitytxt_vec"),
                         ("obj", "object_vec")]:
        assert rel_err(leaves[mine].grad.cpu(), ref_grads[theirs]) < gt, mine
    assert close(head.logit_scale.grad.item(), ref_grads["logit_scale"].item(), 1e-2 if dtype != torch.float32 else 1e-4, 1e-4)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_full_size_properties(dtype):
    """c4-sized inputs: properties that hold without running the oracle at full size."""
    w = syn.WORKLOADS["c4"]
    B = 256
    img, txt, ls = syn.contrastive_inputs(B, w.T, w.D, 1, "trained", dtype=dtype)
    lpi, lpt, idx = syn.contrastive_labels(B, w.T)
    li, lt, dimg, dtxt, dls = run_contrastive(img, txt, ls, lpi, lpt, idx)
    tol = 1e-4 if dtype == torch.float32 else 3e-2
    # the loss is invariant to the scale of each embedding => gradient orthogonal to the embedding
    cosi = (dimg.double() * img.double()).sum(1) / (dimg.double().norm(dim=1) * img.double().norm(dim=1) + 1e-30)
    cost = (dtxt.double() * txt.double()).sum(1) / (dtxt.double().norm(dim=1) * txt.double().norm(dim=1) + 1e-30)
    assert cosi.abs().max() < tol and cost.abs().max() < tol
    assert li >= 0 and lt >= 0
    # permuting the samples of the OT batch permutes the distances and leaves the loss unchanged
    etxt, obj, tnum, onum = syn.ot_inputs(64, w.M, w.N, w.D, 2, "ragged", dtype=dtype)
    loss, dist, dt_, do_ = run_ot(etxt, obj, tnum, onum)
    perm = torch.randperm(64, generator=torch.Generator().manual_seed(0))
    loss_p, dist_p, dtp, dop = run_ot(etxt[perm], obj[perm], tnum[perm], onum[perm])
    assert torch.equal(dist_p, dist[perm]) and torch.equal(dtp, dt_[perm]) and torch.equal(dop, do_[perm])
    assert close(loss_p, loss, 1e-6)
    # padded nodes receive exactly zero gradient; scale invariance of the cosine cost again
    assert (dt_[tnum == 0] == 0).all() and (do_[onum == 0] == 0).all()
    valid = tnum == 1
    c = (dt_.double() * etxt.double()).sum(-1)[valid] / (dt_.double().norm(dim=-1)[valid] * etxt.double().norm(dim=-1)[valid] + 1e-30)
    assert c.abs().max() < (1e-3 if dtype == torch.float32 else 5e-2)
    assert (dist >= -1e-6).all()


def test_error_behaviour_on_gpu():
    head = ce.ClipEventHead().cuda()
    crit = ce.CriterionContrastive("ce")
    a, b = head(torch.randn(4, 16, device="cuda"), torch.randn(8, 16, device="cuda"))
    with pytest.raises(RuntimeError):
        crit(a, b, index_pos=None)
    with pytest.raises(RuntimeError):     # fp16 is not a supported embedding dtype
        F_.contrastive_over_batch(torch.randn(4, 16, device="cuda").half(), torch.randn(8, 16, device="cuda").half(),
                                  torch.tensor(1.0, device="cuda"), torch.arange(4).cuda(), torch.arange(8).cuda(),
                                  torch.arange(4).cuda())
    with pytest.raises(RuntimeError, match="multiple of 8"):
        F_.contrastive_over_batch(torch.randn(4, 12, device="cuda"), torch.randn(8, 12, device="cuda"),
                                  torch.tensor(1.0, device="cuda"), torch.arange(4).cuda(), torch.arange(8).cuda() // 2,
                                  torch.arange(4).cuda() * 2)
    with pytest.raises(RuntimeError):     # more than 64 text nodes
        ce.CriterionAlignment()(torch.randn(2, 65, 16, device="cuda"), torch.randn(2, 9, 16, device="cuda"),
                                torch.ones(2, 65, dtype=torch.int64, device="cuda"),
                                torch.ones(2, 9, dtype=torch.int64, device="cuda"))


# ------------------------------------------------------------------------------------------
# over-instance image side: 'ce' and 'bce'  (model_clip.py:509-520, 624-651)
# ------------------------------------------------------------------------------------------
def run_instance(img, txt, ls, lpi, lpt, idx, loss):
    head = ce.ClipEventHead(constrastive_overbatch=False).cuda()
    with torch.no_grad():
        head.logit_scale.copy_(ls)
    ig, tg = img.cuda().requires_grad_(True), txt.cuda().requires_grad_(True)
    a, b = head(ig, tg)
    out = ce.CriterionContrastive(loss)(a, b, lpi.cuda(), lpt.cuda(), index_pos=idx.cuda(),
                                        constrastive_overbatch=False)
    (out["loss_i"].float() + out["loss_t"].float()).backward()
    torch.cuda.synchronize()
    return out["loss_i"].item(), out["loss_t"].item(), ig.grad.cpu(), tg.grad.cpu(), head.logit_scale.grad.item(), a


@pytest.mark.parametrize("name,loss", [("contrastive_small_instance", "ce"), ("contrastive_small_bce", "bce")])
def test_contrastive_instance_golden_full(name, loss):
    g = load_golden(name)
    li, lt, dimg, dtxt, dls, lazy = run_instance(_t(g["image_features"]), _t(g["text_features"]),
                                                 torch.tensor(syn.LOGIT_SCALE_INIT), _t(g["labels_per_image"]),
                                                 _t(g["labels_per_text"]), _t(g["index_pos"]), loss)
    assert tuple(lazy.shape) == g["logits_per_image"].shape
    assert rel_err(lazy.materialize().cpu(), g["logits_per_image"]) < 1e-5
    assert close(li, g["loss_i"], F32_LOSS_RTOL, F32_LOSS_ATOL)
    assert close(lt, g["loss_t"], F32_LOSS_RTOL, F32_LOSS_ATOL)
    assert rel_err(dimg, g["dimg"]) < F32_GRAD
    assert rel_err(dtxt, g["dtxt"]) < F32_GRAD
    assert close(dls, g["dlogit_scale"], 1e-4, 1e-5)


@pytest.mark.parametrize("loss", ["ce", "bce"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_contrastive_instance_vs_oracle_c2(loss, dtype):
    B, T, D = 256, 9, 512
    img, txt, ls = syn.contrastive_inputs(B, T, D, 17, "trained", dtype=dtype)
    lpi, lpt, idx = syn.contrastive_labels(B, T, overbatch=False)
    if loss == "bce":
        lpi = torch.zeros(B, T)
        lpi[:, 0] = 1.0
    ref_losses, ref_grads = orc.loss_head_step(img.double(), txt.double(), ls.double(), lpi.double() if loss == "bce" else lpi,
                                               lpt, idx, overbatch=False, kind=loss)
    li, lt, dimg, dtxt, dls, _ = run_instance(img, txt, ls, lpi, lpt, idx, loss)
    if dtype == torch.float32:
        assert close(li, ref_losses["loss_i"].item(), F32_LOSS_RTOL, F32_LOSS_ATOL)
        assert close(lt, ref_losses["loss_t"].item(), F32_LOSS_RTOL, F32_LOSS_ATOL)
        assert rel_err(dimg, ref_grads["image_features"]) < F32_GRAD
        assert rel_err(dtxt, ref_grads["text_features"]) < F32_GRAD
        assert close(dls, ref_grads["logit_scale"].item(), 1e-4, 1e-5)
    else:
        assert close(li, ref_losses["loss_i"].item(), 4e-3, 1e-4) and close(lt, ref_losses["loss_t"].item(), 4e-3, 1e-4)
        assert rel_err(dimg, ref_grads["image_features"]) < BF16_GRAD
        assert rel_err(dtxt, ref_grads["text_features"]) < BF16_GRAD


def test_criterion_mode_errors():
    head = ce.ClipEventHead(constrastive_overbatch=True).cuda()
    a, b = head(torch.randn(4, 16, device="cuda"), torch.randn(8, 16, device="cuda"))
    idx = torch.arange(4, device="cuda") * 2
    with pytest.raises(RuntimeError, match="constrastive_overbatch=false"):
        ce.CriterionContrastive("bce")(a, b, index_pos=idx, constrastive_overbatch=True)
    with pytest.raises(RuntimeError, match="kl"):
        ce.CriterionContrastive("kl")(a, b, index_pos=idx)
    with pytest.raises(RuntimeError, match="does not match"):
        ce.CriterionContrastive("ce")(a, b, index_pos=idx, constrastive_overbatch=False)


# ------------------------------------------------------------------------------------------
# the tcgen05 GEMM engine by itself, and the loss head at the benchmark's size
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("entry", ["ce_debug_gemm", "ce_debug_gemm_pair"])
@pytest.mark.parametrize("M,N,K,a_mn,b_mn,split_k", [
    (256, 256, 64, 0, 0, 1), (100, 200, 72, 0, 0, 1), (1024, 2304, 512, 0, 0, 1), (256, 256, 2048, 0, 0, 4),
    (256, 256, 64, 1, 0, 1), (256, 256, 64, 0, 1, 1), (520, 512, 1000, 0, 1, 1), (2304, 512, 1024, 1, 1, 3),
    (4608, 512, 640, 1, 1, 1),      # tall A, two column blocks: column-fastest tile order
    (384, 256, 128, 0, 0, 1),       # odd number of 128-row blocks: the pair's second CTA idles on the last unit
])
def test_tcgen05_gemm_engine_bit_exact(entry, M, N, K, a_mn, b_mn, split_k):
    """C = A B^t through the TMA + tcgen05 main loop (single-CTA and CTA-pair flavours), K-major
    and MN-major operands, split-K: small-integer inputs make every product and sum exact."""
    from clip_event_b200 import _lib as L
    lib = L.load()
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    A = torch.randint(-3, 4, (M, K), generator=g).float()
    B = torch.randint(-3, 4, (N, K), generator=g).float()
    Ad = (A.t().contiguous() if a_mn else A).bfloat16().cuda()
    Bd = (B.t().contiguous() if b_mn else B).bfloat16().cuda()
    C = torch.full((M, N), float("nan"), device="cuda")
    rc = getattr(lib, entry)(Ad.data_ptr(), Bd.data_ptr(), C.data_ptr(), M, N, K, L.dtype_code(torch.bfloat16),
                             a_mn, b_mn, split_k, torch.cuda.current_stream().cuda_stream)
    assert rc == 0, L.last_error()
    torch.cuda.synchronize()
    assert torch.equal(C.cpu(), A @ B.t())


@pytest.mark.parametrize("B,T,D", [(4096, 9, 512), (2000, 9, 520)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_benchmark_size_vs_torch_fp32(dtype, B, T, D):
    """The bench workload (c3: 4096 images x 36864 descriptions, D = 512) against the same maths in
    plain PyTorch fp32 on the GPU -- the only test large enough to reach the CTA-pair GEMMs, the
    column-fastest tile order and several waves of the persistent kernels.  The second shape hits
    the same paths with ragged edges in every dimension (rows, columns, a partial k-block)."""
    img, txt, ls = syn.contrastive_inputs(B, T, D, 3, "trained", dtype=dtype)
    lpi, lpt, idx = syn.contrastive_labels(B, T)
    li, lt, dimg, dtxt, dls = run_contrastive(img, txt, ls, lpi, lpt, idx)
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        i32 = img.float().cuda().requires_grad_(True)
        t32 = txt.float().cuda().requires_grad_(True)
        l32 = ls.float().cuda().requires_grad_(True)
        logits = l32.exp() * (i32 / i32.norm(dim=1, keepdim=True)) @ (t32 / t32.norm(dim=1, keepdim=True)).t()
        ref_i = torch.nn.functional.cross_entropy(logits, lpi.cuda())
        ref_t = torch.nn.functional.cross_entropy(logits.t()[idx.cuda()], lpt.cuda()[idx.cuda()])
        (ref_i + ref_t).backward()
        torch.cuda.synchronize()
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    if dtype == torch.float32:
        lr, la, gt = 2e-5, 4e-6, 5e-5        # the fp32 reference itself carries ~1e-6 at this size
    else:
        lr, la, gt = BF16_LOSS_RTOL, 0.0, BF16_GRAD
    assert close(li, ref_i.item(), lr, la) and close(lt, ref_t.item(), lr, la)
    assert rel_err(dimg, i32.grad.cpu()) < gt
    assert rel_err(dtxt, t32.grad.cpu()) < gt
    assert close(dls, l32.grad.item(), 1e-2 if dtype != torch.float32 else 2e-4, 1e-4)


def test_integration_md_ctypes_stubs_run():
    """The two ctypes stubs printed in INTEGRATION.md (route B) are executed verbatim against the
    built library and must agree with the shipped Python mirror."""
    import os
    import re
    from clip_event_b200 import _lib as L
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    ns = {}
    for blk in blocks:
        if "def ot_loss_and_grads" in blk or "def contrastive_losses_and_grads" in blk:
            exec(blk.replace('"libclip_event_b200.so"', repr(L.LIB_PATH)), ns)
    assert "ot_loss_and_grads" in ns and "contrastive_losses_and_grads" in ns
    w = syn.WORKLOADS["c2"]
    for dtype in (torch.float32, torch.bfloat16):
        img, txt, ls = syn.contrastive_inputs(w.B, w.T, w.D, 5, "trained", dtype=dtype)
        lpi, lpt, idx = syn.contrastive_labels(w.B, w.T)
        li, lt, dimg, dtxt, dls = run_contrastive(img, txt, ls, lpi, lpt, idx)
        a, b, gi, gt, gl = ns["contrastive_losses_and_grads"](img.cuda(), txt.cuda(), ls.float().reshape(1).cuda(),
                                                             lpi.cuda(), lpt.cuda(), idx.cuda())
        torch.cuda.synchronize()
        assert a.item() == li and b.item() == lt
        # split-K partial sums meet in red.add, so gradients repeat to rounding, not bit for bit (a few bf16
        # elements per 10^5 flip by one unit in the last place between two runs)
        assert rel_err(gi, dimg) < 1e-4 and rel_err(gt, dtxt) < 1e-4 and close(gl.item(), dls, 1e-5, 1e-7)
        etxt, obj, tnum, onum = syn.ot_inputs(32, w.M, w.N, w.D, 6, "ragged", dtype=dtype)
        loss, dist, dt_, do_ = run_ot(etxt, obj, tnum, onum)
        l2, d2, dt2, do2 = ns["ot_loss_and_grads"](etxt.cuda(), obj.cuda(), tnum.cuda(), onum.cuda())
        torch.cuda.synchronize()
        assert l2.item() == loss and torch.equal(d2.cpu(), dist)
        assert torch.equal(dt2.cpu(), dt_) and torch.equal(do2.cpu()[:, 1:], do_[:, 1:])
