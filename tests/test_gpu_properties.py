"""Full-size checks of the CUDA path (BASELINE.json configs c3 / c4 at their own batch sizes): parity against the
fp64 oracle where the oracle finishes in seconds, and the size-independent properties of tests/test_properties.py
where it does not.  Needs a B200: run with -m gpu.

Tolerances as in tests/test_gpu_parity.py (north_star): fp32 loss 1e-5 / gradients 2e-5, bf16 loss 2e-3 / gradients 1e-2.
"""
import math

import pytest
import torch

from conftest import rel_err
import clip_event_b200 as ce
from clip_event_b200 import functional as F_
from clip_event_b200 import synthetic as syn
from oracle import clip_event_oracle as orc

pytestmark = pytest.mark.gpu


def close(a, b, rtol, atol=0.0):
    return abs(float(a) - float(b)) <= rtol * abs(float(b)) + atol


def _run_ot(txt, obj, tnum, onum):
    tg, og = txt.cuda().requires_grad_(True), obj.cuda().requires_grad_(True)
    loss, dist = F_.ot_alignment(tg, og, tnum.cuda(), onum.cuda())
    loss.backward()
    torch.cuda.synchronize()
    return loss.item(), dist.float().cpu(), tg.grad.float().cpu(), og.grad.float().cpu()


def _oracle_ot_chunked(txt, obj, tnum, onum, chunk=128):
    """fp64 closed-form distances + gradients of 0.01 * sum(dist), a chunk of samples at a time (memory)."""
    ds, dxs, dys = [], [], []
    for s in range(0, txt.shape[0], chunk):
        e = slice(s, s + chunk)
        n = txt[e].shape[0]
        d, dx, dy = orc.ot_closed_form_grads(txt[e].double(), obj[e].double()[:, 1:], tnum[e] == 0, onum[e][:, 1:] == 0,
                                             torch.full((n,), 0.01, dtype=torch.float64))
        ds.append(d), dxs.append(dx), dys.append(dy)
    return torch.cat(ds), torch.cat(dxs), torch.cat(dys)


@pytest.mark.parametrize("wl,dtype", [("c3", torch.bfloat16), ("c4", torch.bfloat16), ("c4", torch.float32)])
def test_ot_full_batch_vs_oracle(wl, dtype):
    """The whole batch of the bench workloads (c3: 4096 x 16x50x512 -> streaming kernel; c4: 1024 x 32x257x768 ->
    TMA-fed cost / ragged solver / TMA-fed gradient in bf16, the three-launch mma.sync path in fp32): the work split
    of the persistent kernels depends on the batch, so the small-batch parity cases do not cover it."""
    w = syn.WORKLOADS[wl]
    txt, obj, tnum, onum = syn.ot_inputs(w.B, w.M, w.N, w.D, 23, "ragged", dtype=dtype)
    d_ref, dx_ref, dy_ref = _oracle_ot_chunked(txt, obj, tnum, onum)
    loss, dist, dtxt, dobj = _run_ot(txt, obj, tnum, onum)
    lt, gt = (1e-5, 2e-5) if dtype == torch.float32 else (2e-3, 1e-2)
    assert close(loss, 0.01 * d_ref.sum().item(), lt)
    assert rel_err(dist, d_ref) < lt
    assert rel_err(dtxt, dx_ref) < gt and rel_err(dobj[:, 1:], dy_ref) < gt
    # per sample, not only in the aggregate: no sample may be off by more than 5x the aggregate tolerance
    per = (dobj[:, 1:].double() - dy_ref).flatten(1).norm(dim=1) / dy_ref.flatten(1).norm(dim=1).clamp_min(1e-30)
    assert per.max().item() < 5 * gt
    assert (dobj[:, 0] == 0).all() and (dtxt[tnum == 0] == 0).all() and (dobj[onum == 0] == 0).all()


@pytest.mark.parametrize("wl", ["c3", "c4"])
def test_ot_full_batch_properties(wl):
    """Invariants of the cosine-cost OT distance on the bf16 path at full batch: node vectors rescaled by powers
    of two (exact in bf16) leave every distance unchanged and scale the gradients by the inverse factor; the
    result of a sample does not depend on which batch it is solved in; pad contents are never read into the result."""
    w = syn.WORKLOADS[wl]
    dt = torch.bfloat16
    txt, obj, tnum, onum = syn.ot_inputs(w.B, w.M, w.N, w.D, 29, "ragged", dtype=dt)
    loss, dist, dtxt, dobj = _run_ot(txt, obj, tnum, onum)
    assert (dist >= -1e-6).all() and torch.isfinite(dtxt).all() and torch.isfinite(dobj).all()
    # x2 on the text nodes, x0.5 on the image nodes
    loss2, dist2, dtxt2, dobj2 = _run_ot(txt * 2, obj * 0.5, tnum, onum)
    assert rel_err(dist2, dist) < 1e-4 and close(loss2, loss, 1e-4)
    assert rel_err(dtxt2 * 2, dtxt) < 2e-3 and rel_err(dobj2 * 0.5, dobj) < 2e-3
    # a slice of the batch solved alone
    s = slice(w.B // 2 - 37, w.B // 2 + 64)
    _, dist_s, dtxt_s, dobj_s = _run_ot(txt[s].contiguous(), obj[s].contiguous(), tnum[s].contiguous(), onum[s].contiguous())
    assert rel_err(dist_s, dist[s]) < 1e-4
    assert rel_err(dtxt_s, dtxt[s]) < 2e-3 and rel_err(dobj_s, dobj[s]) < 2e-3
    # finite garbage in the padded slots
    txt_g = torch.where((tnum == 0).unsqueeze(-1), torch.full_like(txt, 3.0), txt)
    obj_g = torch.where((onum == 0).unsqueeze(-1), torch.full_like(obj, -5.0), obj)
    loss_g, dist_g, dtxt_g, dobj_g = _run_ot(txt_g, obj_g, tnum, onum)
    assert rel_err(dist_g, dist) < 1e-6 and close(loss_g, loss, 1e-6)
    assert rel_err(dtxt_g, dtxt) < 1e-4 and rel_err(dobj_g[:, 1:], dobj[:, 1:]) < 1e-4
    assert (dtxt_g[tnum == 0] == 0).all() and (dobj_g[onum == 0] == 0).all()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_contrastive_known_answer_full_size(dtype):
    """c3 size (4096 images x 36864 descriptions), every feature the same vector: all logits are equal, so
    loss_i = ln(B*T), loss_t = ln(B) and the feature gradients vanish -- a known answer that needs no oracle run."""
    w = syn.WORKLOADS["c3"]
    B, T, D = w.B, w.T, w.D
    v = torch.randn(1, D, generator=torch.Generator().manual_seed(5))
    img = v.expand(B, D).contiguous().to(dtype).cuda().requires_grad_(True)
    txt = v.expand(B * T, D).contiguous().to(dtype).cuda().requires_grad_(True)
    ls = torch.tensor(syn.LOGIT_SCALE_INIT).cuda().requires_grad_(True)
    lpi, lpt, idx = syn.contrastive_labels(B, T)
    li, lt = F_.contrastive_over_batch(img, txt, ls, lpi.cuda(), lpt.cuda(), idx.cuda())
    (li + lt).backward()
    torch.cuda.synchronize()
    tol = 1e-5 if dtype == torch.float32 else 2e-3
    assert close(li.item(), math.log(B * T), tol) and close(lt.item(), math.log(B), tol)
    # a healthy gradient row at this size has norm ~ s / (B |v|) ~ 1.5e-4; these must vanish against that
    scale = math.exp(syn.LOGIT_SCALE_INIT) / (B * v.norm().item())
    lim = (1e-3 if dtype == torch.float32 else 2e-2) * scale
    assert img.grad.float().norm(dim=1).max().item() < lim and txt.grad.float().norm(dim=1).max().item() < lim
    # dlogit_scale = sum(G * L) = s * sum(G): 0 here; bf16 mode keeps the exponentials as bf16 (2^-9 relative each)
    assert abs(ls.grad.item()) < (1e-3 if dtype == torch.float32 else 1e-1)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_contrastive_scale_and_order_invariance_full_size(dtype):
    """c3 size: power-of-two rescaling of the features (exact in both dtypes) leaves the losses unchanged and scales
    the gradients by the inverse; moving the images together with their description blocks permutes the gradients."""
    w = syn.WORKLOADS["c3"]
    B, T, D = w.B, w.T, w.D
    img, txt, ls = syn.contrastive_inputs(B, T, D, 3, "trained", dtype=dtype)
    lpi, lpt, idx = (t.cuda() for t in syn.contrastive_labels(B, T))

    def run(i, t):
        ig, tg, sg = i.cuda().requires_grad_(True), t.cuda().requires_grad_(True), ls.cuda().requires_grad_(True)
        li, lt = F_.contrastive_over_batch(ig, tg, sg, lpi, lpt, idx)
        (li + lt).backward()
        torch.cuda.synchronize()
        return li.item(), lt.item(), ig.grad.float().cpu(), tg.grad.float().cpu(), sg.grad.item()

    li, lt, di, dt_, dls = run(img, txt)
    lt_tol, g_tol = (1e-5, 1e-4) if dtype == torch.float32 else (2e-3, 1e-2)
    li2, lt2, di2, dt2, dls2 = run(img * 4, txt * 0.5)
    assert close(li2, li, lt_tol, 2e-6) and close(lt2, lt, lt_tol, 2e-6)
    assert rel_err(di2 * 4, di) < g_tol and rel_err(dt2 * 0.5, dt_) < g_tol and close(dls2, dls, g_tol, 1e-5)
    p = torch.randperm(B, generator=torch.Generator().manual_seed(1))
    li3, lt3, di3, dt3, _ = run(img[p].contiguous(), txt.view(B, T, D)[p].reshape(B * T, D).contiguous())
    assert close(li3, li, lt_tol, 2e-6) and close(lt3, lt, lt_tol, 2e-6)
    assert rel_err(di3, di[p]) < g_tol and rel_err(dt3.view(B, T, D), dt_.view(B, T, D)[p]) < g_tol
