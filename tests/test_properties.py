"""Property tests of the oracle (CPU, hypothesis): invariants of the reference's algorithm that do not depend on
the problem size.  The GPU suite checks the same invariants on the CUDA path at BASELINE sizes
(tests/test_gpu_properties.py); here they pin the checker itself on randomly drawn shapes, masks and seeds.

Reference lines: model_ot.py:8-84 (cosine cost, IPOT, trace), model_clip.py:495-528,620-662 (similarity + InfoNCE).
"""
import math

import torch
from hypothesis import given, settings, strategies as st

from clip_event_b200 import synthetic as syn
from oracle import clip_event_oracle as orc

SET = dict(max_examples=20, deadline=None, derandomize=True)


def _ot_case(B, M, N, D, seed, masks):
    txt, obj, tnum, onum = syn.ot_inputs(B, M, N, D, seed, masks)
    return txt.double(), obj.double()[:, 1:], tnum == 0, onum[:, 1:] == 0


@settings(**SET)
@given(B=st.integers(1, 4), M=st.integers(1, 9), N=st.integers(1, 13), D=st.sampled_from([4, 16, 40]),
       seed=st.integers(0, 10_000), masks=st.sampled_from(["full", "ragged", "scattered"]),
       iters=st.sampled_from([1, 10, 50]), beta=st.sampled_from([0.3, 0.5, 1.0]))
def test_ipot_plan_marginals_and_pads(B, M, N, D, seed, masks, iters, beta):
    """model_ot.py:55-62: the last update of an iteration is sigma, so every valid text node carries exactly
    1/x_len of the plan's mass (total mass 1); pads carry none; the plan is non-negative."""
    x, y, xp, yp = _ot_case(B, M, N, D, seed, masks)
    C = orc.cost_matrix_cosine(x, y)
    jp = xp.unsqueeze(-1) | yp.unsqueeze(-2)
    C = C.masked_fill(jp, 0)
    xl = (M - xp.sum(1)).double()
    yl = (N - yp.sum(1)).double()
    T = orc.ipot(C, xl, xp, yl, yp, jp, beta, iters, 1)                      # [B, N, M]
    assert (T >= 0).all() and (T.masked_select(jp.transpose(1, 2)) == 0).all()
    col = T.sum(1)                                                           # mass per text node
    want = (1.0 / xl).view(B, 1).expand(B, M).masked_fill(xp, 0)
    assert torch.allclose(col, want, rtol=1e-9, atol=1e-12)
    assert torch.allclose(T.sum((1, 2)), torch.ones(B, dtype=torch.float64), rtol=1e-9)


@settings(**SET)
@given(B=st.integers(1, 4), M=st.integers(1, 9), N=st.integers(1, 13), D=st.sampled_from([8, 24]),
       seed=st.integers(0, 10_000), masks=st.sampled_from(["full", "ragged", "scattered", "edge"]))
def test_ot_distance_invariances(B, M, N, D, seed, masks):
    """Cosine cost: the distance ignores the length of every node vector, the order of the nodes inside a set
    and whatever sits in padded slots; it is non-negative and 0 for an empty set (model_ot.py:62,73-74)."""
    x, y, xp, yp = _ot_case(B, M, N, D, seed, masks)
    d = orc.optimal_transport_dist(x, y, xp, yp)
    assert (d >= -1e-12).all()
    empty = xp.all(1) | yp.all(1)
    assert (d[empty] == 0).all()
    g = torch.Generator().manual_seed(seed)
    # per-node positive rescaling
    sx = torch.rand(B, M, 1, generator=g, dtype=torch.float64) * 3 + 0.25
    sy = torch.rand(B, N, 1, generator=g, dtype=torch.float64) * 3 + 0.25
    assert torch.allclose(orc.optimal_transport_dist(x * sx, y * sy, xp, yp), d, rtol=1e-9, atol=1e-12)
    # node order
    pm, pn = torch.randperm(M, generator=g), torch.randperm(N, generator=g)
    dp = orc.optimal_transport_dist(x[:, pm], y[:, pn], xp[:, pm], yp[:, pn])
    assert torch.allclose(dp, d, rtol=1e-9, atol=1e-12)
    # pad contents
    x2 = torch.where(xp.unsqueeze(-1), torch.full_like(x, 7.5), x)
    y2 = torch.where(yp.unsqueeze(-1), torch.full_like(y, -3.25), y)
    assert torch.allclose(orc.optimal_transport_dist(x2, y2, xp, yp), d, rtol=1e-12, atol=0)


@settings(**SET)
@given(B=st.integers(1, 3), M=st.integers(1, 7), N=st.integers(1, 9), D=st.sampled_from([8, 24]),
       seed=st.integers(0, 10_000), masks=st.sampled_from(["full", "ragged", "scattered"]))
def test_ot_closed_form_gradients_equal_autograd(B, M, N, D, seed, masks):
    """with the plan detached (model_ot.py:81,83) the gradient is the cost's gradient contracted with
    the plan; padded nodes get none, and every node's gradient is orthogonal to the node (scale invariance)."""
    x, y, xp, yp = _ot_case(B, M, N, D, seed, masks)
    xg, yg = x.clone().requires_grad_(True), y.clone().requires_grad_(True)
    up = torch.linspace(0.5, 1.5, B, dtype=torch.float64)
    (orc.optimal_transport_dist(xg, yg, xp, yp) * up).sum().backward()
    d, dx, dy = orc.ot_closed_form_grads(x, y, xp, yp, up)
    assert torch.allclose(dx, xg.grad, rtol=1e-8, atol=1e-12) and torch.allclose(dy, yg.grad, rtol=1e-8, atol=1e-12)
    assert (dx[xp] == 0).all() and (dy[yp] == 0).all()
    assert ((dx * x).sum(-1).abs() <= 1e-10 * (dx.norm(dim=-1) * x.norm(dim=-1) + 1e-30) + 1e-14).all()


@settings(**SET)
@given(B=st.integers(2, 9), T=st.integers(1, 5), D=st.sampled_from([8, 32]), seed=st.integers(0, 10_000),
       kind=st.sampled_from(["iid", "trained"]), ls=st.sampled_from([0.0, math.log(1 / 0.07), math.log(100.0)]))
def test_contrastive_invariances(B, T, D, seed, kind, ls):
    """model_clip.py:495-528,633-662: features are L2-normalised first, so the losses ignore their lengths and the
    gradients are orthogonal to them; permuting the images together with their description blocks changes nothing;
    the closed form used by the kernels equals autograd through the criterion."""
    img, txt, _ = syn.contrastive_inputs(B, max(T, 1), D, seed, kind if T > 1 else "iid")
    img, txt = img.double(), txt.double()
    s = torch.tensor(ls, dtype=torch.float64)
    lpi, lpt, idx = syn.contrastive_labels(B, T)
    li, lt, dimg, dtxt, dls = orc.contrastive_closed_form(img, txt, s, lpi, lpt, idx)
    ig, tg, sg = img.clone().requires_grad_(True), txt.clone().requires_grad_(True), s.clone().requires_grad_(True)
    a, b = orc.similarity_logits(ig, tg, sg)
    out = orc.contrastive_criterion(a, b, lpi, lpt, index_pos=idx)
    (out["loss_i"] + out["loss_t"]).backward()
    assert torch.allclose(li, out["loss_i"], rtol=1e-10) and torch.allclose(lt, out["loss_t"], rtol=1e-10, atol=1e-14)
    assert torch.allclose(dimg, ig.grad, rtol=1e-7, atol=1e-12) and torch.allclose(dtxt, tg.grad, rtol=1e-7, atol=1e-12)
    assert torch.allclose(dls, sg.grad, rtol=1e-7, atol=1e-12)
    assert ((dimg * img).sum(1).abs() <= 1e-9 * dimg.norm(dim=1) * img.norm(dim=1) + 1e-14).all()
    assert ((dtxt * txt).sum(1).abs() <= 1e-9 * dtxt.norm(dim=1) * txt.norm(dim=1) + 1e-14).all()
    # lengths
    g = torch.Generator().manual_seed(seed)
    si = torch.rand(B, 1, generator=g, dtype=torch.float64) * 4 + 0.1
    stx = torch.rand(B * T, 1, generator=g, dtype=torch.float64) * 4 + 0.1
    li2, lt2 = orc.contrastive_closed_form(img * si, txt * stx, s, lpi, lpt, idx)[:2]
    assert torch.allclose(li2, li, rtol=1e-10) and torch.allclose(lt2, lt, rtol=1e-10, atol=1e-14)
    # sample order: image b and its T descriptions move together, the canonical labels stay
    p = torch.randperm(B, generator=g)
    li3, lt3, dimg3 = orc.contrastive_closed_form(img[p], txt.view(B, T, D)[p].reshape(B * T, D), s, lpi, lpt, idx)[:3]
    assert torch.allclose(li3, li, rtol=1e-10) and torch.allclose(lt3, lt, rtol=1e-10, atol=1e-14)
    assert torch.allclose(dimg3, dimg[p], rtol=1e-8, atol=1e-13)


@settings(**SET)
@given(B=st.integers(2, 12), T=st.integers(1, 4), D=st.sampled_from([8, 32]), seed=st.integers(0, 1000))
def test_contrastive_known_answer_when_all_logits_are_equal(B, T, D, seed):
    """Every image and every description the same vector: all logits equal, so the image-side loss is ln(B*T), the
    text-side loss ln(B), and no feature receives a gradient -- a known answer at any size."""
    g = torch.Generator().manual_seed(seed)
    v = torch.randn(1, D, generator=g, dtype=torch.float64)
    img, txt = v.expand(B, D).contiguous(), v.expand(B * T, D).contiguous()
    lpi, lpt, idx = syn.contrastive_labels(B, T)
    li, lt, dimg, dtxt, dls = orc.contrastive_closed_form(img, txt, torch.tensor(math.log(1 / 0.07), dtype=torch.float64),
                                                          lpi, lpt, idx)
    assert abs(li.item() - math.log(B * T)) < 1e-9 and abs(lt.item() - math.log(B)) < 1e-9
    assert dimg.abs().max() < 1e-9 and dtxt.abs().max() < 1e-9 and abs(float(dls)) < 1e-9
