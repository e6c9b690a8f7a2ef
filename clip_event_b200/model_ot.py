"""CUDA-backed mirror of the reference's ``model_ot.py`` (same names, arguments, results).

Reference: src/clip-event/model_ot.py:8-84 (itself based on UNITER's model/ot.py).  Every function
runs hand-written sm_100a kernels through the C ABI; CPU tensors raise.
"""
from __future__ import annotations

import torch

from . import _lib as L
from . import functional as F_


def cost_matrix_cosine(x, y, eps=1e-5):
    """model_ot.py:8-18.  [B,L_x,D],[B,L_y,D] -> cosine distance [B,L_x,L_y] (fp32).  No autograd:
    the differentiable path is :func:`optimal_transport_dist`."""
    assert x.dim() == y.dim()
    assert x.size(0) == y.size(0)
    assert x.size(2) == y.size(2)
    L.require_cuda(x, y)
    dt = L.dtype_code(x.dtype)
    xc, yc = x.detach().contiguous(), y.detach().contiguous()
    B, M, D = xc.shape
    N = yc.shape[1]
    out = torch.empty(B, M, N, dtype=torch.float32, device=x.device)
    L.check(L.load().ce_ot_cost_matrix(xc.data_ptr(), yc.data_ptr(), B, M, N, D, dt, float(eps),
                                       out.data_ptr(), L.stream_ptr()), "cost_matrix_cosine")
    return out


def trace(x):
    """model_ot.py:21-29.  Batched trace of [B,n,n]."""
    b, m, n = x.size()
    assert m == n
    L.require_cuda(x)
    xc = x.detach().to(torch.float32).contiguous()
    out = torch.empty(b, dtype=torch.float32, device=x.device)
    L.check(L.load().ce_ot_trace(xc.data_ptr(), b, n, out.data_ptr(), L.stream_ptr()), "trace")
    return out.to(x.dtype)


@torch.no_grad()
def ipot(C, x_len, x_pad, y_len, y_pad, joint_pad, beta, iteration, k):
    """model_ot.py:32-63.  [B,M,N] cost (already zeroed at pads) -> transport plan [B,N,M].

    ``x_len`` / ``y_len`` / ``joint_pad`` are implied by the pads and accepted for signature
    compatibility.  The reference raises for k > 1 (its sigma keeps shape [b,1,m] after the first
    inner step); the kernel implements the intended recurrence for any k >= 1.
    """
    L.require_cuda(C, x_pad, y_pad)
    b, m, n = C.size()
    Cc = C.detach().to(torch.float32).contiguous()
    xp = x_pad.to(torch.bool).contiguous().view(torch.uint8)
    yp = y_pad.to(torch.bool).contiguous().view(torch.uint8)
    plan = torch.empty(b, n, m, dtype=torch.float32, device=C.device)
    L.check(L.load().ce_ot_ipot(Cc.data_ptr(), xp.data_ptr(), yp.data_ptr(), b, m, n, float(beta),
                                int(iteration), int(k), plan.data_ptr(), L.stream_ptr()), "ipot")
    return plan.to(C.dtype)


def optimal_transport_dist(txt_emb, img_emb, txt_pad, img_pad, cost=None,
                           beta=0.5, iteration=50, k=1):
    """model_ot.py:66-84.  [B,M,D],[B,N,D],[B,M],[B,N] -> OT distance [B], differentiable w.r.t. the
    embeddings through the cost only (the plan is detached, model_ot.py:81-83)."""
    if cost is not None:
        # reference path with a caller-supplied cost: mask, solve, contract -- no embedding gradient
        joint_pad = txt_pad.unsqueeze(-1) | img_pad.unsqueeze(-2)
        cost = cost.masked_fill(joint_pad, 0)
        T = ipot(cost.detach(), None, txt_pad, None, img_pad, joint_pad, beta, iteration, k)
        return (cost * T.transpose(1, 2).to(cost.dtype)).sum((1, 2))
    _, dist = F_.ot_alignment(txt_emb, img_emb, txt_pad.to(torch.bool), img_pad.to(torch.bool),
                              drop_slot0=False, beta=beta, iters=iteration, k=k, loss_scale=1.0)
    return dist.to(txt_emb.dtype)
