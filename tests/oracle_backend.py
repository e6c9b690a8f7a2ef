"""CPU stand-in for clip_event_b200.distributed.CudaBackend, built on plain torch (test-only).

Implements the four C-ABI phases (include/clip_event_b200.h) with the same inputs/outputs so the
collective choreography in clip_event_b200.distributed can be exercised with gloo on CPU.
"""
import math

import torch

LOG2E = 1.0 / math.log(2.0)


class OracleBackend:
    def _logits2(self, img_all, txt, ls):
        ih = img_all.double() / img_all.double().norm(dim=-1, keepdim=True)
        th = txt.double() / txt.double().norm(dim=-1, keepdim=True)
        return ls.double().exp() * (ih @ th.t()) * LOG2E, ih, th      # base-2 scaled logits

    def fwd_partial(self, img_all, txt, ls, labels_i_all, labels_t, index_pos, col_offset):
        L2, ih, th = self._logits2(img_all, txt, ls)
        R, C = L2.shape
        m = L2.max(dim=1).values
        l = torch.exp2(L2 - m.view(-1, 1)).sum(1)
        lab = labels_i_all - col_offset
        local = (lab >= 0) & (lab < C)
        lab_logit = torch.zeros(R, dtype=torch.float64)
        rows = torch.arange(R)[local]
        lab_logit[rows] = L2[rows, lab[local]] / LOG2E
        row_part = torch.stack([m, l, lab_logit, torch.zeros(R, dtype=torch.float64)], 1).float()
        Lp = L2[:, index_pos]                                    # [R, P]
        col_lse2 = torch.logsumexp(Lp / LOG2E, 0) * LOG2E
        lab_rows = labels_t[index_pos]
        item_t = col_lse2 / LOG2E - Lp[lab_rows, torch.arange(index_pos.numel())] / LOG2E
        sums = torch.tensor([item_t.sum().item(), float(index_pos.numel()), 0.0, 0.0])
        state = dict(L2=L2, ih=ih, th=th, col_lse2=col_lse2, lab=lab, local=local, R=R)
        return torch.cat([row_part.reshape(-1), sums]), state

    def fwd_finish(self, stats_all, world, state):
        R = state["R"]
        row_part_all, sums_all = stats_all[:, : R * 4], stats_all[:, R * 4:]
        rp = row_part_all.reshape(world, -1, 4).double()
        m = rp[:, :, 0].max(0).values
        l = (rp[:, :, 1] * torch.exp2(rp[:, :, 0] - m)).sum(0)
        lse2 = m + torch.log2(l)
        state["lse2_row"] = lse2
        loss_i = (lse2 / LOG2E - rp[:, :, 2].sum(0)).mean()
        sa = sums_all.reshape(world, 4).double()
        loss_t = sa[:, 0].sum() / sa[:, 1].sum()
        return loss_i.float(), loss_t.float()

    def bwd_partial(self, img_all, txt, ls, labels_i_all, labels_t, index_pos, col_offset, g_i, g_t,
                    R_total, P_total, state):
        L2, ih, th = state["L2"], state["ih"], state["th"]
        R, C = L2.shape
        G = torch.exp2(L2 - state["lse2_row"].view(-1, 1))
        rows = torch.arange(R)[state["local"]]
        G[rows, state["lab"][state["local"]]] -= 1
        G = G * (g_i.double() / R_total)
        Gc = torch.exp2(L2[:, index_pos] - state["col_lse2"].view(1, -1))
        Gc[labels_t[index_pos], torch.arange(index_pos.numel())] -= 1
        G.index_add_(1, index_pos, Gc * (g_t.double() / P_total))
        s = ls.double().exp()
        dimg_hat = (s * G @ th).float()
        dth = s * G.t() @ ih
        nt = txt.double().norm(dim=-1, keepdim=True)
        dtxt = ((dth - th * (th * dth).sum(-1, keepdim=True)) / nt).to(txt.dtype)
        dls = (G * L2 / LOG2E).sum().float().reshape(1)
        return dtxt, dimg_hat, dls

    def bwd_finish(self, img_rows, dimg_hat_rows):
        x = img_rows.double()
        n = x.norm(dim=-1, keepdim=True)
        xh = x / n
        d = dimg_hat_rows.double()
        return ((d - xh * (xh * d).sum(-1, keepdim=True)) / n).to(img_rows.dtype)
