"""world_size-2 gloo runs of the sharded loss head's host choreography (CPU, no GPU).

The per-rank kernels are replaced by tests/oracle_backend.py; what is under test is
clip_event_b200.distributed: gathers, offsets, statistics exchange, reduce-scatter and the
gradient semantics.  Oracle = single-process reference maths on the concatenated global batch.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import rel_err
from clip_event_b200 import distributed as cd
from clip_event_b200 import synthetic as syn
from oracle import clip_event_oracle as orc
from oracle_backend import OracleBackend


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, B, T, D, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        img, txt, ls = syn.contrastive_inputs(B, T, D, 5, "trained")
        lo, hi = cd.shard_bounds(B, world, rank)
        b = hi - lo
        li, lt, ip = cd.global_labels_for_rank(b, T, world, rank)
        img_l = img[lo:hi].clone().requires_grad_(True)
        txt_l = txt[lo * T:hi * T].clone().requires_grad_(True)
        ls_l = ls.clone().requires_grad_(True)
        loss_i, loss_t = cd.global_contrastive(img_l, txt_l, ls_l, li, lt, ip, compute=OracleBackend())
        # OT shard: fake local kernel = oracle alignment criterion
        etxt, obj, tnum, onum = syn.ot_inputs(B, 4, 6, 16, 6, "ragged")
        e_l = etxt[lo:hi].clone().requires_grad_(True)
        o_l = obj[lo:hi].clone().requires_grad_(True)

        def ot_fn(a, b_, c, d):
            return orc.alignment_criterion(a, b_, c, d)["loss_ot"], None
        loss_ot = cd.sharded_alignment(e_l, o_l, tnum[lo:hi], onum[lo:hi], ot_fn=ot_fn)
        (2.0 * loss_i + 0.5 * loss_t + loss_ot).backward()
        # numpy copies, pickled by value: a torch tensor travels as a shared-memory handle that the parent must fetch
        # from THIS process, which may have exited by then
        q.put((rank, loss_i.item(), loss_t.item(), loss_ot.item(),
               *[t.grad.detach().numpy().copy() for t in (img_l, txt_l, ls_l, e_l, o_l)]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_global_contrastive_matches_single_process(world):
    B, T, D = 8, 3, 16
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, T, D, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted([q.get(timeout=300) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0

    img, txt, ls = syn.contrastive_inputs(B, T, D, 5, "trained")
    lpi, lpt, idx = syn.contrastive_labels(B, T)
    li, lt, dimg, dtxt, dls = orc.contrastive_closed_form(img.double(), txt.double(), ls.double(), lpi, lpt, idx,
                                                          g_i=2.0, g_t=0.5)
    etxt, obj, tnum, onum = syn.ot_inputs(B, 4, 6, 16, 6, "ragged")
    _, g_ot = orc.loss_head_step(img, txt, ls, lpi, lpt, idx, etxt, obj, tnum, onum)
    ot_total = orc.alignment_criterion(etxt, obj, tnum, onum)["loss_ot"].item()
    b = B // world
    for rank, loss_i, loss_t, loss_ot, gi, gt, gls, ge, go in results:
        assert abs(loss_i - li.item()) < 1e-5 and abs(loss_t - lt.item()) < 1e-5
        assert abs(loss_ot - ot_total) < 1e-6          # replicated global sum
        assert rel_err(gi, dimg[rank * b:(rank + 1) * b]) < 1e-5
        assert rel_err(gt, dtxt[rank * b * T:(rank + 1) * b * T]) < 1e-5
        assert abs(gls.item() - dls.item()) < 1e-4 * max(1.0, abs(dls.item()))
        assert rel_err(ge, g_ot["entitytxt_vec"][rank * b:(rank + 1) * b]) < 1e-5
        assert rel_err(go, g_ot["object_vec"][rank * b:(rank + 1) * b]) < 1e-5


def test_shard_helpers():
    assert cd.shard_bounds(4096, 8, 3) == (1536, 2048)
    with pytest.raises(RuntimeError):
        cd.shard_bounds(10, 4, 0)
    li, lt, ip = cd.global_labels_for_rank(2, 3, 4, 1)
    assert li.tolist() == [6, 9] and lt.tolist() == [2, 2, 2, 3, 3, 3] and ip.tolist() == [0, 3]


# --------------------------------------------------------------------------------------------
# DDP convention: encoder parameters under DistributedDataParallel (gradient AVERAGING)
# must receive the gradient of the GLOBAL-mean loss; logit_scale the full gradient.
# --------------------------------------------------------------------------------------------
class _TinyEncoders(torch.nn.Module):
    def __init__(self, d_in, d):
        super().__init__()
        g = torch.Generator().manual_seed(77)
        self.img = torch.nn.Linear(d_in, d, bias=False)
        self.txt = torch.nn.Linear(d_in, d, bias=False)
        self.logit_scale = torch.nn.Parameter(torch.tensor(syn.LOGIT_SCALE_INIT))
        with torch.no_grad():
            self.img.weight.copy_(torch.randn(d, d_in, generator=g) * 0.3)
            self.txt.weight.copy_(torch.randn(d, d_in, generator=g) * 0.3)

    def forward(self, xi, xt):
        return self.img(xi), self.txt(xt), self.logit_scale


def _ddp_worker(rank, world, port, B, T, D, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        import clip_event_b200 as ce
        from clip_event_b200.model_clip import LazyLogits
        xi, xt, _ = syn.contrastive_inputs(B, T, 12, 9, "trained")
        lo, hi = cd.shard_bounds(B, world, rank)
        b = hi - lo
        model = torch.nn.parallel.DistributedDataParallel(_TinyEncoders(12, D))
        fi, ft, ls = model(xi[lo:hi], xt[lo * T:hi * T])
        crit = ce.CriterionContrastive("ce", group=True, ddp_average=True, compute=OracleBackend())
        lpi, lpt, idx = syn.contrastive_labels(b, T)            # per-rank labels, as the collate_fn builds them
        ld = crit(LazyLogits(fi, ft, ls, "per_image"), LazyLogits(ft, fi, ls, "per_text"), lpi, lpt, index_pos=idx)
        sum(ld.values()).backward()
        m = model.module
        # numpy copies: a tensor in the queue shares its storage through a file descriptor the exiting worker takes with it
        q.put((rank, ld["loss_i"].item(), ld["loss_t"].item(), m.img.weight.grad.numpy().copy(), m.txt.weight.grad.numpy().copy(),
               float(m.logit_scale.grad)))
    finally:
        dist.destroy_process_group()


def test_ddp_parameter_gradients_match_single_process():
    world, B, T, D = 2, 8, 3, 16
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ddp_worker, args=(r, world, port, B, T, D, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted([q.get(timeout=300) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    # single process: the same encoders on the whole batch, reference loss (global mean), autograd
    xi, xt, _ = syn.contrastive_inputs(B, T, 12, 9, "trained")
    enc = _TinyEncoders(12, D).double()
    fi, ft, ls = enc(xi.double(), xt.double())
    lpi, lpt, idx = syn.contrastive_labels(B, T)
    a, b_ = orc.similarity_logits(fi, ft, ls, True)
    ld = orc.contrastive_criterion(a, b_, lpi, lpt, idx, True, "ce")
    sum(ld.values()).backward()
    for rank, li, lt, gwi, gwt, gls in results:
        assert abs(li - ld["loss_i"].item()) < 1e-5 and abs(lt - ld["loss_t"].item()) < 1e-5
        assert rel_err(gwi, enc.img.weight.grad) < 2e-5      # DDP average of world-scaled shards = true gradient
        assert rel_err(gwt, enc.txt.weight.grad) < 2e-5
        assert abs(gls - enc.logit_scale.grad.item()) < 1e-4 * max(1.0, abs(enc.logit_scale.grad.item()))


# --------------------------------------------------------------------------------------------
# the one-call sharded step: three exchange steps, losses and gradients formed together
# --------------------------------------------------------------------------------------------
def _cpu_ot_eager(etxt, obj, tnum, onum, need_grad, stream_ptr):
    tp, ip = tnum == 0, onum[:, 1:] == 0
    d, dx, dy = orc.ot_closed_form_grads(etxt.double(), obj.double()[:, 1:], tp, ip,
                                         torch.full((etxt.shape[0],), 0.01, dtype=torch.float64))
    dobj = torch.zeros_like(obj)
    dobj[:, 1:] = dy.to(obj.dtype)
    return (0.01 * d.sum()).float().reshape(1), dx.to(etxt.dtype), dobj, None


def _step_worker(rank, world, port, B, T, D, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        img, txt, ls = syn.contrastive_inputs(B, T, D, 5, "trained")
        etxt, obj, tnum, onum = syn.ot_inputs(B, 4, 6, 16, 6, "ragged")
        lo, hi = cd.shard_bounds(B, world, rank)
        b = hi - lo
        li, lt, ip = cd.global_labels_for_rank(b, T, world, rank)
        lv = [t.clone().requires_grad_(True) for t in (img[lo:hi], txt[lo * T:hi * T], ls, etxt[lo:hi], obj[lo:hi])]
        a, b_, c = cd.global_loss_head_step(lv[0], lv[1], lv[2], li, lt, ip, lv[3], lv[4], tnum[lo:hi], onum[lo:hi],
                                            compute=OracleBackend(), ot_eager=_cpu_ot_eager)
        (a + b_ + c).backward()
        q.put((rank, a.item(), b_.item(), c.item(), [t.grad.numpy().copy() for t in lv]))
    finally:
        dist.destroy_process_group()


def test_global_loss_head_step_matches_single_process():
    world, B, T, D = 2, 8, 3, 16
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_step_worker, args=(r, world, port, B, T, D, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted([q.get(timeout=300) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    img, txt, ls = syn.contrastive_inputs(B, T, D, 5, "trained")
    lpi, lpt, idx = syn.contrastive_labels(B, T)
    li, lt, dimg, dtxt, dls = orc.contrastive_closed_form(img.double(), txt.double(), ls.double(), lpi, lpt, idx)
    etxt, obj, tnum, onum = syn.ot_inputs(B, 4, 6, 16, 6, "ragged")
    _, g_ot = orc.loss_head_step(img, txt, ls, lpi, lpt, idx, etxt, obj, tnum, onum)
    ot_total = orc.alignment_criterion(etxt, obj, tnum, onum)["loss_ot"].item()
    b = B // world
    for rank, loss_i, loss_t, loss_ot, (gi, gt, gls, ge, go) in results:
        assert abs(loss_i - li.item()) < 1e-5 and abs(loss_t - lt.item()) < 1e-5
        assert abs(loss_ot - ot_total) < 1e-6 * max(1.0, abs(ot_total))      # replicated global sum, packed in the stats record
        assert rel_err(gi, dimg[rank * b:(rank + 1) * b]) < 1e-5
        assert rel_err(gt, dtxt[rank * b * T:(rank + 1) * b * T]) < 1e-5
        assert abs(gls.item() - dls.item()) < 1e-4 * max(1.0, abs(dls.item()))
        assert rel_err(ge, g_ot["entitytxt_vec"][rank * b:(rank + 1) * b]) < 1e-5
        assert rel_err(go, g_ot["object_vec"][rank * b:(rank + 1) * b]) < 1e-5
