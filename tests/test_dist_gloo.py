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
        q.put((rank, loss_i.item(), loss_t.item(), loss_ot.item(), img_l.grad, txt_l.grad, ls_l.grad,
               e_l.grad, o_l.grad))
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
    results = sorted([q.get(timeout=60) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
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
