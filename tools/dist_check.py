"""Multi-GPU parity check (launch with torchrun, one rank per GPU): the sharded loss head against
the single-process oracle on the rank-order concatenation of every rank's inputs."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from clip_event_b200 import distributed as cd  # noqa: E402
from clip_event_b200 import synthetic as syn  # noqa: E402
from oracle import clip_event_oracle as orc  # noqa: E402


def rel(a, b):
    a, b = a.detach().double().cpu().flatten(), b.detach().double().cpu().flatten()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    for dtype, (lt, gt) in ((torch.float32, (1e-5, 3e-5)), (torch.bfloat16, (2e-3, 1e-2))):
        for (B, T, D, M, N) in ((64 * world, 5, 512, 8, 50), (96 * world, 9, 768, 16, 50)):
            img, txt, ls = syn.contrastive_inputs(B, T, D, 21, "trained", dtype=dtype)
            etxt, obj, tnum, onum = syn.ot_inputs(B, M, N, D, 22, "ragged", dtype=dtype)
            lo, hi = cd.shard_bounds(B, world, rank)
            b = hi - lo
            li_, lt_, ip_ = cd.global_labels_for_rank(b, T, world, rank, device=dev)
            leaves = dict(img=img[lo:hi], txt=txt[lo * T:hi * T], etxt=etxt[lo:hi], obj=obj[lo:hi])
            leaves = {k: v.to(dev).requires_grad_(True) for k, v in leaves.items()}
            lsg = ls.to(dev).requires_grad_(True)
            loss_i, loss_t = cd.global_contrastive(leaves["img"], leaves["txt"], lsg,
                                                   li_ if D == 512 else None, lt_, ip_)   # explicit and canonical labels
            loss_ot = cd.sharded_alignment(leaves["etxt"], leaves["obj"], tnum[lo:hi].to(dev), onum[lo:hi].to(dev))
            (loss_i + loss_t + loss_ot).backward()
            torch.cuda.synchronize()
            if rank == 0 or True:
                lpi, lpt, idx = syn.contrastive_labels(B, T)
                ri, rt, rdi, rdt, rdls = orc.contrastive_closed_form(img.double(), txt.double(), ls.double(), lpi, lpt, idx)
                tp, ip = tnum == 0, onum[:, 1:] == 0
                d_ref, dx_ref, dy_ref = orc.ot_closed_form_grads(etxt.double(), obj.double()[:, 1:], tp, ip,
                                                                 torch.full((B,), 0.01, dtype=torch.float64))
                errs = dict(
                    loss_i=abs(loss_i.item() - ri.item()) / abs(ri.item()), loss_t=abs(loss_t.item() - rt.item()),
                    loss_ot=abs(loss_ot.item() - 0.01 * d_ref.sum().item()) / abs(0.01 * d_ref.sum().item()),
                    dimg=rel(leaves["img"].grad, rdi[lo:hi]), dtxt=rel(leaves["txt"].grad, rdt[lo * T:hi * T]),
                    dls=abs(lsg.grad.item() - rdls.item()) / max(1.0, abs(rdls.item())),
                    detxt=rel(leaves["etxt"].grad, dx_ref[lo:hi]), dobj=rel(leaves["obj"].grad[:, 1:], dy_ref[lo:hi]))
                good = (errs["loss_i"] < lt + 2e-6 and errs["loss_t"] < lt * abs(rt.item()) + (2e-6 if dtype == torch.float32 else 1e-4)
                        and errs["loss_ot"] < lt and errs["dimg"] < gt and errs["dtxt"] < gt and errs["detxt"] < gt
                        and errs["dobj"] < gt and errs["dls"] < (1e-4 if dtype == torch.float32 else 1e-2))
                ok = ok and good
                print("rank %d %s B=%d T=%d D=%d: %s %s" % (rank, str(dtype)[6:], B, T, D, "OK " if good else "BAD",
                                                         " ".join("%s=%.2e" % kv for kv in errs.items())), flush=True)
                # the same step through the one-call sharded loss head (what bench.py --gpus N times), with an
                # upstream gradient of 2 on every loss: one scale launch over all five gradient buffers
                lv = dict(img=img[lo:hi], txt=txt[lo * T:hi * T], etxt=etxt[lo:hi], obj=obj[lo:hi])
                lv = {k: v.to(dev).requires_grad_(True) for k, v in lv.items()}
                ls2 = ls.to(dev).requires_grad_(True)
                f_i, f_t, f_o = cd.global_loss_head_step(lv["img"], lv["txt"], ls2, li_ if D == 512 else None, lt_, ip_,
                                                         lv["etxt"], lv["obj"], tnum[lo:hi].to(dev), onum[lo:hi].to(dev))
                (2.0 * (f_i + f_t + f_o)).backward()
                torch.cuda.synchronize()
                e2 = dict(loss_i=abs(f_i.item() - ri.item()) / abs(ri.item()), loss_t=abs(f_t.item() - rt.item()),
                          loss_ot=abs(f_o.item() - 0.01 * d_ref.sum().item()) / abs(0.01 * d_ref.sum().item()),
                          dimg=rel(lv["img"].grad, 2 * rdi[lo:hi]), dtxt=rel(lv["txt"].grad, 2 * rdt[lo * T:hi * T]),
                          dls=abs(ls2.grad.item() - 2 * rdls.item()) / max(1.0, abs(2 * rdls.item())),
                          detxt=rel(lv["etxt"].grad, 2 * dx_ref[lo:hi]), dobj=rel(lv["obj"].grad[:, 1:], 2 * dy_ref[lo:hi]))
                good2 = (e2["loss_i"] < lt + 2e-6 and e2["loss_t"] < lt * abs(rt.item()) + (2e-6 if dtype == torch.float32 else 1e-4)
                         and e2["loss_ot"] < lt and e2["dimg"] < gt and e2["dtxt"] < gt and e2["detxt"] < gt
                         and e2["dobj"] < gt and e2["dls"] < (1e-4 if dtype == torch.float32 else 1e-2))
                ok = ok and good2
                print("rank %d %s B=%d one-call step: %s %s" % (rank, str(dtype)[6:], B, "OK " if good2 else "BAD",
                                                              " ".join("%s=%.2e" % kv for kv in e2.items())), flush=True)
    flag = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(flag)
    dist.destroy_process_group()
    if flag.item():
        sys.exit(1)
    if rank == 0:
        print("dist_check: all ranks OK")


if __name__ == "__main__":
    main()
