"""Parity + timing of the fused OT kernel against the fp64 oracle (run once with CE_OT_FUSED=1 and once
with CE_OT_FUSED=0 for the three-kernel path):  python tools/ot_fused_check.py [time]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from clip_event_b200 import functional as F_, synthetic as syn
from oracle import clip_event_oracle as orc


def rel(a, b):
    a, b = a.detach().double().cpu().flatten(), b.detach().double().cpu().flatten()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def check(B, M, N, D, masks, kind, iters=50, beta=0.5, seed=13):
    txt, obj, tnum, onum = syn.ot_inputs(B, M, N, D, seed, masks, kind, dtype=torch.bfloat16)
    tp, ip = tnum == 0, onum[:, 1:] == 0
    d_ref, dx_ref, dy_ref = orc.ot_closed_form_grads(txt.double(), obj.double()[:, 1:], tp, ip,
                                                     torch.full((B,), 0.01, dtype=torch.float64), beta=beta, iteration=iters)
    t, o = txt.cuda().requires_grad_(True), obj.cuda().requires_grad_(True)
    loss, dist = F_.ot_alignment(t, o, tnum.cuda(), onum.cuda(), beta=beta, iters=iters)
    loss.backward()
    torch.cuda.synchronize()
    e = dict(loss=abs(loss.item() - 0.01 * d_ref.sum().item()) / max(abs(0.01 * d_ref.sum().item()), 1e-30),
             dist=rel(dist, d_ref), dtxt=rel(t.grad, dx_ref), dobj=rel(o.grad[:, 1:], dy_ref),
             slot0=float(o.grad[:, 0].abs().max()), finite=bool(torch.isfinite(t.grad).all() and torch.isfinite(o.grad).all()))
    ok = e["loss"] < 2e-3 and e["dtxt"] < 1e-2 and e["dobj"] < 1e-2 and e["slot0"] == 0 and e["finite"]
    print("%s B=%d M=%d N=%d D=%d %s/%s it=%d beta=%.1f: %s" % ("OK " if ok else "BAD", B, M, N, D, masks, kind, iters, beta,
          " ".join("%s=%.2e" % (k, v) for k, v in e.items())), flush=True)
    return ok


def timeit(wl, B=None):
    w = syn.WORKLOADS[wl]
    B = B or w.B
    etxt, obj, tnum, onum = syn.ot_inputs(B, w.M, w.N, w.D, 0, "ragged", dtype=torch.bfloat16)
    etxt, obj, tnum, onum = etxt.cuda(), obj.cuda(), tnum.cuda(), onum.cuda()
    eg, og = etxt.clone().requires_grad_(True), obj.clone().requires_grad_(True)

    def full():
        eg.grad = None; og.grad = None
        l, _ = F_.ot_alignment(eg, og, tnum, onum)
        l.backward()
    for _ in range(3):
        full()
    torch.cuda.synchronize()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        full()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        full()
    torch.cuda.synchronize()
    n = 30
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(5):
        g.replay()
    e0.record()
    for _ in range(n):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    byts = 2.0 * (w.M + w.N) * w.D * 2 * B
    print("time %s B=%d fused=%s: %.1f us  -> %.0f GB/s algorithmic (%.3f of 6546)" % (
        wl, B, os.environ.get("CE_OT_FUSED", "1"), us, byts / us / 1e3, byts / us / 1e3 / 6546), flush=True)


if __name__ == "__main__":
    ok = True
    cases = [(5, 4, 7, 128, "full", "iid"), (7, 6, 9, 128, "edge", "iid"), (6, 5, 11, 256, "scattered", "iid"),
             (32, 8, 50, 512, "ragged", "iid"), (16, 16, 50, 512, "full", "correlated"), (300, 16, 50, 512, "ragged", "iid"),
             (9, 16, 64, 768, "ragged", "iid"), (500, 13, 37, 384, "edge", "iid"), (3, 1, 1, 128, "full", "iid")]
    for c in cases:
        ok = check(*c) and ok
    for iters in (10, 25, 100):
        for beta in (0.3, 0.5):
            ok = check(64, 16, 50, 512, "ragged", "iid", iters, beta) and ok
    # bench size (28 samples per CTA: the steady state of the job ring) against the fp64 oracle on a slice of CTA 0 .. 147's samples
    w = syn.WORKLOADS["c3"]
    txt, obj, tnum, onum = syn.ot_inputs(w.B, w.M, w.N, w.D, 0, "ragged", dtype=torch.bfloat16)
    t, o = txt.cuda().requires_grad_(True), obj.cuda().requires_grad_(True)
    loss, dist = F_.ot_alignment(t, o, tnum.cuda(), onum.cuda())
    loss.backward()
    torch.cuda.synchronize()
    sel = torch.arange(0, w.B, 13)
    tp, ip = tnum[sel] == 0, onum[sel][:, 1:] == 0
    d_ref, dx_ref, dy_ref = orc.ot_closed_form_grads(txt[sel].double(), obj[sel].double()[:, 1:], tp, ip,
                                                     torch.full((sel.numel(),), 0.01, dtype=torch.float64))
    e = dict(dist=rel(dist.cpu()[sel], d_ref), dtxt=rel(t.grad.cpu()[sel], dx_ref), dobj=rel(o.grad.cpu()[sel][:, 1:], dy_ref),
             finite=bool(torch.isfinite(t.grad).all() and torch.isfinite(o.grad).all()), loss=float(loss))
    good = e["dist"] < 1e-4 and e["dtxt"] < 1e-2 and e["dobj"] < 1e-2 and e["finite"] and abs(e["loss"] - 39.0049) < 0.05
    print("%s c3 full size: %s" % ("OK " if good else "BAD", " ".join("%s=%.3e" % kv for kv in e.items())), flush=True)
    ok = ok and good
    print("ALL OK" if ok else "SOME BAD", flush=True)
    if len(sys.argv) > 1:
        timeit("c3")
        timeit("c2")
