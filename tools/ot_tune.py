"""Times ce_ot_fwd_bwd (forward only = cost + IPOT, and full) for a workload:
python tools/ot_tune.py c4 bf16 [batch]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from clip_event_b200 import functional as F_, synthetic as syn
wl, dt = sys.argv[1], (torch.bfloat16 if sys.argv[2] == "bf16" else torch.float32)
w = syn.WORKLOADS[wl]
B = w.B if len(sys.argv) < 4 else int(sys.argv[3])
etxt, obj, tnum, onum = syn.ot_inputs(B, w.M, w.N, w.D, 0, "ragged", dtype=dt)
etxt, obj, tnum, onum = etxt.cuda(), obj.cuda(), tnum.cuda(), onum.cuda()
def fwd():
    with torch.no_grad():
        return F_.ot_alignment(etxt, obj, tnum, onum)
eg, og = etxt.clone().requires_grad_(True), obj.clone().requires_grad_(True)
def full():
    eg.grad = None; og.grad = None
    l, _ = F_.ot_alignment(eg, og, tnum, onum)
    l.backward()
def time_it(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s): fn()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    with torch.cuda.graph(g): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
l, d = fwd()
print("%s %s B=%d: fwd(cost+ipot) %.1f us   full %.1f us   loss %.6f" % (
    wl, sys.argv[2], B, time_it(fwd), time_it(full), l.item()))
# packed (variable-length) layout of the same batch: SURVEY 8f-3
if dt == torch.bfloat16 and w.M <= 16 and w.N <= 64 and w.D <= 512:
    tp = F_.pack_nodes(etxt, tnum)
    ip = F_.pack_nodes(obj, onum, drop_first=True)
    trg, irg = tp.rows.clone().requires_grad_(True), ip.rows.clone().requires_grad_(True)
    tpg, ipg = F_.PackedNodes(trg, tp.offsets, tp.max_count), F_.PackedNodes(irg, ip.offsets, ip.max_count)
    def full_packed():
        trg.grad = None; irg.grad = None
        l, _ = F_.ot_alignment_packed(tpg, ipg)
        l.backward()
    lp, _ = F_.ot_alignment_packed(tp, ip)
    nbytes = 2 * 2 * (tp.rows.numel() + ip.rows.numel())
    tpk = time_it(full_packed)
    print("%s packed: full %.1f us  (%d of %d node rows are valid: %.0f MB in + out -> %.0f GB/s)  loss %.6f" % (
        wl, tpk, tp.rows.shape[0] + ip.rows.shape[0], B * (w.M + w.N), nbytes / 1e6, nbytes / tpk / 1e3, lp.item()))
