import os, sys, torch
sys.path.insert(0, "/root/repo")
from clip_event_b200 import functional as F_, synthetic as syn
w = syn.WORKLOADS["c3"]
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
for masks in ("full", "ragged"):
    etxt, obj, tnum, onum = syn.ot_inputs(w.B, w.M, w.N, w.D, 0, masks, dtype=torch.bfloat16)
    etxt, obj, tnum, onum = etxt.cuda(), obj.cuda(), tnum.cuda(), onum.cuda()
    eg, og = etxt.clone().requires_grad_(True), obj.clone().requires_grad_(True)
    def full():
        eg.grad = None; og.grad = None
        l, _ = F_.ot_alignment(eg, og, tnum, onum); l.backward()
    tp = F_.pack_nodes(etxt, tnum); ip = F_.pack_nodes(obj, onum, drop_first=True)
    trg, irg = tp.rows.clone().requires_grad_(True), ip.rows.clone().requires_grad_(True)
    tpg, ipg = F_.PackedNodes(trg, tp.offsets, tp.max_count), F_.PackedNodes(irg, ip.offsets, ip.max_count)
    def fullp():
        trg.grad = None; irg.grad = None
        l, _ = F_.ot_alignment_packed(tpg, ipg); l.backward()
    def fwdp():
        with torch.no_grad(): F_.ot_alignment_packed(tp, ip)
    def fwd():
        with torch.no_grad(): F_.ot_alignment(etxt, obj, tnum, onum)
    print(masks, "padded fwd %.1f full %.1f | packed fwd %.1f full %.1f  rows %d" % (time_it(fwd), time_it(full), time_it(fwdp), time_it(fullp), tp.rows.shape[0] + ip.rows.shape[0]))
