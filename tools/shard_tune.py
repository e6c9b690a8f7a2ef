"""Per-rank compute of the column-sharded contrastive step WITHOUT collectives (single GPU):
what one of W ranks executes at global batch B (R = B gathered images, C = B*T/W local columns)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from clip_event_b200 import distributed as cd, synthetic as syn
wl, W = sys.argv[1], int(sys.argv[2])
w = syn.WORKLOADS[wl]; dt = torch.bfloat16
img, txt, ls = syn.contrastive_inputs(w.B, w.T, w.D, 0, "trained", dtype=dt)
b = w.B // W
dev = torch.device("cuda")
li, lt, ip = cd.global_labels_for_rank(b, w.T, W, 0, device=dev)
img_all = img.to(dev); txt_l = txt[: b * w.T].to(dev); lsd = ls.to(dev).float().reshape(1)
lab_all = torch.arange(w.B, device=dev) * w.T
be = cd.CudaBackend()
one = torch.ones(1, device=dev)
def fwd():
    stats, state = be.fwd_partial(img_all, txt_l, lsd, lab_all, lt, ip, 0)
    sa = stats.repeat(W, 1)   # stand-in for the gathered records (values irrelevant for timing)
    return be.fwd_finish(sa, W, state), state
def full():
    (_, _), state = fwd()
    dtxt, dimg_hat, dls = be.bwd_partial(img_all, txt_l, lsd, lab_all, lt, ip, 0, one, one, w.B, w.B, state)
    return be.bwd_finish(img_all[:b], dimg_hat[:b])
def time_it(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
print("%s W=%d per-rank compute: fwd %.1f us  fwd+bwd %.1f us" % (wl, W, time_it(lambda: fwd()), time_it(full)))
