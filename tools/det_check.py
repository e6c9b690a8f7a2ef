import sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from clip_event_b200 import functional as F_, synthetic as syn
w = syn.WORKLOADS["c2"]
img, txt, ls = syn.contrastive_inputs(w.B, w.T, w.D, 5, "trained", dtype=torch.bfloat16)
lpi, lpt, idx = syn.contrastive_labels(w.B, w.T)
def run():
    i = img.cuda().requires_grad_(True); t = txt.cuda().requires_grad_(True); l = ls.float().cuda().requires_grad_(True)
    li, lt = F_.contrastive_over_batch(i, t, l, lpi.cuda(), lpt.cuda(), idx.cuda())
    (li + lt).backward()
    torch.cuda.synchronize()
    return i.grad.clone(), t.grad.clone(), l.grad.clone()
a = run(); b = run(); c = run()
for x, y in zip(a, b): print("run1 vs run2 equal:", torch.equal(x, y), (x.float() - y.float()).abs().max().item())
for x, y in zip(a, c): print("run1 vs run3 equal:", torch.equal(x, y))
d = (a[0].float() - b[0].float())
nz = d.nonzero()
print("differing elements:", nz.shape[0], "of", d.numel(), "rows:", sorted(set(nz[:, 0].tolist()))[:20])
rel = (d.abs() / a[0].float().abs().clamp_min(1e-30))
print("max rel diff %.3e  median |grad| %.3e  max|d| at value %.3e" % (rel.max().item(), a[0].float().abs().median().item(), a[0].float().flatten()[d.abs().flatten().argmax()].item()))
