"""Programmatic dependent launch along the contrastive kernel chain (csrc/ce_common.cuh, CE_LAUNCH_CHAIN): the same
loss-head step captured as a CUDA graph without and with the programmatic edges must give the same losses and
gradients on every replay, and a step run eagerly with CE_PDL=0 must match the default (tools/pdl_check.py)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("workload,dtype", [("c2", "bf16"), ("c2", "fp32")])
def test_pdl_graph_replays_match_plain_graph(workload, dtype):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "pdl_check.py"), workload, dtype, "12"],
                       capture_output=True, text=True, timeout=420, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "pdl_check: OK" in r.stdout


def test_eager_step_same_with_and_without_pdl(monkeypatch):
    import clip_event_b200 as ce
    from clip_event_b200 import synthetic as syn

    w = syn.WORKLOADS["c2"]
    img, txt, ls = syn.contrastive_inputs(w.B, w.T, w.D, 11, "trained", dtype=torch.bfloat16)
    lpi, lpt, idx = (t.cuda() for t in syn.contrastive_labels(w.B, w.T))
    crit = ce.CriterionContrastive("ce")
    head = ce.ClipEventHead().cuda()

    def run(mode):
        monkeypatch.setenv("CE_PDL", mode)       # read by the library at every launch
        i = img.cuda().requires_grad_(True)
        t = txt.cuda().requires_grad_(True)
        head.logit_scale.grad = None
        li, lt = head(i, t)
        out = crit(li, lt, lpi, lpt, index_pos=idx, constrastive_overbatch=head.constrastive_overbatch)
        (out["loss_i"] + out["loss_t"]).backward()
        torch.cuda.synchronize()
        return [out["loss_i"].float(), out["loss_t"].float(), i.grad.float(), t.grad.float(), head.logit_scale.grad.float()]

    a, b = run("0"), run("1")
    for x, y in zip(a, b):
        scale = x.abs().max().clamp_min(1e-30)
        assert float(((x - y).abs().max() / scale).item()) <= 2e-2   # run-to-run accumulation order only
