"""Host-side mirror of the reference interface: argument handling and error behaviour (CPU only)."""
import pytest
import torch

import clip_event_b200 as ce
from clip_event_b200 import synthetic as syn
from clip_event_b200.model_clip import LazyLogits


def test_criterion_contrastive_rejects_unknown_loss_like_the_reference():
    # model_clip.py:631
    with pytest.raises(RuntimeError, match="Invalid constrastive_loss"):
        ce.CriterionContrastive("hinge")
    for ok in ("ce", "bce", "kl"):
        ce.CriterionContrastive(ok)


def test_criterion_contrastive_needs_lazy_logits():
    crit = ce.CriterionContrastive("ce")
    with pytest.raises(RuntimeError, match="LazyLogits"):
        crit(torch.randn(2, 4), torch.randn(4, 2), index_pos=torch.tensor([0, 2]))


def test_lazy_logits_shapes():
    img, txt = torch.randn(4, 8), torch.randn(12, 8)
    ls = torch.tensor(1.0)
    assert LazyLogits(img, txt, ls, "per_image").shape == (4, 12)
    assert LazyLogits(txt, img, ls, "per_text").shape == (12, 4)
    assert LazyLogits(img, txt, ls, "per_image", per_instance=True).shape == (4, 3)
    assert LazyLogits(img, txt, ls, "per_image").size(1) == 12


def test_head_parameters_and_flags():
    head = ce.ClipEventHead(constrastive_overbatch=False, alignment=True)
    assert abs(head.logit_scale.item() - syn.LOGIT_SCALE_INIT) < 1e-6      # model_clip.py:330
    assert head.constrastive_overbatch is False and head.alignment is True
    head.set_hyps(True, False)
    assert head.constrastive_overbatch is True and head.alignment is False
    a, b = torch.randn(2, 3, 8), torch.randn(2, 5, 8)
    assert head.sim_entity(a, b) == (a, b)


def test_mask2pad():
    m = torch.tensor([[1, 1, 0], [0, 0, 0]])
    assert ce.CriterionAlignment().mask2pad(m).tolist() == [[False, False, True], [True, True, True]]


def test_synthetic_generators_are_seeded_and_shaped():
    a = syn.contrastive_inputs(4, 3, 16, 7, "trained")
    b = syn.contrastive_inputs(4, 3, 16, 7, "trained")
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    assert a[0].shape == (4, 16) and a[1].shape == (12, 16)
    t, o, tn, on = syn.ot_inputs(5, 4, 6, 8, 3, "edge")
    assert t.shape == (5, 4, 8) and o.shape == (5, 7, 8) and tn.shape == (5, 4) and on.shape == (5, 7)
    assert tn[0].sum() == 0 and on[1, 1:].sum() == 0 and on[:, 0].all()
    for w in syn.WORKLOADS.values():
        assert w.T == w.K + 1


def test_tile_walk_covers_every_work_item_once(tmp_path):
    """The persistent GEMM kernels' division-free work walk (umma_gemm.cuh: TileWalk), compiled for
    the host: over 720 geometries (row units x column blocks x K splits x workers x both raster
    orders) every item is visited exactly once and decoded to the right tile."""
    import os
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "tilewalk_host_test.cu")
    exe = str(tmp_path / "tilewalk_host_test")
    r = subprocess.run([nvcc, "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-o", exe, src],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "tilewalk ok" in r.stdout, r.stdout + r.stderr
