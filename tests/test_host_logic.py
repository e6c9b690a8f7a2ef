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


def test_criterion_contrastive_accepts_dense_logits_but_has_no_cpu_path():
    # model_clip.py:633-662 takes any logits tensors; the drop-in does too (CUDA row kernels) -- on CPU
    # tensors it raises instead of falling back
    crit = ce.CriterionContrastive("ce")
    with pytest.raises(RuntimeError, match="no CPU path"):
        crit(torch.randn(2, 4), torch.randn(4, 2), torch.tensor([0, 2]), torch.tensor([0, 0, 1, 1]), index_pos=torch.tensor([0, 2]))
    with pytest.raises(RuntimeError, match="global negatives need the features"):
        ce.CriterionContrastive("ce", group=True)(torch.randn(2, 4), torch.randn(4, 2), index_pos=torch.tensor([0, 2]))


def test_alignment_masks_are_num_semantics_for_every_dtype():
    # the reference applies mask2pad(x) = (x == 0) whatever the dtype (model_clip.py:673-676,688-690)
    from clip_event_b200 import functional as F_
    b = torch.tensor([[True, False]])
    assert F_.num_mask(b).dtype == torch.int64 and F_.num_mask(b).tolist() == [[1, 0]]
    i = torch.tensor([[1, 0]])
    assert F_.num_mask(i) is i


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


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm) prints one JSON
    line with the contract's keys; run here on the smallest workload."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "c1",
                        "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "loss fwd+bwd samples/sec" and line["unit"] == "samples/s"
    assert line["higher_is_better"] is True and line["n_gpus"] >= 1 and line["value"] > 0
    assert line["e2e"] == {"value": line["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == line["value"] and "sample" in cb
    assert "workload" in line["config"]
    # both arms print the same workload description
    sys.path.insert(0, root)
    import bench
    from clip_event_b200 import synthetic as syn_
    assert line["config"] == bench.config_dict(syn_.WORKLOADS["c1"], 1, "bf16")
