"""The C-ABI shared library: builds for sm_100a, loads, exports every symbol the header declares,
and refuses to run without a B200 (no CPU fallback).  CPU only."""
import os
import re
import subprocess

import pytest
import torch

from clip_event_b200 import _lib as L
from clip_event_b200 import build as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "clip_event_b200.h")


def _declared():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ce_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_loads():
    path = B.build()
    assert os.path.exists(path)
    lib = L.load(build_if_missing=False)
    assert lib.ce_version() >= 100


def test_every_declared_symbol_is_exported_and_bound():
    names = _declared()
    assert len(names) >= 15
    out = subprocess.run(["nm", "-D", "--defined-only", L.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (ce_[a-z0-9_]+)", out))
    missing = [n for n in names if n not in exported]
    assert not missing, "declared in the header but not exported: %s" % missing
    unbound = [n for n in names if n not in L.SIGNATURES]
    assert not unbound, "declared in the header but not bound in _lib.SIGNATURES: %s" % unbound
    extra = [n for n in L.SIGNATURES if n not in names]
    assert not extra, "bound but not declared in the header: %s" % extra


def test_built_for_sm100a_with_tcgen05_and_tma():
    sass = subprocess.run(["cuobjdump", "-sass", L.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    assert "UTCHMMA" in sass, "tcgen05.mma missing from SASS"
    assert "UTMALDG" in sass, "TMA loads missing from SASS"
    assert "LDTM" in sass, "tcgen05.ld missing from SASS"
    assert "UTCHMMA.2CTA" in sass, "tcgen05.mma.cta_group::2 (CTA-pair GEMM) missing from SASS"
    assert "UTMASTG" in sass, "TMA stores (gradient-tile epilogue) missing from SASS"
    # programmatic dependent launch along the contrastive chain: griddepcontrol.wait / .launch_dependents
    assert "ACQBULK" in sass and "PREEXIT" in sass, "griddepcontrol (programmatic dependent launch) missing from SASS"


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    lib = L.load()
    assert lib.ce_device_check() == -5          # CE_ERR_ARCH
    assert "no CPU fallback" in L.last_error() or "sm_" in L.last_error()
    import clip_event_b200 as ce
    x = torch.randn(2, 3, 8)
    with pytest.raises(RuntimeError):
        ce.cost_matrix_cosine(x, x)
    with pytest.raises(RuntimeError):
        ce.CriterionAlignment()(x, x, torch.ones(2, 3, dtype=torch.int64), torch.ones(2, 3, dtype=torch.int64))
    head = ce.ClipEventHead()
    with pytest.raises(RuntimeError):
        head(torch.randn(2, 8), torch.randn(4, 8))
