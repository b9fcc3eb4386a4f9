"""The C-ABI library loads, exports every symbol include/fedfr_b200.h declares, and fails loudly without a GPU."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def native():
    import __graft_entry__ as g
    g.build()
    from fedfr_b200 import _native
    return _native


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "fedfr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b((?:pfc|fedavg)_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(native):
    syms = _declared_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(native.lib, s), f"{s} declared in include/fedfr_b200.h but not exported"
        assert s in native.SIGNATURES, f"{s} has no ctypes signature in fedfr_b200/_native.py"
    assert native.lib.pfc_version() >= 100


def test_every_developer_hook_is_declared_and_exported(native):
    """include/fedfr_b200_dev.h lists the tuning / profiling hooks; every pfc_set_* / pfc_profile_* / pfc_launch_count the
    library exports is declared in one of the two headers (no undeclared exports)."""
    import subprocess
    text = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "fedfr_b200_dev.h")).read(), flags=re.S)
    dev = set(re.findall(r"\b(pfc_[a-z0-9_]+)\s*\(", text))
    assert len(dev) >= 15
    for s in dev:
        assert hasattr(native.lib, s), f"{s} declared in include/fedfr_b200_dev.h but not exported"
    out = subprocess.run(["nm", "-D", "--defined-only", native.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T ((?:pfc|fedavg)_[a-z0-9_]+)$", out, flags=re.M))
    undeclared = exported - dev - set(_declared_symbols())
    assert not undeclared, f"exported but declared in neither header: {sorted(undeclared)}"


def test_pure_host_entries(native):
    assert native.lib.fedavg_table_bytes(475, 40) > 475 * 40 * 8
    assert native.lib.pfc_sample_workspace_bytes(250000) > 1024
    assert native.lib.pfc_fwd_num_partials(0, 0, 512, 0) == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_gpu_fails_loudly(native):
    rc = native.lib.pfc_query_device(0, None, None, None)
    assert rc != 0
    with pytest.raises(RuntimeError):
        native.check(rc, "pfc_query_device")
    import fedfr_b200
    with pytest.raises(RuntimeError):
        fedfr_b200.PartialFC(0, 0, 1, 4, False, fedfr_b200.CosFace(), 10)
    with pytest.raises(RuntimeError):
        fedfr_b200.FedPavg([{"a": torch.zeros(4)}], [1.0])


def test_unsupported_margin_is_an_error():
    from fedfr_b200.losses import margin_params, CosFace
    assert margin_params(CosFace(s=30.0, m=0.4)) == (30.0, pytest.approx(0.4), 0)

    class ArcFace:                    # the reference's class is matched by name and read as a descriptor
        s, m = 64.0, 0.5
    assert margin_params(ArcFace()) == (64.0, 0.5, 1)

    class CurricularFace:
        s, m = 64.0, 0.5
    with pytest.raises(NotImplementedError):
        margin_params(CurricularFace())


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under fedfr_b200/ may import, load or execute it."""
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|oracle[/\\.]_?(build|ref)|liboracle|oracle_c", re.M)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "fedfr_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                assert not pat.search(open(os.path.join(dirpath, f)).read()), f
