"""The C-ABI shared library loads and exports every symbol include/scone_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "scone_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", " ", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(sb_[a-z_0-9]+)\s*\(", txt)))


def test_header_declares_the_boundary():
    syms = declared_symbols()
    for s in ("sb_create", "sb_load_geometry", "sb_load_mg_data", "sb_define_tallies", "sb_run_cycle", "sb_resample",
              "sb_tally_read", "sb_bank_upload", "sb_bank_download", "sb_source_generate", "sb_destroy", "sb_last_error"):
        assert s in syms


def test_library_exports_every_declared_symbol():
    import scone_b200
    path = scone_b200.library_path()
    assert os.path.exists(path), "libscone_b200.so not built: run __graft_entry__.build()"
    lib = ctypes.CDLL(path)
    for s in declared_symbols():
        assert hasattr(lib, s), "missing export: " + s


def test_no_cpu_fallback():
    """Without a CUDA device the engine refuses to come up (it never routes through the oracle)."""
    import scone_b200
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(scone_b200.EngineError):
        scone_b200.EigenPhysicsPackage(os.path.join(ROOT, "decks", "c5g7", "c5g7_2d"))


def test_product_does_not_reference_oracle():
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "scone_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", ".sh")):
                t = open(os.path.join(base, f), errors="ignore").read()
                if re.search(r'#include\s+"[^"]*oracle|import\s+oracle|from\s+tests|oracle_lib|liboracle', t):
                    bad.append(f)
    assert not bad, bad
