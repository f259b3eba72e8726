"""CPU-only: the C-ABI library builds for sm_100a, loads, and exports exactly the entry points
include/thevc_cuda.h declares.  No compute call is made (there is no GPU here and no fallback)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "thevc_cuda.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(tvc_[a-z0-9_A-Z]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from thevc_b200 import capi
    from thevc_b200 import build as b
    b.build()
    assert os.path.exists(capi.lib_path())
    L = capi.load()
    decl = declared_symbols()
    assert len(decl) >= 35
    for name in decl:
        assert hasattr(L, name), "library does not export %s" % name
    # the binding table covers the header, and nothing else
    assert sorted(capi.SIGNATURES) == decl
    assert L.tvc_abi_version() == 1


def test_library_is_sm100a_and_has_native_kernels():
    from thevc_b200 import capi
    out = subprocess.run(["cuobjdump", "-lelf", capi.lib_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", "_ZN3tvc15k_me_sad_tablesILi8ELi2ELb1EEEvNS_6MeMapsEiiiiPK13tvc_me_centerPtii",
                           capi.lib_path()], capture_output=True, text=True).stdout
    assert "VABSDIFF4" in sass          # u8 SIMD SAD
    assert "UTMALDG" in sass            # TMA staging of the search window


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from thevc_b200 import TLibCuda, TvcError
    with pytest.raises(TvcError):
        TLibCuda(416, 240)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "thevc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in src and "hm_oracle" not in src and "libhmref" not in src, f
