"""The C-ABI library: builds for sm_100a, loads, exports every symbol that
include/swirld_b200.h declares, and refuses to run without a GPU (no CPU
fallback).  No compute calls here."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "swirld_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sw_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from swirld_b200 import build, engine
    build.build()
    L = engine.load_library()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), "libswirld_b200.so does not export %s" % n
    assert sorted(engine.SYMBOLS) == names, "engine.SYMBOLS out of date with the header"
    assert L.sw_version() >= 100


def test_sass_is_sm100a():
    import shutil
    import subprocess
    from swirld_b200 import build
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", build.LIB], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from swirld_b200 import engine
    with pytest.raises(engine.EngineError) as ei:
        engine.Engine(4, 16)
    assert ei.value.code == -4


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "py-swirld_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                s = open(os.path.join(dp, f)).read()
                assert "import oracle" not in s and "liboracle" not in s and "ref_harness" not in s, f
