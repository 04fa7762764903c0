"""The C-ABI library builds for sm_100a, loads on a CPU-only box, exports every symbol that
include/ehic_b200.h declares, and refuses to compute without a device (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "ehic_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ehic_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    from ensemble_hic_b200 import _lib, build
    path = build.build()
    exported = subprocess.check_output(["nm", "-D", "--defined-only", path], text=True)
    names = {l.split()[-1] for l in exported.splitlines() if l.strip()}
    declared = _declared()
    assert len(declared) >= 20
    missing = [s for s in declared if s not in names]
    assert not missing, missing
    assert sorted(_lib.SYMBOLS) == declared          # the ctypes binding covers the whole header
    L = _lib.lib()
    for s in declared:
        assert hasattr(L, s)
    assert L.ehic_version() >= 100


def test_library_is_built_for_sm_100a_only():
    from ensemble_hic_b200 import build
    out = subprocess.check_output(["cuobjdump", "--list-elf", build.build()], text=True)
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback():
    """without a compute-capability-10 device every compute entry point fails loudly"""
    from ensemble_hic_b200 import _lib
    L = _lib.lib()
    if L.ehic_device_count() > 0:
        pytest.skip("a B200 is present")
    from ensemble_hic_b200.engine import Model
    with pytest.raises(_lib.EhicError) as e:
        Model(1, np.ones(4))
    assert e.value.code == _lib.ERR_NODEVICE
    n = C.c_int64()
    x = np.zeros((4, 3))
    rc = L.ehic_nblist_host(4, x.ctypes.data, 1.0, None, 0, C.byref(n), 0)
    assert rc == _lib.ERR_NODEVICE
    assert b"no CPU fallback" in L.ehic_last_error()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "ensemble_hic_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "ehic_oracle" not in txt, f
