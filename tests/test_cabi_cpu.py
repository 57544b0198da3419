"""The C-ABI library loads and exports every symbol include/mtp_b200.h declares (no compute calls; CPU only)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "mtp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mtp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from mtp_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        from mtp_b200 import build
        build.build()
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 5
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/mtp_b200.h but not exported"
    # every symbol the Python binding uses is declared in the header
    for name in _lib.exported_symbols():
        assert name in declared, f"{name} bound in _lib.py but missing from include/mtp_b200.h"
    assert lib.mtp_version() >= 100


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from mtp_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.MtpError):
        _lib.load()
