"""The C-ABI library must build for sm_100a without a GPU, load, and export every symbol include/t2v_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "t2v_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(t2v_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    from t2v_b200 import native
    path = native.build()
    lib = ctypes.CDLL(path)
    names = declared_symbols()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted(native.EXPORTS) == names, set(native.EXPORTS) ^ set(names)
    assert lib.t2v_version() == 2


def test_loader_has_no_fallback(monkeypatch, tmp_path):
    from t2v_b200 import native
    monkeypatch.setattr(native, "LIB_PATH", str(tmp_path / "missing.so"))
    monkeypatch.setattr(native, "_lib", None)
    try:
        native.lib()
    except RuntimeError as e:
        assert "missing" in str(e)
    else:
        raise AssertionError("loading a missing library must raise")
