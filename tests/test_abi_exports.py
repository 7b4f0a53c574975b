"""CPU-side check: the built C-ABI library loads and exports every symbol include/boss_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "boss_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(boss_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    import boss_b200  # noqa: F401  builds nothing; raises if the .so is missing
    from boss_b200 import _lib
    names = _declared()
    assert len(names) >= 20
    so = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(so, n), f"{n} declared in include/boss_b200.h but not exported"
    for n in _lib.EXPORTS:
        assert n in names, f"{n} bound in _lib.py but not declared in the header"


def test_no_compute_without_gpu_is_loud():
    import torch
    if torch.cuda.is_available():
        return
    from boss_b200 import _lib
    import pytest
    with pytest.raises(_lib.BossError):
        _lib.init(0)
