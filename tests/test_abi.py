"""The C-ABI library builds, loads without a GPU and exports every symbol include/apr_b200.h declares."""
import ctypes
import os
import re

from apr_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "apr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(apr_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert _declared() == sorted(_lib.EXPORTED_SYMBOLS)


def test_library_builds_and_exports_all_symbols():
    path = _lib.build_library()
    L = ctypes.CDLL(path)
    for name in _declared():
        assert hasattr(L, name), name
    assert _lib.lib().apr_abi_version() == 1
    assert _lib.lib().apr_status_string(3) == b"workspace too small"


def test_workspace_size_queries_need_no_gpu():
    L = _lib.lib()
    assert L.apr_train_workspace_bytes(4, 512, 64) > 2 * 2 * 512 * 64 * 4
    assert L.apr_train_workspace_bytes(4, 512, 66) == -1  # d must be a multiple of 4
    assert L.apr_eval_workspace_bytes(100, 10, 64) > 0
    assert L.apr_eval_workspace_bytes(100, 129, 64) == -1


def test_no_cpu_fallback():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from apr_b200 import engine
    with pytest.raises(RuntimeError):
        engine.require_cuda()
