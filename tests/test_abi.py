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
    assert _lib.lib().apr_abi_version() == 2
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


def test_train_layout_is_consistent_without_a_gpu():
    """apr_train_layout / apr_train_workspace_bytes are host arithmetic (no CUDA call): the regions the sharded driver
    broadcasts (apr_b200/distributed.py) must lie inside the workspace, be 256-byte aligned and not overlap."""
    import ctypes

    from apr_b200 import _lib
    L = _lib.lib()
    for S, B, d in [(1, 1, 4), (3, 100, 20), (16, 512, 64), (64, 65536, 128), (4, 1 << 20, 256)]:
        out = (ctypes.c_int64 * 13)()
        assert L.apr_train_layout(S, B, d, out) == 0
        total, Sc = int(out[0]), int(out[1])
        assert total == L.apr_train_workspace_bytes(S, B, d) and 1 <= Sc <= S
        names = ["ucnt", "icnt", "iall", "nslow", "seg_hdr", "rec", "iu_item", "hdr", "npair", "nfast", "pairs"]
        off = dict(zip(names, [int(v) for v in out[2:13]]))
        size = {"ucnt": 4 * S, "icnt": 4 * S, "iall": 4 * S, "nslow": 4 * S, "npair": 4 * S, "nfast": 4 * S, "hdr": 256,
                "seg_hdr": 32 * B * S, "rec": 16 * B * S, "iu_item": 4 * B * S, "pairs": 48 * (B // 2 + 1) * S}
        spans = sorted((off[k], off[k] + size[k], k) for k in names)
        for (a0, a1, ka), (b0, b1, kb) in zip(spans, spans[1:]):
            assert a1 <= b0, (ka, kb)
        for a0, a1, k in spans:
            assert a0 % 256 == 0 and 0 <= a0 and a1 <= total, k
    assert L.apr_train_workspace_bytes(0, 1, 4) == -1 and L.apr_train_workspace_bytes(1, 1, 3) == -1
