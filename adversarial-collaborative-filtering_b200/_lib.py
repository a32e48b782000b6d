"""ctypes binding of libapr_b200.so (the C ABI declared in include/apr_b200.h) and its build recipe.

There is no CPU fallback: if the shared library is missing or cannot be loaded, ``lib()`` raises.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import threading
from ctypes import POINTER, c_char_p, c_double, c_float, c_int32, c_int64, c_uint32, c_void_p

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_DIR = os.path.dirname(PKG_DIR)
CSRC_DIR = os.path.join(PKG_DIR, "csrc")
LIB_DIR = os.path.join(PKG_DIR, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libapr_b200.so")
SOURCES = ["api.cu", "train.cu", "sampler.cu", "eval.cu", "eval_tc.cu", "loader.cu", "keras_mf.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
              "-shared"]

APR_OK, APR_E_ARG, APR_E_ALIGN, APR_E_WORKSPACE, APR_E_CUDA, APR_E_UNSUPPORTED = range(6)


class AprError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__("libapr_b200 status %d: %s" % (status, msg))
        self.status = status


def _sources():
    return [os.path.join(CSRC_DIR, s) for s in SOURCES if os.path.exists(os.path.join(CSRC_DIR, s))]


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = _sources() + [os.path.join(CSRC_DIR, f) for f in os.listdir(CSRC_DIR) if f.endswith((".cuh", ".h"))]
    deps.append(os.path.join(REPO_DIR, "include", "apr_b200.h"))
    return any(os.path.getmtime(p) > t for p in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    """nvcc-compile every CUDA source for sm_100a into the in-tree shared library (one object per source, compiled in
    parallel; an object is rebuilt when its source or any header is newer)."""
    if not force and not needs_build():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    obj_dir = os.path.join(LIB_DIR, "obj")
    os.makedirs(obj_dir, exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")
    flags = [f for f in NVCC_FLAGS if f != "-shared"]
    hdrs = [os.path.join(CSRC_DIR, f) for f in os.listdir(CSRC_DIR) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(REPO_DIR, "include", "apr_b200.h"))
    t_hdr = max(os.path.getmtime(h) for h in hdrs)
    jobs, objs = [], []
    for src in _sources():
        obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(t_hdr, os.path.getmtime(src)):
            cmd = [nvcc] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
            jobs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)))
    for cmd, proc in jobs:
        _, err = proc.communicate()
        if proc.returncode != 0:
            raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), err[-4000:]))
        if verbose:
            print(err)
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH] + objs
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (" ".join(cmd), res.stderr[-4000:]))
    return LIB_PATH


_P = c_void_p  # device pointers travel as integers

_SIGNATURES = {
    "apr_abi_version": (ctypes.c_int, []),
    "apr_status_string": (c_char_p, [ctypes.c_int]),
    "apr_last_cuda_error": (c_char_p, []),
    "apr_device_info": (ctypes.c_int, [POINTER(c_int32), POINTER(c_int32), POINTER(c_int32)]),
    "apr_context_create": (ctypes.c_int, [c_int32]),
    "apr_context_destroy": (ctypes.c_int, [c_int32]),
    "apr_context_stats": (ctypes.c_int, [POINTER(c_int64)]),
    "apr_init_truncated_normal": (ctypes.c_int, [_P, c_int64, c_int32, c_float, c_uint32, c_uint32, c_uint32, _P]),
    "apr_fill_f32": (ctypes.c_int, [_P, c_int64, c_float, _P]),
    "apr_sample_epoch": (ctypes.c_int, [_P, _P, c_int64, c_int32, c_int32, _P, _P, c_int32, c_uint32, c_uint32, c_int32,
                                        _P, _P, _P, _P, _P, _P]),
    "apr_sample_epoch_shard": (ctypes.c_int, [_P, _P, c_int64, c_int32, c_int32, _P, _P, c_int32, c_uint32, c_uint32, c_int32,
                                              c_int32, c_int32, c_int32, _P, _P, _P, _P, _P, _P]),
    "apr_select_dns": (ctypes.c_int, [_P, _P, c_int32, _P, _P, c_int64, c_int32, _P, _P]),
    "apr_train_workspace_bytes": (c_int64, [c_int32, c_int32, c_int32]),
    "apr_train_workspace_init": (ctypes.c_int, [_P, c_int64, _P]),
    "apr_train_steps": (ctypes.c_int, [_P, _P, _P, _P, c_int64, c_int64, c_int32, _P, _P, _P, c_int32, c_int32, c_float,
                                       c_float, c_float, c_float, c_int32, c_int32, _P, c_int64, _P, _P]),
    "apr_train_steps_random": (ctypes.c_int, [_P, _P, _P, _P, c_int64, c_int64, c_int32, _P, _P, _P, c_int32, c_int32,
                                              c_float, c_float, c_float, c_float, c_uint32, c_uint32, _P, c_int64, _P, _P]),
    "apr_train_status": (ctypes.c_int, [_P, POINTER(c_int32), _P]),
    "apr_train_prepare": (ctypes.c_int, [_P, _P, _P, c_int32, c_int32, c_int32, c_int64, c_int64, _P, c_int64, _P]),
    "apr_train_run": (ctypes.c_int, [_P, _P, _P, _P, c_int64, c_int64, c_int32, _P, _P, _P, c_int32, c_int32, c_float,
                                     c_float, c_float, c_float, c_int32, c_int32, _P, c_int64, _P, _P]),
    "apr_train_layout": (ctypes.c_int, [c_int32, c_int32, c_int32, POINTER(c_int64)]),
    "apr_train_prepare_range": (ctypes.c_int, [_P, _P, _P, c_int32, c_int32, c_int32, c_int64, c_int64, _P, c_int64,
                                               c_int32, c_int32, c_int32, _P]),
    "apr_train_stage_sharded": (ctypes.c_int, [POINTER(c_void_p)] * 6 + [c_int32, c_int32, c_int32, c_int32, c_int32,
                                               c_float, c_float, c_float, c_float, c_int32, _P, c_int64, _P, c_int32,
                                               c_int32, _P]),
    "apr_train_steps_sharded": (ctypes.c_int, [POINTER(c_void_p)] * 7 + [c_int32, c_int32, c_int32, c_int32, c_int32,
                                               c_float, c_float, c_float, c_float, c_int32, _P, c_int64, _P, c_int32,
                                               c_int32, _P, _P]),
    "apr_train_unique_counts": (ctypes.c_int, [_P, c_int64, c_int32, c_int32, c_int32, POINTER(c_int32), _P]),
    "apr_loss_acc": (ctypes.c_int, [_P, _P, c_int32, _P, _P, _P, c_int32, c_int32, _P, _P]),
    "apr_score_pairs": (ctypes.c_int, [_P, _P, c_int32, _P, _P, c_int64, _P, _P]),
    "apr_score_all_items": (ctypes.c_int, [_P, _P, c_int32, _P, c_int32, c_int32, c_int32, _P, _P]),
    "apr_eval_candidates": (ctypes.c_int, [_P, _P, c_int32, _P, _P, _P, c_int32, _P, _P, _P]),
    "apr_eval_workspace_bytes": (c_int64, [c_int32, c_int32, c_int32]),
    "apr_eval_fullrank": (ctypes.c_int, [_P, _P, c_int32, _P, _P, c_int32, c_int32, c_int32, _P, _P, c_int32, _P, _P, _P,
                                         c_int32, _P, c_int64, _P]),
    "apr_eval_tc_workspace_bytes": (c_int64, [c_int32, c_int32, c_int32]),
    "apr_eval_fullrank_tc": (ctypes.c_int, [_P, _P, c_int32, _P, _P, c_int32, c_int32, c_int32, _P, _P, _P, _P, c_int64, _P,
                                            _P]),
    "apr_eval_tc_topk_workspace_bytes": (c_int64, [c_int32, c_int32, c_int32, c_int32]),
    "apr_eval_fullrank_tc_topk": (ctypes.c_int, [_P, _P, c_int32, _P, _P, c_int32, c_int32, c_int32, _P, _P, c_int32, _P, _P,
                                                 _P, ctypes.c_uint64, _P, _P, c_int64, _P, _P]),
    "apr_eval_tc_ambiguous": (ctypes.c_int, [_P, c_int32, c_int32, c_int32, c_int32, POINTER(c_int32), _P]),
    "apr_topk_merge": (ctypes.c_int, [_P, _P, c_int32, c_int32, c_int32, _P, _P, _P]),
    "apr_eval_tc_timing": (ctypes.c_int, [c_int32, POINTER(c_float)]),
    "apr_sum_squares": (ctypes.c_int, [_P, c_int64, _P, _P]),
    "apr_keras_step": (ctypes.c_int, [_P] * 8 + [c_int64, c_int64, c_int32, _P, _P, _P, _P, c_int32, c_int32, c_float, c_float,
                                      c_float, c_int64, _P, _P]),
    "apr_loader_workspace_bytes": (c_int64, [c_int64, c_int64]),
    "apr_tsv_count_lines": (ctypes.c_int, [_P, c_int64, _P, c_int64, POINTER(c_int64), _P]),
    "apr_tsv_parse": (ctypes.c_int, [_P, c_int64, _P, c_int64, c_int32, _P, _P, _P, _P, _P, _P, _P, _P]),
    "apr_loader_train_rows": (ctypes.c_int, [_P, c_int64, c_int32, _P, _P, _P, c_int64, _P]),
    "apr_loader_csr": (ctypes.c_int, [_P, _P, c_int64, c_int64, _P, _P, _P, _P, c_int64, _P]),
    "apr_loader_unique_pairs": (ctypes.c_int, [_P, _P, _P, c_int64, _P, _P, _P, _P, c_int64, _P]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES.keys())

_lock = threading.Lock()
_lib = None


def lib() -> ctypes.CDLL:
    """Load (once) the in-tree CUDA library.  Raises if it has not been built -- no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libapr_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` -- "
                "this package has no CPU fallback." % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        if L.apr_abi_version() != 2:
            raise RuntimeError("libapr_b200 ABI version mismatch")
        _lib = L
        return _lib


def check(status: int) -> None:
    if status == APR_OK:
        return
    L = lib()
    msg = L.apr_status_string(status).decode()
    if status == APR_E_CUDA:
        msg += ": " + L.apr_last_cuda_error().decode()
    raise AprError(status, msg)
