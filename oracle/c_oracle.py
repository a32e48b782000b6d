"""ctypes loader of oracle/_build/liboracle.so (the C/OpenMP restatement; test infrastructure, see apr_oracle_c.c)."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "liboracle.so")
_lib = None


def load(build: bool = True):
    global _lib
    if _lib is None:
        if build and (not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(os.path.join(HERE, "apr_oracle_c.c"))):
            subprocess.check_call(["make", "-s", "-C", HERE])
        L = ctypes.CDLL(SO)
        L.apr_oracle_threads.restype = ctypes.c_int
        L.apr_oracle_step.restype = ctypes.c_int
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def threads() -> int:
    return int(load().apr_oracle_threads())


def set_threads(n: int) -> int:
    """Use n OpenMP threads from now on (whatever OMP_NUM_THREADS says); returns the count in effect."""
    load().apr_oracle_set_threads(ctypes.c_int(int(n)))
    return threads()


def step(P, Q, aP, aQ, u, i, j, lr, reg, reg_adv, eps, adver) -> None:
    """In place on float32 C-contiguous tables; u, i, j int32 [B]."""
    u, i, j = [np.ascontiguousarray(x, dtype=np.int32).reshape(-1) for x in (u, i, j)]
    for t in (P, Q, aP, aQ):
        assert t.dtype == np.float32 and t.flags.c_contiguous
    rc = load().apr_oracle_step(_p(P), _p(Q), _p(aP), _p(aQ), ctypes.c_int(P.shape[1]), _p(u), _p(i), _p(j),
                                ctypes.c_int(u.size), ctypes.c_float(lr), ctypes.c_float(reg), ctypes.c_float(reg_adv),
                                ctypes.c_float(eps), ctypes.c_int(int(adver)))
    if rc:
        raise MemoryError("apr_oracle_step")


def positions(P, Q, users, test_item, num_items, excl_ptr, excl_idx) -> np.ndarray:
    users = np.ascontiguousarray(users, dtype=np.int32)
    test_item = np.ascontiguousarray(test_item, dtype=np.int32)
    excl_ptr = np.ascontiguousarray(excl_ptr, dtype=np.int64)
    excl_idx = np.ascontiguousarray(excl_idx, dtype=np.int32)
    out = np.zeros(users.size, dtype=np.int32)
    load().apr_oracle_positions(_p(P), _p(Q), ctypes.c_int(P.shape[1]), _p(users), _p(test_item), ctypes.c_int(users.size),
                                ctypes.c_int(num_items), _p(excl_ptr), _p(excl_idx), _p(out))
    return out
